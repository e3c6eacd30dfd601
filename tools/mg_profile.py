"""MG-GCR solve with the per-level wall-clock profile switched on (GPU box)."""
import ctypes as C, os, sys, time
import numpy as np
os.environ.setdefault("QUDA_B200_MG_PROFILE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quda_b200 as q
from tests import oracle_util as ou
import bench
o = ou.load_oracle()
L = q.lib(); L.initQuda(0)
L.setVerbosityQuda(q.QUDA_SUMMARIZE, b"", None)
res = bench.run_mg_leg(q, L, o, (32, 32, 32, 64), int(os.environ.get('QB_PRECOND', '4')))
print({k: v for k, v in res.items() if not isinstance(v, dict) or k.startswith('multi_src')})
L.endQuda()
