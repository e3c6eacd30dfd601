"""Batched (multi-RHS) fine hop: ms per launch and per member against the single-field kernel (GPU box)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quda_b200 as q
from tests import oracle_util as ou
o = ou.load_oracle(); X = (32, 32, 32, 64); o.set_dims(X)
g = o.gauge(1, True, 1.0, 137)
L = q.lib(); L.initQuda(0)
gp = q.gauge_param(X, cuda_prec=4, reconstruct=12)
L.loadGaugeQuda((C.c_void_p * 4)(*[a.ctypes.data for a in g]), C.byref(gp))
p = q.invert_param(cuda_prec=4)
for n in (1, 2, 4, 6, 8, 12):
    ms = L.timeDslashBatchQudaB200(C.byref(p), 0, n, 50, None)
    print("BATCH nbatch=%d ms=%.4f per_member_us=%.2f compulsory_GBs=%.0f" % (n, ms, ms * 1e3 / n, (384 + 192 * n) * o.Vh / ms / 1e6), flush=True)
L.endQuda()
