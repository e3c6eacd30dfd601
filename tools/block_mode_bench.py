"""12-source block MG-GCR at 32^3x64 (GPU box): split-tf32 (mode 3, fp32-accurate) against plain tf32 (mode 1) in the multi-RHS tensor-core
coarse operator, fp32 and fp16 preconditioner storage.  Usage: python tools/block_mode_bench.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import quda_b200 as q  # noqa: E402
from tests import oracle_util as ou  # noqa: E402

if __name__ == "__main__":
    L = q.lib()
    L.initQuda(0)
    oracle = ou.load_oracle()
    X = (32, 32, 32, 64)
    for half_storage in (False, True):
        for mode in ("3", "1"):
            os.environ["QB_BLOCK_MG_MODE"] = mode
            r = bench.run_mg_leg(q, L, oracle, X, 4, half_storage=half_storage, full=False, pc=True, multi_src=True)
            m = r["multi_src_12_point_sources"]["block"]
            print("BLOCKMODE half_storage=%d mode=%s solve1=%.4f iters1=%d block_s_per_src=%.4f block_iters=%d worst_res=%.2e" %
                  (half_storage, mode, r["solve_seconds"], r["iterations"], m["seconds_per_source"], m["iterations"], m["worst_true_res"]), flush=True)
    L.endQuda()
