"""One 12-source block MG-GCR solve at 32^3x64 on the even-odd hierarchy (for launch lists under ncu).  Usage: python tools/block_solve_once.py"""
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
    os.environ["QB_BENCH_BLOCK_ONLY"] = "1"
    r = bench.run_mg_leg(q, L, oracle, (32, 32, 32, 64), 4, half_storage=False, full=False, pc=True, multi_src=True)
    m = r["multi_src_12_point_sources"]["block"]
    print("BLOCK s_per_src=%.4f iters=%d" % (m["seconds_per_source"], m["iterations"]), flush=True)
    L.endQuda()
