"""Times the multi-RHS tensor-core coarse operator against the single-RHS kernel on the level-1 grid of a 2-level hierarchy.
    python tools/mrhs_bench.py [X Y Z T] ; prints one line per (nrhs, mode)."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quda_b200 as q  # noqa: E402
from tests import oracle_util as ou  # noqa: E402

X = tuple(int(a) for a in sys.argv[1:5]) if len(sys.argv) >= 5 else (32, 32, 32, 64)
nvec = int(os.environ.get("QB_NVEC", "24"))
o = ou.load_oracle()
o.set_dims(X)
L = q.lib()
L.initQuda(0)
g = o.weak_gauge(eps=0.25, antiperiodic=False, seed=4711)
gp = q.gauge_param(X, cuda_prec=8, reconstruct=12, cuda_prec_sloppy=4, cuda_prec_precondition=4, t_boundary=q.QUDA_PERIODIC_T)
L.loadGaugeQuda((C.c_void_p * 4)(*[a.ctypes.data for a in g]), C.byref(gp))
ip = q.invert_param(kappa=0.1248, mu=0.004, cuda_prec=8, solution_type=q.QUDA_MAT_SOLUTION)
ip.cuda_prec_sloppy = 4; ip.cuda_prec_precondition = 4; ip.solve_type = q.QUDA_DIRECT_SOLVE; ip.inv_type = q.QUDA_GCR_INVERTER
mgp = q.multigrid_param(ip, n_level=2, geo_block=((4, 4, 4, 4),), n_vec=(nvec,), setup_maxiter=5, setup_tol=1e-1, run_verify=False)
mg = L.newMultigridQuda(C.byref(mgp))
info = (C.c_int * 8)()
L.mgLevelInfoQudaB200(mg, 0, info)
Vc, N = int(np.prod(info[0:4])), info[7]
niter = int(os.environ.get("QB_NITER", "50"))
t1 = L.mgTimeQudaB200(mg, 1, 0, niter)
link_bytes = Vc * 9 * N * N * 8
print(json.dumps({"coarse_sites": Vc, "N": N, "single_rhs_ms": t1, "single_rhs_gbs": (link_bytes + Vc * 10 * N * 8) / t1 / 1e6}))
MODES = [int(m) for m in os.environ.get("QB_MODES", "1,3").split(",")]
NRHS = [int(m) for m in os.environ.get("QB_NRHS", "1,4,8,12,16,24,32,48,64").split(",")]
for mode in MODES:
    for nrhs in NRHS:
        if nrhs > L.mgMrhsMaxRhsQudaB200(mg, 1, mode):
            continue
        ms = L.mgTimeMrhsQudaB200(mg, 1, 0, nrhs, mode, niter)
        flops = Vc * nrhs * (9 * 8 * N * N)
        byts = link_bytes + Vc * 2 * nrhs * N * 8  # links once + every vector read once and written once
        print(json.dumps({"mode": mode, "nrhs": nrhs, "ms": round(ms, 4), "ms_per_rhs": round(ms / nrhs, 4), "speedup_vs_single": round(t1 * nrhs / ms, 2),
                          "tflops": round(flops / ms / 1e9, 2), "hbm_gbs_compulsory": round(byts / ms / 1e6, 1)}))
L.destroyMultigridQuda(mg)
L.endQuda()
