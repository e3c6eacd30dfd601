"""Where newMultigridQuda spends its time: builds the 3-level hierarchy of the bench (32^3x64, 4^4 / 2^4 aggregates, 24 vectors) several times with
verbosity QUDA_SUMMARIZE so that the per-level log lines (null vectors, coarse operator, level setup) are printed.  QB_PC=1: even-odd hierarchy."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quda_b200 as q  # noqa: E402
from tests import oracle_util as ou  # noqa: E402

X = (32, 32, 32, 64)
o = ou.load_oracle(); o.set_dims(X)
g = o.weak_gauge(eps=0.25, antiperiodic=False, seed=4711)
L = q.lib(); L.initQuda(0)
gp = q.gauge_param(X, cuda_prec=8, reconstruct=12, cuda_prec_sloppy=4, cuda_prec_precondition=4, t_boundary=q.QUDA_PERIODIC_T)
L.loadGaugeQuda((C.c_void_p * 4)(*[a.ctypes.data for a in g]), C.byref(gp))
for rep in range(3):
    for pc in ((0, 1) if os.environ.get("QB_PC", "both") == "both" else (int(os.environ["QB_PC"]),)):
        ip = q.invert_param(kappa=0.1248, mu=0.004, cuda_prec=8, solution_type=q.QUDA_MAT_SOLUTION)
        ip.cuda_prec_sloppy = 4; ip.cuda_prec_precondition = 4; ip.solve_type = q.QUDA_DIRECT_SOLVE; ip.inv_type = q.QUDA_GCR_INVERTER
        ip.verbosity = q.QUDA_SUMMARIZE
        mgp = q.multigrid_param(ip, n_level=3, geo_block=((4, 4, 4, 4), (2, 2, 2, 2)), n_vec=(24, 24), setup_maxiter=500, setup_tol=5e-6, run_verify=False,
                                solve_type=q.QUDA_DIRECT_PC_SOLVE if pc else q.QUDA_DIRECT_SOLVE)
        t0 = time.perf_counter()
        mg = L.newMultigridQuda(C.byref(mgp))
        print(f"SETUP rep {rep} pc {pc}: {time.perf_counter() - t0:.3f} s", flush=True)
        L.destroyMultigridQuda(mg)
L.endQuda()
