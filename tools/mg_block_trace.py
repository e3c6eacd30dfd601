"""One block MG-GCR solve (invertMultiSrcQuda, QB_NSRC sources, default 12) at 32^3x64 bracketed by cudaProfilerStart/Stop, for
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ... python tools/mg_block_trace.py
(the launch list of the solve alone: which kernels the 0.13 s go to)."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quda_b200 as q
from tests import oracle_util as ou
o = ou.load_oracle()
L = q.lib(); L.initQuda(0)
X = (32, 32, 32, 64)
kappa, mu = 0.1248, 0.004
o.set_dims(X)
g = o.weak_gauge(eps=0.25, antiperiodic=False, seed=4711)
gp = q.gauge_param(X, cuda_prec=8, reconstruct=12, cuda_prec_sloppy=4, cuda_prec_precondition=4, t_boundary=q.QUDA_PERIODIC_T)
L.loadGaugeQuda((C.c_void_p * 4)(*[a.ctypes.data for a in g]), C.byref(gp))
def inv_param():
    p = q.invert_param(kappa=kappa, mu=mu, cuda_prec=8, solution_type=q.QUDA_MAT_SOLUTION)
    p.cuda_prec_sloppy = 4; p.cuda_prec_precondition = 4
    p.solve_type = q.QUDA_DIRECT_SOLVE; p.inv_type = q.QUDA_GCR_INVERTER
    p.gcrNkrylov = 20; p.tol = 1e-9; p.maxiter = 5000; p.reliable_delta = 1e-4
    return p
ip = inv_param()
mgp = q.multigrid_param(ip, n_level=3, geo_block=((4, 4, 4, 4), (2, 2, 2, 2)), n_vec=(24, 24), nu_pre=2, nu_post=2, setup_maxiter=500, setup_tol=5e-6, run_verify=False)
mg = L.newMultigridQuda(C.byref(mgp))
nsrc = int(os.environ.get("QB_NSRC", "12"))
bs = []
for k in range(nsrc):
    bk = np.zeros(o.V * 24); bk[2 * k] = 1.0
    bs.append(bk)
xs = [np.zeros(o.V * 24) for _ in range(nsrc)]
p = inv_param(); p.inv_type_precondition = q.QUDA_MG_INVERTER; p.preconditioner = mg; p.num_src = nsrc
px, pb = (C.c_void_p * nsrc)(*[a.ctypes.data for a in xs]), (C.c_void_p * nsrc)(*[a.ctypes.data for a in bs])
L.invertMultiSrcQuda(px, pb, C.byref(p))
rtl = C.CDLL("libcudart.so")
rtl.cudaProfilerStart()
L.invertMultiSrcQuda(px, pb, C.byref(p))
rtl.cudaProfilerStop()
print("SOLVE", p.iter, p.secs, p.true_res)
L.destroyMultigridQuda(mg); L.endQuda()
