"""Times the level-0 coarse-link build only (2-level setup with a token null-vector solve) at the bench lattice."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quda_b200 as q
from tests import oracle_util as ou
X = tuple(int(a) for a in os.environ.get("QB_X", "32,32,32,64").split(","))
o = ou.load_oracle(); o.set_dims(X)
L = q.lib(); L.initQuda(0)
g = o.weak_gauge(eps=0.25, antiperiodic=False, seed=4711)
gp = q.gauge_param(X, cuda_prec=8, reconstruct=12, cuda_prec_sloppy=4, cuda_prec_precondition=4, t_boundary=q.QUDA_PERIODIC_T)
L.loadGaugeQuda((C.c_void_p * 4)(*[a.ctypes.data for a in g]), C.byref(gp))
ip = q.invert_param(kappa=0.1248, mu=0.004, cuda_prec=8, solution_type=q.QUDA_MAT_SOLUTION)
ip.cuda_prec_sloppy = 4; ip.cuda_prec_precondition = 4; ip.solve_type = q.QUDA_DIRECT_SOLVE; ip.inv_type = q.QUDA_GCR_INVERTER
ip.verbosity = 1
mgp = q.multigrid_param(ip, n_level=2, geo_block=((4, 4, 4, 4),), n_vec=(int(os.environ.get("QB_NVEC", "24")),), setup_maxiter=2, setup_tol=1e-1, run_verify=False)
mg = L.newMultigridQuda(C.byref(mgp))
L.destroyMultigridQuda(mg)
L.endQuda()
