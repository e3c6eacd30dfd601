"""Twisted-clover preconditioned hop A^-1 D at 32^3x64, fp32 recon-12: us per application (timeDslashQudaB200, CUDA events)."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quda_b200 as q
from tests import oracle_util as ou
o = ou.load_oracle(); X = (32, 32, 32, 64); o.set_dims(X)
g = o.gauge(1, True, 1.0, 137)
c = o.clover(norm=0.1, diag=1.0, seed=4242)
L = q.lib(); L.initQuda(0)
for prec in (4, 8):
    gp = q.gauge_param(X, cuda_prec=prec, reconstruct=12)
    L.loadGaugeQuda((C.c_void_p * 4)(*[a.ctypes.data for a in g]), C.byref(gp))
    def param(dag):
        p = q.invert_param(kappa=0.1, mu=0.01, cuda_prec=prec, dslash_type=q.QUDA_TWISTED_CLOVER_DSLASH, dagger=dag)
        p.clover_cpu_prec = 8
        p.clover_cuda_prec = p.clover_cuda_prec_sloppy = p.clover_cuda_prec_precondition = prec
        p.clover_order = q.QUDA_PACKED_CLOVER_ORDER; p.clover_coeff = 1.0
        p.compute_clover = p.compute_clover_inverse = p.return_clover = p.return_clover_inverse = 0
        return p
    p = param(0)
    L.loadCloverQuda(c.ctypes.data_as(C.c_void_p), None, C.byref(p))
    sp = o.drand(o.Vh * 24, 137)
    fi = L.newSpinorQudaB200(1, prec); fo = L.newSpinorQudaB200(1, prec)
    L.loadSpinorQudaB200(fi, sp.ctypes.data_as(C.c_void_p), C.byref(p))
    for dag in (0, 1):
        p = param(dag)
        L.timeDslashQudaB200(fo, fi, C.byref(p), 0, 5, None)
        ms = L.timeDslashQudaB200(fo, fi, C.byref(p), 0, 50, None)
        print("TMC prec=%d dagger=%d us=%.1f" % (prec, dag, ms * 1e3), flush=True)
    L.freeSpinorQudaB200(fi); L.freeSpinorQudaB200(fo)
L.endQuda()
