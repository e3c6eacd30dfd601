"""Twisted-clover even-odd hop A^-1 D at 32^3x64 (GPU box): hop + inverse clover block fused into one launch against the two-launch form.
Usage: python tools/clover_bench.py ; prints one line per (precision, QB_FUSE_CLOVER)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, os, ctypes as C, numpy as np
sys.path.insert(0, %(root)r)
import quda_b200 as q
from tests import oracle_util as ou
o = ou.load_oracle(); X = (32, 32, 32, 64); o.set_dims(X)
g = o.gauge(1, True, 1.0, 137); sp = o.drand(o.Vh * 24, 137)
cl = o.clover(norm=0.1, diag=1.0, seed=4242)
L = q.lib(); L.initQuda(0)
for prec in (4, 2, 8):
    gp = q.gauge_param(X, cuda_prec=prec, reconstruct=12)
    L.loadGaugeQuda((C.c_void_p * 4)(*[a.ctypes.data for a in g]), C.byref(gp))
    p = q.invert_param(kappa=0.1, mu=0.05, cuda_prec=prec, dslash_type=q.QUDA_TWISTED_CLOVER_DSLASH)
    p.clover_cpu_prec = 8
    p.clover_cuda_prec = p.clover_cuda_prec_sloppy = p.clover_cuda_prec_precondition = prec
    p.clover_order = q.QUDA_PACKED_CLOVER_ORDER
    p.clover_coeff = 1.0
    p.compute_clover = p.compute_clover_inverse = p.return_clover = p.return_clover_inverse = 0
    L.loadCloverQuda(cl.ctypes.data_as(C.c_void_p), None, C.byref(p))
    fi = L.newSpinorQudaB200(1, prec); fo = L.newSpinorQudaB200(1, prec)
    L.loadSpinorQudaB200(fi, sp.ctypes.data_as(C.c_void_p), C.byref(p))
    L.timeDslashQudaB200(fo, fi, C.byref(p), 0, 10, None)
    ms = L.timeDslashQudaB200(fo, fi, C.byref(p), 0, 100, None)
    print("CLOVER fuse=%%s prec=%%d us=%%.2f" %% (os.environ.get("QB_FUSE_CLOVER", "1"), prec, ms * 1e3), flush=True)
    L.freeSpinorQudaB200(fi); L.freeSpinorQudaB200(fo)
L.endQuda()
"""

if __name__ == "__main__":
    for fuse in ("1", "0"):
        env = dict(os.environ, QB_FUSE_CLOVER=fuse)
        r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], env=env, capture_output=True, text=True)
        sys.stdout.write("".join(l + "\n" for l in r.stdout.splitlines() if l.startswith("CLOVER")))
        if r.returncode:
            sys.stdout.write(r.stderr[-2000:])
