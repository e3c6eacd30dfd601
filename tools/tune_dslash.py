"""Tuning sweep for the fine Dslash kernel (GPU box): block size x build variant x precision.
Usage: python tools/tune_dslash.py [lib_path ...]; prints one line per configuration."""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, os, ctypes as C, numpy as np
sys.path.insert(0, %(root)r)
import quda_b200 as q
from tests import oracle_util as ou
o = ou.load_oracle(); X=(32,32,32,64); o.set_dims(X)
g = o.gauge(1, True, 1.0, 137); sp = o.drand(o.Vh*24, 137)
L = q.lib(); L.initQuda(0)
for prec, recon in eval(os.environ.get('QB_TUNE_CASES', '((2,12),(4,12),(2,18),(4,18),(8,12),(4,8))')):
    gp = q.gauge_param(X, cuda_prec=prec, reconstruct=recon)
    L.loadGaugeQuda((C.c_void_p*4)(*[a.ctypes.data for a in g]), C.byref(gp))
    p = q.invert_param(cuda_prec=prec)
    fi = L.newSpinorQudaB200(1, prec); fo = L.newSpinorQudaB200(1, prec)
    L.loadSpinorQudaB200(fi, sp.ctypes.data_as(C.c_void_p), C.byref(p))
    for bs in (64, 128):
        L.setDslashBlockSizeQudaB200(bs)
        L.timeDslashQudaB200(fo, fi, C.byref(p), 0, 10, None)
        ms = L.timeDslashQudaB200(fo, fi, C.byref(p), 0, int(os.environ.get('QB_TUNE_ITER', '100')), None)
        print("TUNE lib=%%s prec=%%d recon=%%d block=%%d us=%%.2f" %% (os.path.basename(q.LIB_PATH), prec, recon, bs, ms*1e3), flush=True)
    L.freeSpinorQudaB200(fi); L.freeSpinorQudaB200(fo)
L.endQuda()
"""

if __name__ == "__main__":
    libs = sys.argv[1:] or [os.path.join(ROOT, "quda-qkxtm-multigrid_b200", "libquda_b200.so")]
    for lib in libs:
        env = dict(os.environ, QUDA_B200_LIB=os.path.abspath(lib))
        r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], env=env, capture_output=True, text=True)
        sys.stdout.write("".join(l + "\n" for l in r.stdout.splitlines() if l.startswith("TUNE")))
        if r.returncode:
            sys.stdout.write(r.stderr[-2000:])
