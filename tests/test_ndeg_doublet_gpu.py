"""Non-degenerate twisted-mass doublet (twist_flavor = QUDA_TWIST_NONDEG_DOUBLET) through the C ABI, as the reference's
tests/dslash_test.cpp does for its ndeg_tm branch (:270-275, :686-775): dslashQuda / MatQuda / MatDagMatQuda for every parity, matpc
type and dagger, and invertQuda solves, against the oracle's tm_ndeg_* (pinned bit for bit to the reference's objects).
Host fields: [parity][flavour][x_cb][spin][colour][re, im].  The two flavours go through one batched Wilson hop on the device.
Tolerances: fp64 <= 1e-13, fp32 <= 1e-6 relative L2."""
import ctypes as C

import numpy as np
import pytest

from tests.oracle_util import rel_l2

pytestmark = pytest.mark.gpu

KAPPA, MU, EPS = 0.1, 0.01, 0.03
TOL = {8: 1e-13, 4: 1e-6}


def vp(a):
    return a.ctypes.data_as(C.c_void_p)


def setup(quda, oracle, X, prec, recon=12, seed=137):
    oracle.set_dims(X)
    g = oracle.gauge(kind=1, antiperiodic=True, seed=seed)
    gp = quda.gauge_param(X, cuda_prec=prec, reconstruct=recon)
    quda.lib().loadGaugeQuda((C.c_void_p * 4)(*[a.ctypes.data for a in g]), C.byref(gp))
    sp = oracle.drand(4 * oracle.Vh * 24, seed=4711)   # full doublet field [even doublet | odd doublet]
    return g, sp


def param(q, prec, **kw):
    p = q.invert_param(kappa=KAPPA, mu=MU, cuda_prec=prec, flavor=q.QUDA_TWIST_NONDEG_DOUBLET, **kw)
    p.epsilon = EPS
    p.Ls = 2
    return p


@pytest.mark.parametrize("prec", [8, 4])
def test_ndeg_dslash_matpc_mat_all_variants(quda, oracle, prec):
    q, L = quda, quda.lib()
    X = (8, 8, 8, 8)
    g, sp = setup(q, oracle, X, prec)
    F = oracle.Vh * 24
    even = sp[: 2 * F].copy()
    worst = 0.0
    for matpc in range(4):
        for dag in (0, 1):
            for parity in (0, 1):
                p = param(q, prec, matpc=matpc, dagger=dag)
                out = np.zeros(2 * F)
                L.dslashQuda(vp(out), vp(even), C.byref(p), parity)
                err = rel_l2(out, oracle.tm_ndeg_dslash(g, even, KAPPA, MU, EPS, parity, matpc, dag))
                worst = max(worst, err)
                assert err <= TOL[prec], ("dslash", matpc, dag, parity, err)
            p = param(q, prec, matpc=matpc, dagger=dag, solution_type=q.QUDA_MATPC_SOLUTION)
            out = np.zeros(2 * F)
            L.MatQuda(vp(out), vp(even), C.byref(p))
            ref = oracle.tm_ndeg_matpc(g, even, KAPPA, MU, EPS, matpc, dag)
            err = rel_l2(out, ref)
            worst = max(worst, err)
            assert err <= TOL[prec], ("matpc", matpc, dag, err)
            # MatDagMat = M^dag M  (dslash_test type 3)
            p2 = param(q, prec, matpc=matpc, dagger=dag, solution_type=q.QUDA_MATPCDAG_MATPC_SOLUTION)
            L.MatDagMatQuda(vp(out), vp(even), C.byref(p2))
            ref2 = oracle.tm_ndeg_matpc(g, oracle.tm_ndeg_matpc(g, even, KAPPA, MU, EPS, matpc, dag), KAPPA, MU, EPS, matpc, 1 - dag)
            err = rel_l2(out, ref2)
            assert err <= 2 * TOL[prec], ("matdagmat", matpc, dag, err)
    for dag in (0, 1):
        p = param(q, prec, dagger=dag, solution_type=q.QUDA_MAT_SOLUTION)
        out = np.zeros(4 * F)
        L.MatQuda(vp(out), vp(sp), C.byref(p))
        err = rel_l2(out, oracle.tm_ndeg_mat(g, sp, KAPPA, MU, EPS, dag))
        worst = max(worst, err)
        assert err <= TOL[prec], ("mat", dag, err)
    print(f"ndeg doublet prec={prec}: worst rel-L2 = {worst:.2e}")


@pytest.mark.parametrize("solve,matpc", [("pc", 0), ("pc", 2), ("pc", 1), ("full", 0)])
def test_ndeg_invert(quda, oracle, solve, matpc):
    """invertQuda on the doublet: GCR, even-odd preconditioned (symmetric and asymmetric) and unpreconditioned, fp64 / fp32 sloppy;
    the solution is checked with the host operator."""
    q, L = quda, quda.lib()
    X = (8, 8, 8, 8)
    oracle.set_dims(X)
    g = oracle.weak_gauge(eps=0.3, antiperiodic=True, seed=17)
    gp = q.gauge_param(X, cuda_prec=8, reconstruct=12, cuda_prec_sloppy=4)
    L.loadGaugeQuda((C.c_void_p * 4)(*[a.ctypes.data for a in g]), C.byref(gp))
    kappa, mu, eps = 0.12, 0.05, 0.08
    b = np.random.default_rng(2).standard_normal(4 * oracle.Vh * 24)
    x = np.zeros_like(b)
    p = q.invert_param(kappa=kappa, mu=mu, cuda_prec=8, flavor=q.QUDA_TWIST_NONDEG_DOUBLET, matpc=matpc, solution_type=q.QUDA_MAT_SOLUTION)
    p.epsilon = eps; p.Ls = 2
    p.cuda_prec_sloppy = 4; p.cuda_prec_precondition = 4
    p.solve_type = q.QUDA_DIRECT_PC_SOLVE if solve == "pc" else q.QUDA_DIRECT_SOLVE
    p.inv_type = q.QUDA_GCR_INVERTER
    p.gcrNkrylov = 20; p.tol = 1e-9; p.maxiter = 2000; p.reliable_delta = 1e-4
    L.invertQuda(vp(x), vp(b), C.byref(p))
    res = np.linalg.norm(b - oracle.tm_ndeg_mat(g, x, kappa, mu, eps, 0)) / np.linalg.norm(b)
    print(f"ndeg invert {solve} matpc={matpc}: {p.iter} iterations, host residual {res:.2e}, reported {p.true_res:.2e}")
    assert res < 5e-9 and p.iter > 0


@pytest.mark.parametrize("mask", [8, 12, 15])
def test_ndeg_doublet_on_partitioned_lattice_self_exchange(mask):
    """The doublet on a partitioned lattice (halo path on one GPU, the reference's --partition trick): the flavours go through the
    pack / exchange / interior / boundary path one after the other.  Runs in a subprocess (partitioning is fixed at communicator set-up)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = f"""
import sys, ctypes as C, numpy as np
sys.path.insert(0, {root!r})
import quda_b200 as q
from tests import oracle_util as ou
o = ou.load_oracle(); X=(8,4,6,8); o.set_dims(X)
g = o.gauge(1, True, 1.0, 137); sp = o.drand(4*o.Vh*24, 4711); even = sp[:2*o.Vh*24].copy()
L = q.lib(); L.initQuda(0); L.commDimPartitionedSetQudaB200({mask})
gp = q.gauge_param(X, cuda_prec=8, reconstruct=12)
L.loadGaugeQuda((C.c_void_p*4)(*[a.ctypes.data for a in g]), C.byref(gp))
worst = 0.0
def par(**kw):
    p = q.invert_param(cuda_prec=8, flavor=q.QUDA_TWIST_NONDEG_DOUBLET, **kw); p.epsilon = 0.03; p.Ls = 2
    return p
for parity, matpc, dag in [(0,0,0),(1,0,1),(0,2,1),(1,3,0)]:
    p = par(matpc=matpc, dagger=dag)
    out = np.zeros(2*o.Vh*24)
    L.dslashQuda(out.ctypes.data_as(C.c_void_p), even.ctypes.data_as(C.c_void_p), C.byref(p), parity)
    worst = max(worst, ou.rel_l2(out, o.tm_ndeg_dslash(g, even, 0.1, 0.01, 0.03, parity, matpc, dag)))
p = par(solution_type=q.QUDA_MAT_SOLUTION)
out = np.zeros(2*o.V*24)
L.MatQuda(out.ctypes.data_as(C.c_void_p), sp.ctypes.data_as(C.c_void_p), C.byref(p))
worst = max(worst, ou.rel_l2(out, o.tm_ndeg_mat(g, sp, 0.1, 0.01, 0.03, 0)))
L.endQuda()
print("WORST", worst)
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert float(r.stdout.strip().split("WORST")[-1]) <= 1e-13
