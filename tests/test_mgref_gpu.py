"""GPU multigrid pieces against the REFERENCE'S OWN object code: oracle/_ref/libmgref.so holds the
unmodified host paths of lib/transfer.cpp, transfer_util.cu (block Gram-Schmidt), prolongator.cu, restrictor.cu, coarse_op.cu(h)
(calculateY), coarsecoarse_op.cu and dslash_coarse.cu (CPU coarseDslash); tests/test_mgref.py pins that library to the fine oracle on the
CPU.  Here the GPU's near-null vectors are handed to the reference, which builds ITS transfer operator and coarse links from them; then
  a13  V / block orthogonalisation      P and R of both agree on random vectors and unit vectors
  a11, a12  prolongator / restrictor    same comparison, full and single-parity variants
  a9   fine -> coarse links             element-wise Y (8 directions), X after undoing the -kappa fold; full and preconditioned coarsening
  a10  coarse -> coarser links          level-2 operator vs the reference's CoarseCoarseOp fed with the GPU's level-1 null vectors
  a8   coarse Dslash (+ even-odd PC)    GPU M_c v and M_c,pc v vs the reference's ApplyCoarse on the reference's links
Tolerances are fp32 rounding through the block orthogonalisation (the reference does classical, the GPU modified Gram-Schmidt, both
with fp64 sums): 2e-5 relative, stated at each assert.  Xinv is the one piece without reference code (MAGMA, un-vendored)."""
import ctypes as C

import numpy as np
import pytest

from tests import oracle_util as ou
from tests.oracle_util import rel_l2
from tests.test_multigrid_gpu import coords_of, full_index, load_gauge, mg_inv_param, vp

pytestmark = pytest.mark.gpu
TOL = 2e-5


@pytest.fixture(scope="module")
def mgref():
    r = ou.load_mgref()
    if r is None:
        pytest.skip("oracle/_ref/libmgref.so is absent")
    return r


def as_c(a):
    a = np.asarray(a, dtype=np.float64)
    return a[0::2] + 1j * a[1::2]


def gpu_links(L, mg, level, which, V, N):
    nd = 1 if which == 1 else 9
    out = np.zeros(V * nd * N * N * 2, dtype=np.float32)
    L.mgCoarseLinksQudaB200(mg, level, which, vp(out))
    z = as_c(out).reshape(V, nd, N, N)
    return z[:, 0] if nd == 1 else z


def neighbour_table(Xc):
    """full index of x + mu for every coarse site"""
    V = int(np.prod(Xc)); Vh = V // 2
    nb = np.zeros((V, 4), dtype=np.int64)
    for par in (0, 1):
        for cb in range(Vh):
            x = list(coords_of(cb, par, Xc))
            for mu in range(4):
                y = list(x); y[mu] = (y[mu] + 1) % Xc[mu]
                nb[par * Vh + cb, mu] = full_index(*y, Xc)
    return nb


def compare_links(Lg, ref_co, kappa, Xc):
    """GPU links (row-major, -kappa folded, every link on its output site) against the reference's Y / X (lib/dslash_coarse.cu:49-203)"""
    Y, Xr = ref_co.links("Y"), ref_co.links("X")
    nb = neighbour_table(Xc)
    scale = np.abs(Y).max()
    errs = {}
    for mu in range(4):
        fwd = -Lg[:, 2 * mu] / kappa                                   # Y_{mu+4}(x)
        bwd = -np.conj(np.swapaxes(Lg[nb[:, mu], 2 * mu + 1], 1, 2)) / kappa   # Y_mu(x) = -L_{2mu+1}(x+mu)^dag / kappa
        errs[f"Y{mu + 4}"] = np.abs(fwd - Y[mu + 4]).max() / scale
        errs[f"Y{mu}"] = np.abs(bwd - Y[mu]).max() / scale
    errs["X"] = np.abs(Lg[:, 8] - Xr).max() / np.abs(Xr).max()
    return errs


def build_mg(q, oracle, X, blocks, nvecs, n_level, kappa, mu, pc, seed=99, eps=0.3, antiperiodic=True):
    L = q.lib()
    oracle.set_dims(X)
    g = oracle.weak_gauge(eps=eps, antiperiodic=antiperiodic, seed=seed)
    load_gauge(q, g, X, prec=8, sloppy=4, precond=4, recon=18, antiperiodic=antiperiodic)
    ip = mg_inv_param(q, kappa, mu)
    mgp = q.multigrid_param(ip, n_level=n_level, geo_block=blocks, n_vec=nvecs, setup_maxiter=30, setup_tol=1e-3,
                            solve_type=q.QUDA_DIRECT_PC_SOLVE if pc else q.QUDA_DIRECT_SOLVE)
    mg = L.newMultigridQuda(C.byref(mgp))
    return g, mg, mgp


def null_vectors(L, mg, level, n, length):
    B = []
    for k in range(n):
        b = np.zeros(length, dtype=np.float32)
        L.mgNullVectorQudaB200(mg, level, k, vp(b))
        B.append(b)
    return B


@pytest.mark.parametrize("pc", [False, True])
def test_two_level_pieces_against_reference_code(quda, oracle, mgref, pc):
    q, L = quda, quda.lib()
    X, bs, nvec, kappa, mu = (4, 4, 4, 8), (2, 2, 2, 2), 24, 0.124, 0.05
    a = 2 * kappa * mu
    g, mg, mgp = build_mg(q, oracle, X, (bs,), (nvec,), 2, kappa, mu, pc)
    Vf = int(np.prod(X))
    B = null_vectors(L, mg, 0, nvec, Vf * 24)
    T = mgref.transfer(B, X, 4, 3, bs, 2)
    assert T.geo_bs == tuple(mgp.geo_block_size[0][0:4])
    rng = np.random.default_rng(2)
    # ---- a11 / a12 / a13: P, R (and through them V) ----
    worst_p = worst_r = 0.0
    for trial in range(6):
        c = rng.standard_normal(2 * T.nc).astype(np.float32)
        if trial >= 3:   # unit vectors: single columns of V
            c[:] = 0; c[2 * rng.integers(T.nc)] = 1
        f = rng.standard_normal(2 * T.nf).astype(np.float32)
        fo = np.zeros(2 * T.nf, dtype=np.float32); co_ = np.zeros(2 * T.nc, dtype=np.float32)
        L.mgProlongQudaB200(mg, 0, vp(fo), vp(c))
        L.mgRestrictQudaB200(mg, 0, vp(co_), vp(f))
        worst_p = max(worst_p, rel_l2(fo, T.P(c)))
        worst_r = max(worst_r, rel_l2(co_, T.R(f)))
    print(f"pc={pc}: P vs reference Prolongate {worst_p:.2e}, R vs reference Restrict {worst_r:.2e}")
    assert worst_p < TOL and worst_r < TOL
    # ---- a9: coarse links element-wise ----
    if pc:
        ref_co = mgref.coarse_op(T, g, kappa, -a, "QUDA_TWISTED_MASSPC_DIRAC", "QUDA_MATPC_EVEN_EVEN")   # dirac_twisted_mass.cpp:572-576
    else:
        ref_co = mgref.coarse_op(T, g, kappa, a, "QUDA_TWISTED_MASS_DIRAC")                             # dirac_twisted_mass.cpp:224-228
    N = ref_co.N
    Lg = gpu_links(L, mg, 1, 0, ref_co.V, N)
    errs = compare_links(Lg, ref_co, kappa, T.Xc)
    print(f"pc={pc}: links vs reference calculateY (max |diff| / max |Y|): " + ", ".join(f"{k} {v:.1e}" for k, v in errs.items()))
    assert max(errs.values()) < TOL, errs
    # ---- a8: coarse Dslash, full operator and even-odd preconditioned operator ----
    worst = worst_pc = 0.0
    half = ref_co.V * N   # floats of a single-parity coarse field
    for trial in range(4):
        v = rng.standard_normal(2 * T.nc).astype(np.float32)
        out = np.zeros_like(v)
        L.mgMatQudaB200(mg, 1, 0, vp(out), vp(v))
        worst = max(worst, rel_l2(out, ref_co.apply(v, kappa)))
        # DiracCoarsePC::M, symmetric even-even (lib/dirac_coarse.cpp:245-283): out = in - Dhat_eo Dhat_oe in, Dhat = Yhat hop
        ve = v[:half].copy()
        t = ref_co.apply(ve, kappa, parity=1, dslash=True, clover=False, yhat=True)
        u = ref_co.apply(t, kappa, parity=0, dslash=True, clover=False, yhat=True)
        oe = np.zeros_like(ve)
        L.mgMatQudaB200(mg, 1, 1, vp(oe), vp(ve))
        worst_pc = max(worst_pc, rel_l2(oe, ve - u))
    print(f"pc={pc}: coarse operator vs reference ApplyCoarse {worst:.2e}; even-odd operator vs reference DiracCoarsePC::M {worst_pc:.2e}")
    assert worst < TOL and worst_pc < 5 * TOL   # the PC operator goes through Xinv twice
    ref_co.free(); T.free()
    L.destroyMultigridQuda(mg)


@pytest.mark.parametrize("pc", [False, True])
def test_three_level_coarse_coarse_against_reference_code(quda, oracle, mgref, pc):
    """a10: level-1 -> level-2 links against CoarseCoarseOp (lib/coarsecoarse_op.cu:148-184), full (Y, QUDA_COARSE_DIRAC) and
    preconditioned (Yhat, QUDA_COARSEPC_DIRAC, lib/dirac_coarse.cpp:377-380) coarsening"""
    q, L = quda, quda.lib()
    X, bs, nvec, kappa, mu = (8, 8, 8, 8), (2, 2, 2, 2), 24, 0.124, 0.05
    a = 2 * kappa * mu
    g, mg, mgp = build_mg(q, oracle, X, (bs, bs), (nvec, nvec), 3, kappa, mu, pc)
    Vf = int(np.prod(X))
    B0 = null_vectors(L, mg, 0, nvec, Vf * 24)
    T1 = mgref.transfer(B0, X, 4, 3, bs, 2)
    co1 = mgref.coarse_op(T1, g, kappa, -a if pc else a, "QUDA_TWISTED_MASSPC_DIRAC" if pc else "QUDA_TWISTED_MASS_DIRAC",
                          "QUDA_MATPC_EVEN_EVEN" if pc else "QUDA_MATPC_INVALID")
    N = co1.N
    errs1 = compare_links(gpu_links(L, mg, 1, 0, co1.V, N), co1, kappa, T1.Xc)
    assert max(errs1.values()) < TOL, errs1
    B1 = null_vectors(L, mg, 1, nvec, co1.V * N * 2)
    T2 = mgref.transfer(B1, T1.Xc, 2, nvec, bs, 1)
    assert T2.geo_bs == tuple(mgp.geo_block_size[1][0:4])
    rng = np.random.default_rng(3)
    c = rng.standard_normal(2 * T2.nc).astype(np.float32)
    f = rng.standard_normal(2 * T2.nf).astype(np.float32)
    fo = np.zeros(2 * T2.nf, dtype=np.float32); co_ = np.zeros(2 * T2.nc, dtype=np.float32)
    L.mgProlongQudaB200(mg, 1, vp(fo), vp(c))
    L.mgRestrictQudaB200(mg, 1, vp(co_), vp(f))
    ep, er = rel_l2(fo, T2.P(c)), rel_l2(co_, T2.R(f))
    co2 = mgref.coarse_coarse_op(T2, co1, kappa, pc=pc)
    errs2 = compare_links(gpu_links(L, mg, 2, 0, co2.V, co2.N), co2, kappa, T2.Xc)
    v = rng.standard_normal(2 * T2.nc).astype(np.float32)
    out = np.zeros_like(v)
    L.mgMatQudaB200(mg, 2, 0, vp(out), vp(v))
    eo = rel_l2(out, co2.apply(v, kappa))
    print(f"pc={pc}: level-1 P {ep:.2e} R {er:.2e}; level-2 links vs reference CoarseCoarseOp: " + ", ".join(f"{k} {v_:.1e}" for k, v_ in errs2.items()) +
          f"; level-2 operator vs ApplyCoarse {eo:.2e}")
    assert ep < TOL and er < TOL
    # the level-2 links inherit the rounding of the level-1 ones (built from two slightly different V): a few 1e-5
    assert max(errs2.values()) < 5 * TOL and eo < 5 * TOL, errs2
    co2.free(); T2.free(); co1.free(); T1.free()
    L.destroyMultigridQuda(mg)
