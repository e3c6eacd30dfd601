"""CPU tests (no GPU) of oracle/_ref/libmgref.so -- the reference's own multigrid host code (lib/transfer.cpp, transfer_util.cu,
prolongator.cu, restrictor.cu, coarse_op.cu(h), coarsecoarse_op.cu, dslash_coarse.cu compiled unmodified) -- against the fine-grid
oracle that is bit-pinned to the reference's tests/wilson_dslash_reference.cpp.  They establish WHAT the reference's coarse operators are
(conventions of SURVEY Appendix A.8), so that the GPU parity tests (tests/test_mgref_gpu.py) compare against reference object code and not
against a restatement:
   full coarsening           ApplyCoarse(Y, X) = P^dag (tm_mat) P                         (lib/coarse_op.cuh:1309-1497, dslash_coarse.cu:263-290)
   preconditioned coarsening ApplyCoarse(Y, X) = P^dag A^-1 (tm_mat) P, a = -2 kappa mu   (computeTMAV :202-224, bi-directional links)
   coarse -> coarser         ApplyCoarse(Y2, X2) = P2^dag ApplyCoarse(Y, X) P2            (lib/coarsecoarse_op.cu:148-184)
"""
import numpy as np
import pytest

from tests import oracle_util as ou

X = (4, 4, 4, 8)
BS = (2, 2, 2, 2)
KAPPA, MU = 0.124, 0.05


@pytest.fixture(scope="module")
def mgref():
    r = ou.load_mgref()
    if r is None:
        pytest.skip("oracle/_ref/libmgref.so is absent and /root/reference is not here to build it")
    return r


def as_c(a):
    a = np.asarray(a, dtype=np.float64)
    return a[0::2] + 1j * a[1::2]


def to_r(z, dtype=np.float32):
    out = np.zeros(2 * z.size, dtype=dtype)
    out[0::2] = z.real; out[1::2] = z.imag
    return out


def ainv(z, a, V):
    v = z.reshape(V, 4, 3).copy()
    v[:, :2] *= (1 - 1j * a) / (1 + a * a)
    v[:, 2:] *= (1 + 1j * a) / (1 + a * a)
    return v.reshape(-1)


def dense(apply, n):
    M = np.zeros((0, n), dtype=np.complex128)
    cols = []
    e = np.zeros(2 * n, dtype=np.float32)
    for i in range(n):
        e[:] = 0; e[2 * i] = 1
        cols.append(as_c(apply(e)))
    return np.stack(cols, axis=1)


@pytest.fixture(scope="module")
def setup(oracle, mgref):
    oracle.set_dims(X)
    g = oracle.weak_gauge(eps=0.3, antiperiodic=True, seed=99)
    rng = np.random.default_rng(5)
    # n_vec = 24: the reference's restrictor handles the fine grid in groups of 8 coarse colours (coarse_colors_per_thread = 8 for
    # nColor = 3, lib/restrictor.cu:428) and runs past the end of its arrays for n_vec < 8, so only {24, 32} of its dispatch
    # list {2, 4, 24, 32} (:452-461) are usable on the fine level
    nvec = 24
    B = [rng.standard_normal(oracle.V * 24).astype(np.float32) for _ in range(nvec)]
    T = mgref.transfer(B, X, 4, 3, BS, 2)
    P = dense(lambda e: T.P(e), T.nc)
    return g, T, P, nvec


def test_reference_transfer_operator(oracle, mgref, setup):
    g, T, P, nvec = setup
    assert T.geo_bs == BS and T.Xc == (2, 2, 2, 4)
    # block-orthonormal columns (blockGramSchmidt, transfer_util.cu:327-363), R = P^dag (restrictor.cu:90-125 vs prolongator.cu:102-116)
    assert np.abs(P.conj().T @ P - np.eye(T.nc)).max() < 5e-6
    rng = np.random.default_rng(1)
    v = rng.standard_normal(2 * T.nf).astype(np.float32)
    assert ou.rel_l2(as_c(T.R(v)), P.conj().T @ as_c(v)) < 5e-6
    # single-parity transfers (Transfer::setSiteSubset, multigrid.cpp:300-309): the rows / columns of that parity
    for par in (0, 1):
        half = T.nf // 2
        sl = slice(par * half, (par + 1) * half)
        c = rng.standard_normal(2 * T.nc).astype(np.float32)
        assert ou.rel_l2(as_c(T.P(c, parity=par)), (P @ as_c(c))[sl]) < 5e-6
        assert ou.rel_l2(as_c(T.R(v[2 * sl.start:2 * sl.stop], parity=par)), P[sl].conj().T @ as_c(v)[sl]) < 5e-6
    # V in the reference's packed order [site][spin][colour][vector] holds the same numbers as the columns of P
    V = as_c(T.V()).reshape(oracle.V, 4, 3, nvec)
    Pm = P.reshape(oracle.V, 4, 3, -1)
    s = 7
    col = np.nonzero(np.abs(Pm[s, 0, 0]) > 0)[0]
    assert len(col) == nvec and np.allclose(Pm[s, 0, :, col].T, V[s, 0], atol=1e-7)


def test_reference_coarse_operator_is_the_galerkin_product(oracle, mgref, setup):
    g, T, P, nvec = setup
    a = 2 * KAPPA * MU
    co = mgref.coarse_op(T, g, KAPPA, a, "QUDA_TWISTED_MASS_DIRAC")
    Mc = dense(lambda e: co.apply(e, KAPPA), T.nc)
    MP = np.stack([as_c(oracle.tm_mat(g, to_r(P[:, i], np.float64), KAPPA, MU, 1, 0)) for i in range(T.nc)], axis=1)
    ref = P.conj().T @ MP
    err = np.abs(Mc - ref).max() / np.abs(ref).max()
    print(f"reference full coarsening vs P^dag tm_mat P: {err:.2e}")
    assert err < 2e-6
    # X * Xinv = 1 (Xinv comes from the Gauss-Jordan stand-in for MAGMA, see oracle/mg_ref_shim.cpp)
    Xl, Xi = co.links("X"), co.links("Xinv")
    assert np.abs(np.einsum("sij,sjk->sik", Xl, Xi) - np.eye(co.N)).max() < 1e-5
    co.free()


def test_reference_preconditioned_coarsening(oracle, mgref, setup):
    """DiracTwistedMassPC::createCoarseOp (lib/dirac_twisted_mass.cpp:572-576): a = -2 kappa mu, QUDA_TWISTED_MASSPC_DIRAC"""
    g, T, P, nvec = setup
    a = 2 * KAPPA * MU
    co = mgref.coarse_op(T, g, KAPPA, -a, "QUDA_TWISTED_MASSPC_DIRAC", "QUDA_MATPC_EVEN_EVEN")
    Mc = dense(lambda e: co.apply(e, KAPPA), T.nc)
    AMP = np.stack([ainv(as_c(oracle.tm_mat(g, to_r(P[:, i], np.float64), KAPPA, MU, 1, 0)), a, oracle.V) for i in range(T.nc)], axis=1)
    ref = P.conj().T @ AMP
    err = np.abs(Mc - ref).max() / np.abs(ref).max()
    print(f"reference preconditioned coarsening vs P^dag A^-1 tm_mat P: {err:.2e}")
    assert err < 2e-6
    co.free()


def test_reference_coarse_coarse_operator(oracle, mgref, setup):
    g, T, P, nvec = setup
    a = 2 * KAPPA * MU
    co = mgref.coarse_op(T, g, KAPPA, a, "QUDA_TWISTED_MASS_DIRAC")
    rng = np.random.default_rng(8)
    nv2 = 24
    B2 = [rng.standard_normal(2 * T.nc).astype(np.float32) for _ in range(nv2)]
    T2 = mgref.transfer(B2, T.Xc, 2, nvec, (2, 2, 2, 2), 1)
    # block-size fix-up of transfer.cpp:31-44: x cannot be blocked over its whole length, odd coarse extents are refused
    assert T2.geo_bs == (1, 1, 1, 2) and T2.Xc == (2, 2, 2, 2), (T2.geo_bs, T2.Xc)
    P2 = dense(lambda e: T2.P(e), T2.nc)
    Mc = dense(lambda e: co.apply(e, KAPPA), T.nc)
    for pc in (False, True):
        c2 = mgref.coarse_coarse_op(T2, co, KAPPA, pc=pc)
        M2 = dense(lambda e: c2.apply(e, KAPPA), T2.nc)
        if pc:   # DiracCoarsePC::createCoarseOp coarsens Yhat: the Galerkin product of Xinv M_c (X of the result gets the unit diagonal)
            Xi = co.links("Xinv")
            XM = Mc.reshape(co.V, co.N, -1)
            XM = np.einsum("sij,sjk->sik", Xi, XM).reshape(Mc.shape)
            ref = P2.conj().T @ XM @ P2
        else:
            ref = P2.conj().T @ Mc @ P2
        err = np.abs(M2 - ref).max() / np.abs(ref).max()
        print(f"reference coarse-coarse (pc={pc}): {err:.2e}")
        assert err < 5e-6
        c2.free()
    T2.free()
    co.free()
