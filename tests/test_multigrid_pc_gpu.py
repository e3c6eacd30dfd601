"""Multigrid on the even-odd preconditioned system -- the reference's DEFAULT path: tests/test_util.cpp:1600 (solve_type =
QUDA_DIRECT_PC_SOLVE), tests/multigrid_invert_test.cpp:252 (coarse_grid_solution_type = QUDA_MATPC_SOLUTION), lib/multigrid.cpp:145-155
(preconditioned_coarsen picks matSmooth), lib/dirac_twisted_mass.cpp:580 (DiracTwistedMassPC::createCoarseOp), lib/coarse_op.cuh:202-224
(computeTMAV), :1349-1440 (bi-directional links), lib/dirac_coarse.cpp:377-380 (DiracCoarsePC::createCoarseOp coarsens Yhat) and the
single-parity cycle lib/multigrid.cpp:494-560.

What the reference's preconditioned coarsening builds (read off computeTMAV / computeUV / multiplyVUV): Y_fwd = (A^-1 V)^dag-free form
V^dag A^-1 (1 - gamma_mu) U V(x+mu), Y_bwd^dag = V^dag A^-1 (1 + gamma_mu) U^dag V(x-mu), X = 1 - kappa (local hops): the Galerkin product
of  A^-1 M = 1 - kappa A^-1 D  on the FULL lattice.  Its even-odd Schur complement is the symmetric preconditioned operator
M_pc = 1 - kappa^2 A^-1 D A^-1 D of the oracle's tm_matpc -- both facts are checked here against the bit-pinned oracle."""
import ctypes as C

import numpy as np
import pytest

from tests.oracle_util import rel_l2
from tests.test_multigrid_gpu import as_c, host_residual, load_gauge, mg_inv_param, point_source, vp

pytestmark = pytest.mark.gpu


def to_reals(z):
    out = np.zeros(2 * z.size)
    out[0::2] = z.real; out[1::2] = z.imag
    return out


def ainv(z, a, V):
    """A^-1 = (1 - i a gamma5) / (1 + a^2), gamma5 = diag(1, 1, -1, -1) in the DeGrand-Rossi basis (wilson_dslash_reference.cpp:233-263)"""
    v = z.reshape(V, 4, 3).copy()
    v[:, :2] *= (1 - 1j * a) / (1 + a * a)
    v[:, 2:] *= (1 + 1j * a) / (1 + a * a)
    return v.reshape(-1)


def build(q, oracle, X, blocks, nvecs, n_level, kappa, mu, eps, matpc=None, setup_maxiter=200, setup_tol=5e-6, seed=4711, antiperiodic=False, pc=True):
    L = q.lib()
    oracle.set_dims(X)
    g = oracle.weak_gauge(eps=eps, antiperiodic=antiperiodic, seed=seed)
    load_gauge(q, g, X, prec=8, sloppy=4, precond=4, recon=12, antiperiodic=antiperiodic)
    ip = mg_inv_param(q, kappa, mu)
    if matpc is not None:
        ip.matpc_type = matpc
    mgp = q.multigrid_param(ip, n_level=n_level, geo_block=blocks, n_vec=nvecs, setup_maxiter=setup_maxiter, setup_tol=setup_tol,
                            solve_type=q.QUDA_DIRECT_PC_SOLVE if pc else q.QUDA_DIRECT_SOLVE)
    mg = L.newMultigridQuda(C.byref(mgp))
    return g, mg, mgp, ip


@pytest.mark.parametrize("nvec", [4, 8])
def test_preconditioned_coarse_operator_is_galerkin_product_of_ainv_m(quda, oracle, nvec):
    """dense check on a small lattice: M_c = P^dag (A^-1 M_oracle) P; and A^-1 M_oracle's even-odd Schur complement IS the oracle's tm_matpc"""
    q, L = quda, quda.lib()
    X, bs, kappa, mu = (4, 4, 4, 8), (2, 2, 2, 2), 0.124, 0.05
    a = 2 * kappa * mu
    g, mg, mgp, ip = build(q, oracle, X, (bs,), (nvec,), 2, kappa, mu, 0.3, setup_maxiter=30, setup_tol=1e-3, seed=99, antiperiodic=True)
    info = (C.c_int * 8)()
    L.mgLevelInfoQudaB200(mg, 0, info)
    N = info[7]
    Vf, Vc = int(np.prod(X)), int(np.prod(info[0:4]))
    nf, nc = Vf * 12, Vc * N
    dev = (C.c_double * 3)()
    L.mgVerifyQudaB200(mg, 0, dev)   # identity (3) is R A^-1 M P eta = M_c eta here
    assert dev[0] < 2e-6 and dev[1] < 2e-5 and dev[2] < 2e-5, list(dev)
    P = np.zeros((nf, nc), dtype=np.complex128)
    Mc = np.zeros((nc, nc), dtype=np.complex128)
    e = np.zeros(2 * nc, dtype=np.float32)
    out = np.zeros(2 * nf, dtype=np.float32)
    oc = np.zeros(2 * nc, dtype=np.float32)
    for i in range(nc):
        e[:] = 0; e[2 * i] = 1
        L.mgProlongQudaB200(mg, 0, vp(out), vp(e))
        P[:, i] = as_c(out.astype(np.float64))
        L.mgMatQudaB200(mg, 1, 0, vp(oc), vp(e))
        Mc[:, i] = as_c(oc.astype(np.float64))
    AMP = np.zeros((nf, nc), dtype=np.complex128)
    for i in range(nc):
        AMP[:, i] = ainv(as_c(oracle.tm_mat(g, to_reals(P[:, i]), kappa, mu, 1, 0)), a, Vf)
    ref = P.conj().T @ AMP
    err = np.abs(Mc - ref).max() / np.abs(ref).max()
    print(f"n_vec {nvec}: |M_c - P^dag A^-1 M P|_max / |.|_max = {err:.2e}")
    assert err < 1e-5
    # the unpreconditioned Galerkin product is a different matrix (the test would not notice a missing A^-1 otherwise)
    MP = np.stack([as_c(oracle.tm_mat(g, to_reals(P[:, i]), kappa, mu, 1, 0)) for i in range(nc)], axis=1)
    assert np.abs(Mc - P.conj().T @ MP).max() / np.abs(ref).max() > 1e-3
    # Schur complement of A^-1 M on the even sites = tm_matpc (symmetric, even-even) of the oracle
    rng = np.random.default_rng(3)
    ne = nf // 2
    ve = rng.standard_normal(ne) + 1j * rng.standard_normal(ne)
    w = as_c(oracle.tm_mat(g, to_reals(np.concatenate([ve, np.zeros(ne)])), kappa, mu, 1, 0))   # (A v_e, -kappa D_oe v_e)
    uo = -ainv(np.concatenate([np.zeros(ne), w[ne:]]), a, Vf)[ne:]                               # makes the odd rows of A^-1 M u vanish
    full = ainv(as_c(oracle.tm_mat(g, to_reals(np.concatenate([ve, uo])), kappa, mu, 1, 0)), a, Vf)
    assert np.linalg.norm(full[ne:]) < 1e-13 * np.linalg.norm(full[:ne])
    assert rel_l2(full[:ne], as_c(oracle.tm_matpc(g, to_reals(ve), kappa, mu, 1, 0, 0))) < 1e-13
    L.destroyMultigridQuda(mg)


def pc_solve(q, L, oracle, g, mg, kappa, mu, b, solution_type, matpc=None, tol=1e-8, use_mg=True, prec_sloppy=4):
    p = mg_inv_param(q, kappa, mu, sloppy=prec_sloppy)
    p.solve_type = q.QUDA_DIRECT_PC_SOLVE
    p.solution_type = solution_type
    if matpc is not None:
        p.matpc_type = matpc
    if use_mg:
        p.inv_type_precondition = q.QUDA_MG_INVERTER
        p.preconditioner = mg
    p.gcrNkrylov = 20; p.tol = tol; p.maxiter = 4000 if not use_mg else 200; p.reliable_delta = 1e-4
    x = np.zeros_like(b)
    L.invertQuda(vp(x), vp(b), C.byref(p))
    return x, p


@pytest.mark.parametrize("n_level,X,blocks,nvecs", [(2, (8, 8, 8, 16), ((4, 4, 4, 4),), (8,)),
                                                    (3, (16, 16, 16, 16), ((4, 4, 4, 4), (2, 2, 2, 2)), (8, 8))])
def test_mg_gcr_on_the_even_odd_system(quda, oracle, n_level, X, blocks, nvecs):
    """multigrid_invert_test with its default parameters: outer GCR on the even-odd system (QUDA_DIRECT_PC_SOLVE, MAT solution),
    every level injecting single-parity fields (MATPC); host residual of the FULL system as multigrid_invert_test.cpp:529-577"""
    q, L = quda, quda.lib()
    kappa, mu = 0.1248, 0.004
    g, mg, mgp, ip = build(q, oracle, X, blocks, nvecs, n_level, kappa, mu, 0.25)
    for l in range(n_level - 1):
        dev = (C.c_double * 3)()
        L.mgVerifyQudaB200(mg, l, dev)
        assert dev[0] < 5e-6 and dev[1] < 1e-4 and dev[2] < 5e-5, (l, list(dev))
    b = point_source(oracle.V)
    x, p = pc_solve(q, L, oracle, g, mg, kappa, mu, b, q.QUDA_MAT_SOLUTION)
    res = host_residual(oracle, g, x, b, kappa, mu)
    x0, p0 = pc_solve(q, L, oracle, g, mg, kappa, mu, b, q.QUDA_MAT_SOLUTION, use_mg=False)
    print(f"{n_level}-level MG on the even-odd system: {p.iter} iterations ({p.secs:.3f} s) vs plain even-odd GCR {p0.iter} ({p0.secs:.3f} s); host residual {res:.2e}, setup {mgp.secs:.2f} s")
    assert res < 5e-8 and p.true_res < 2e-8
    assert p.iter < p0.iter / 3, (p.iter, p0.iter)
    assert host_residual(oracle, g, x0, b, kappa, mu) < 5e-8
    # MATPC solution: x_e solves M_pc x_e = b_e (oracle tm_matpc)
    be = oracle.drand(oracle.Vh * 24, seed=9)
    xe, pe = pc_solve(q, L, oracle, g, mg, kappa, mu, be, q.QUDA_MATPC_SOLUTION)
    rpc = np.linalg.norm(be - oracle.tm_matpc(g, xe, kappa, mu, 1, 0, 0)) / np.linalg.norm(be)
    assert rpc < 5e-8 and pe.iter <= p.iter + 3, (rpc, pe.iter, p.iter)
    # a hierarchy built for full-field injection refuses the even-odd outer solve only through an error; the converse works:
    # full outer solve (QUDA_DIRECT_SOLVE) with this hierarchy reduces exactly to the even-odd system inside the cycle
    pf = mg_inv_param(q, kappa, mu)
    pf.inv_type_precondition = q.QUDA_MG_INVERTER; pf.preconditioner = mg
    pf.gcrNkrylov = 20; pf.tol = 1e-8; pf.maxiter = 200; pf.reliable_delta = 1e-4
    xf = np.zeros_like(b)
    L.invertQuda(vp(xf), vp(b), C.byref(pf))
    assert host_residual(oracle, g, xf, b, kappa, mu) < 5e-8 and pf.iter < p0.iter / 3
    L.destroyMultigridQuda(mg)


def test_mg_on_the_odd_odd_system(quda, oracle):
    """matpc_type = QUDA_MATPC_ODD_ODD: transfer parity, coarse even-odd systems and the outer solve all live on the odd sites
    (multigrid.cpp:300-309)"""
    q, L = quda, quda.lib()
    X, kappa, mu = (8, 8, 8, 16), 0.1245, 0.005
    g, mg, mgp, ip = build(q, oracle, X, ((4, 4, 4, 4),), (8,), 2, kappa, mu, 0.25, matpc=q.QUDA_MATPC_ODD_ODD)
    b = point_source(oracle.V)
    x, p = pc_solve(q, L, oracle, g, mg, kappa, mu, b, q.QUDA_MAT_SOLUTION, matpc=q.QUDA_MATPC_ODD_ODD)
    res = host_residual(oracle, g, x, b, kappa, mu)
    x0, p0 = pc_solve(q, L, oracle, g, mg, kappa, mu, b, q.QUDA_MAT_SOLUTION, matpc=q.QUDA_MATPC_ODD_ODD, use_mg=False)
    print(f"odd-odd: {p.iter} MG iterations vs {p0.iter} plain, host residual {res:.2e}")
    assert res < 5e-8 and p.iter < p0.iter / 3
    L.destroyMultigridQuda(mg)


@pytest.mark.parametrize("n_level,X,blocks,nvecs", [(2, (8, 8, 8, 16), ((4, 4, 4, 4),), (8,)),
                                                    (3, (16, 16, 16, 16), ((4, 4, 4, 4), (2, 2, 2, 2)), (8, 8))])
def test_block_mg_multi_src_on_the_even_odd_system(quda, oracle, n_level, X, blocks, nvecs, monkeypatch):
    """invertMultiSrcQuda with the reference's default solve type (QUDA_DIRECT_PC_SOLVE) on a hierarchy coarsened on the even-odd
    system: all sources in lock-step through the single-parity block K-cycle (coarse levels on the multi-RHS tensor-core operator);
    every solution checked with the host operator and against the one-source-at-a-time path"""
    q, L = quda, quda.lib()
    kappa, mu, tol = 0.1248, 0.004, 1e-8
    g, mg, mgp, ip = build(q, oracle, X, blocks, nvecs, n_level, kappa, mu, 0.25)
    nsrc = 5
    rng = np.random.default_rng(11)
    bs = [point_source(oracle.V)] + [rng.standard_normal(oracle.V * 24) for _ in range(nsrc - 1)]

    def solve(block):
        monkeypatch.setenv("QB_BLOCK_MG", "1" if block else "0")
        p = mg_inv_param(q, kappa, mu)
        p.solve_type = q.QUDA_DIRECT_PC_SOLVE
        p.inv_type_precondition = q.QUDA_MG_INVERTER
        p.preconditioner = mg
        p.gcrNkrylov = 20; p.tol = tol; p.maxiter = 200; p.reliable_delta = 1e-4
        p.num_src = nsrc
        xs = [np.zeros(oracle.V * 24) for _ in range(nsrc)]
        L.invertMultiSrcQuda((C.c_void_p * nsrc)(*[a.ctypes.data for a in xs]), (C.c_void_p * nsrc)(*[a.ctypes.data for a in bs]), C.byref(p))
        return xs, p.iter, p.true_res, p.secs

    xb, it_b, tr_b, t_b = solve(True)
    xs, it_s, tr_s, t_s = solve(False)
    L.destroyMultigridQuda(mg)
    worst = max(host_residual(oracle, g, x, b, kappa, mu) for x, b in zip(xb, bs))
    print(f"even-odd block MG ({n_level} levels): {it_b} lock-step iterations in {t_b:.3f} s, sequential {it_s} iterations (sum over {nsrc}) in {t_s:.3f} s; "
          f"worst host residual {worst:.2e}")
    assert worst < 5e-8 and tr_b < 5e-8 and tr_s < 5e-8
    assert it_b < it_s, (it_b, it_s)                       # lock-step count, not the sum: the block path was taken
    assert it_b <= 1.5 * it_s / nsrc + 3, (it_b, it_s)
    for a, b_ in zip(xb, xs):
        assert np.linalg.norm(a - b_) / np.linalg.norm(b_) < 1e-6
