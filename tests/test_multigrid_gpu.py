"""GPU tests of the multigrid path through the C ABI (newMultigridQuda / invertQuda) and the test hooks of
include/quda_b200_ext.h.  Modelled on tests/multigrid_invert_test.cpp (host residual check :529-577) and on
the identities of MG::verify (lib/multigrid.cpp:372-486).  Oracles:
  * fine operator: oracle/tm_oracle.c (tm_mat), bit-pinned to the reference's CPU code;
  * transfer operator / Galerkin coarse operator / Schur complement: numpy restatements below of
    lib/transfer.cpp:220-258 (geo map), lib/prolongator.cu:41-56, lib/restrictor.cu:90-125,
    lib/coarse_op.cuh (M_c = R M P) and lib/dirac_coarse.cpp:250-283 (DiracCoarsePC::M).
"""
import ctypes as C

import numpy as np
import pytest

from tests.oracle_util import rel_l2

pytestmark = pytest.mark.gpu


def vp(a):
    return a.ctypes.data_as(C.c_void_p)


def load_gauge(q, g, X, prec=8, sloppy=4, precond=4, recon=18, antiperiodic=False):
    gp = q.gauge_param(X, cuda_prec=prec, reconstruct=recon, cuda_prec_sloppy=sloppy, cuda_prec_precondition=precond,
                       t_boundary=q.QUDA_ANTI_PERIODIC_T if antiperiodic else q.QUDA_PERIODIC_T)
    q.lib().loadGaugeQuda((C.c_void_p * 4)(*[a.ctypes.data for a in g]), C.byref(gp))


def mg_inv_param(q, kappa, mu, prec=8, sloppy=4, precond=4):
    p = q.invert_param(kappa=kappa, mu=mu, cuda_prec=prec, solution_type=q.QUDA_MAT_SOLUTION)
    p.cuda_prec_sloppy = sloppy
    p.cuda_prec_precondition = precond
    p.solve_type = q.QUDA_DIRECT_SOLVE
    p.inv_type = q.QUDA_GCR_INVERTER
    p.verbosity = q.QUDA_SILENT
    return p


def coords_of(cb, parity, X):
    za = cb // (X[0] // 2); zb = za // X[1]
    y = za - zb * X[1]; t = zb // X[2]; z = zb - t * X[2]
    x = 2 * cb + ((y + z + t + parity) & 1) - za * X[0]
    return x, y, z, t


def full_index(x, y, z, t, X):
    Vh = X[0] * X[1] * X[2] * X[3] // 2
    lex = ((t * X[2] + z) * X[1] + y) * X[0] + x
    return ((x + y + z + t) & 1) * Vh + (lex >> 1)


def as_c(a):
    a = np.asarray(a)
    return a[..., 0::2] + 1j * a[..., 1::2]


@pytest.mark.parametrize("nvec", [4, 8])
def test_transfer_and_galerkin_coarse_operator(quda, oracle, nvec):
    """n_vec = 4: CUDA-core coarse-link build (coarse_op.cu); n_vec = 8: tensor-core build (coarse_op_mma.cu, tcgen05 split tf32)."""
    q, L = quda, quda.lib()
    X = (4, 4, 4, 8)
    bs = (2, 2, 2, 2)
    kappa, mu = 0.124, 0.05
    oracle.set_dims(X)
    g = oracle.weak_gauge(eps=0.3, antiperiodic=True, seed=99)
    load_gauge(q, g, X, antiperiodic=True)
    ip = mg_inv_param(q, kappa, mu)
    mgp = q.multigrid_param(ip, n_level=2, geo_block=(bs,), n_vec=(nvec,), setup_maxiter=30, setup_tol=1e-3)
    mg = L.newMultigridQuda(C.byref(mgp))
    assert mg and mgp.secs > 0
    info = (C.c_int * 8)()
    L.mgLevelInfoQudaB200(mg, 0, info)
    Xc = tuple(info[0:4]); N = info[7]
    assert Xc == (2, 2, 2, 4) and info[4] == nvec and info[5] == 12 and info[6] == 16 and N == 2 * nvec
    Vf, Vc = int(np.prod(X)), int(np.prod(Xc))
    nf, nc = Vf * 12, Vc * N

    dev = (C.c_double * 3)()
    L.mgVerifyQudaB200(mg, 0, dev)
    assert dev[0] < 2e-6 and dev[1] < 2e-5 and dev[2] < 2e-5, list(dev)

    # dense P from unit coarse vectors
    P = np.zeros((nf, nc), dtype=np.complex128)
    e = np.zeros(2 * nc, dtype=np.float32)
    out = np.zeros(2 * nf, dtype=np.float32)
    for i in range(nc):
        e[:] = 0; e[2 * i] = 1
        L.mgProlongQudaB200(mg, 0, vp(out), vp(e))
        P[:, i] = as_c(out.astype(np.float64))
    # (a) block orthonormality  P^dag P = 1
    assert np.abs(P.conj().T @ P - np.eye(nc)).max() < 5e-6
    # (b) support: column (X, S, j) lives on the aggregate X (coords // block) and chirality S (spin // 2)
    Vhf, Vhc = Vf // 2, Vc // 2
    agg = np.zeros(Vf, dtype=np.int64)
    for par in (0, 1):
        for cb in range(Vhf):
            x, y, z, t = coords_of(cb, par, X)
            agg[par * Vhf + cb] = full_index(x // bs[0], y // bs[1], z // bs[2], t // bs[3], Xc)
    Pm = np.abs(P).reshape(Vf, 4, 3, Vc, 2, nvec)
    for s in range(4):
        for S in range(2):
            blk = Pm[:, s, :, :, S, :]
            if S != s // 2:
                assert blk.max() == 0.0
            else:
                mask = (agg[:, None] == np.arange(Vc)[None, :])
                assert (blk.max(axis=(1, 3))[~mask] == 0).all()
    # (c) R = P^dag
    rng = np.random.default_rng(1)
    v = rng.standard_normal(2 * nf).astype(np.float32)
    rc = np.zeros(2 * nc, dtype=np.float32)
    L.mgRestrictQudaB200(mg, 0, vp(rc), vp(v))
    assert rel_l2(as_c(rc.astype(np.float64)), P.conj().T @ as_c(v.astype(np.float64))) < 5e-6
    # (d) the null vectors lie in the range of P
    for k in range(nvec):
        L.mgNullVectorQudaB200(mg, 0, k, vp(out))
        b = as_c(out.astype(np.float64))
        assert np.linalg.norm(P @ (P.conj().T @ b) - b) / np.linalg.norm(b) < 2e-5
    # (e) Galerkin: M_c = P^dag M P with M the reference CPU operator (oracle tm_mat, fp64)
    Mc = np.zeros((nc, nc), dtype=np.complex128)
    oc = np.zeros(2 * nc, dtype=np.float32)
    for i in range(nc):
        e[:] = 0; e[2 * i] = 1
        L.mgMatQudaB200(mg, 1, 0, vp(oc), vp(e))
        Mc[:, i] = as_c(oc.astype(np.float64))
    MP = np.zeros((nf, nc), dtype=np.complex128)
    for i in range(nc):
        col = np.zeros(2 * nf)
        col[0::2] = P[:, i].real; col[1::2] = P[:, i].imag
        MP[:, i] = as_c(oracle.tm_mat(g, col, kappa, mu, 1, 0))
    ref = P.conj().T @ MP
    assert np.abs(Mc - ref).max() / np.abs(ref).max() < 1e-5
    # (f) even-odd preconditioned coarse operator = Schur complement of M_c (uses Xinv)
    ne = Vhc * N
    Mee, Meo, Moe, Moo = Mc[:ne, :ne], Mc[:ne, ne:], Mc[ne:, :ne], Mc[ne:, ne:]
    Xe = np.zeros_like(Mee); Xo = np.zeros_like(Moo)
    for s in range(Vhc):
        sl = slice(s * N, (s + 1) * N)
        Xe[sl, sl] = Mee[sl, sl]; Xo[sl, sl] = Moo[sl, sl]
    assert np.abs(Mee - Xe).max() < 1e-6 * np.abs(Mee).max()  # same-parity couplings only within a site
    Mhat = np.eye(ne) - np.linalg.inv(Xe) @ Meo @ np.linalg.inv(Xo) @ Moe
    vin = rng.standard_normal(2 * ne).astype(np.float32)
    vout = np.zeros_like(vin)
    L.mgMatQudaB200(mg, 1, 1, vp(vout), vp(vin))
    assert rel_l2(as_c(vout.astype(np.float64)), Mhat @ as_c(vin.astype(np.float64))) < 1e-5
    L.destroyMultigridQuda(mg)


def host_residual(oracle, g, x, b, kappa, mu):
    """|b - M x| / |b| with the reference CPU operator (multigrid_invert_test.cpp:529-577)."""
    return np.linalg.norm(b - oracle.tm_mat(g, x, kappa, mu, 1, 0)) / np.linalg.norm(b)


def point_source(V):
    b = np.zeros(V * 24)
    b[0] = 1.0  # first real component, as the reference test's source (multigrid_invert_test.cpp:497-508 sets 12 reals)
    b[2] = 1.0
    return b


@pytest.mark.parametrize("inv,precond", [("gcr", None), ("bicgstab", None), ("gcr", "mr"), ("mr", None)])
def test_krylov_solvers_without_multigrid(quda, oracle, inv, precond):
    """invert_test-style: GCR / BiCGStab / MR, even-odd preconditioned and full, true residual checked on the host."""
    q, L = quda, quda.lib()
    X = (8, 8, 8, 8)
    kappa, mu = 0.12, 0.1
    oracle.set_dims(X)
    g = oracle.weak_gauge(eps=0.3, antiperiodic=True, seed=7)
    load_gauge(q, g, X, antiperiodic=True)
    b = oracle.drand(oracle.V * 24, seed=5)
    for solve_type in (q.QUDA_DIRECT_PC_SOLVE, q.QUDA_DIRECT_SOLVE):
        p = mg_inv_param(q, kappa, mu)
        p.solve_type = solve_type
        p.inv_type = {"gcr": q.QUDA_GCR_INVERTER, "bicgstab": q.QUDA_BICGSTAB_INVERTER, "mr": q.QUDA_MR_INVERTER}[inv]
        p.inv_type_precondition = q.QUDA_MR_INVERTER if precond == "mr" else q.QUDA_INVALID_INVERTER
        p.tol = 1e-9 if inv != "mr" else 1e-3
        p.maxiter = 2000 if inv != "mr" else 300
        p.gcrNkrylov = 16
        p.reliable_delta = 1e-4
        p.maxiter_precondition = 4
        p.tol_precondition = 0.1
        p.omega = 1.0
        x = np.zeros_like(b)
        L.invertQuda(vp(x), vp(b), C.byref(p))
        res = host_residual(oracle, g, x, b, kappa, mu)
        if inv == "mr":
            assert res < 0.05 and p.iter == 300
        else:
            assert res < 5e-9, (inv, precond, solve_type, res, p.true_res)
            assert abs(p.true_res - res) < 0.5 * res + 1e-10
        assert p.iter > 0 and p.secs > 0


@pytest.mark.parametrize("solve", ["normop_pc", "normop"])
def test_cg_on_the_normal_equations(quda, oracle, solve):
    """invert_test's default solver (QUDA_CG_INVERTER, lib/inv_cg_quda.cpp) on M^dag M, fp64 outer / fp32 sloppy with reliable updates:
    x solves M x = b (MAT solution through the normal equations); residual of M checked on the host."""
    q, L = quda, quda.lib()
    X = (8, 8, 8, 8)
    kappa, mu = 0.12, 0.1
    oracle.set_dims(X)
    g = oracle.weak_gauge(eps=0.3, antiperiodic=True, seed=7)
    load_gauge(q, g, X, antiperiodic=True)
    b = oracle.drand(oracle.V * 24, seed=5)
    p = mg_inv_param(q, kappa, mu)
    p.solve_type = q.QUDA_NORMOP_PC_SOLVE if solve == "normop_pc" else q.QUDA_NORMOP_SOLVE
    p.inv_type = q.QUDA_CG_INVERTER
    p.tol = 1e-10; p.maxiter = 4000; p.reliable_delta = 0.1
    x = np.zeros_like(b)
    L.invertQuda(vp(x), vp(b), C.byref(p))
    res = host_residual(oracle, g, x, b, kappa, mu)
    print(f"CG {solve}: {p.iter} iterations, host residual of M x = b: {res:.2e}, reported (normal system) {p.true_res:.2e}")
    assert res < 5e-8 and p.true_res < 2e-10 and p.iter > 0


def run_mg_solve(q, oracle, X, blocks, nvecs, n_level, kappa, mu, eps, tol, precond_prec=4, nu=2, setup_maxiter=200):
    L = q.lib()
    oracle.set_dims(X)
    g = oracle.weak_gauge(eps=eps, antiperiodic=False, seed=4711)
    load_gauge(q, g, X, prec=8, sloppy=4, precond=precond_prec, recon=12)
    ip = mg_inv_param(q, kappa, mu, precond=precond_prec)
    mgp = q.multigrid_param(ip, n_level=n_level, geo_block=blocks, n_vec=nvecs, nu_pre=nu, nu_post=nu, setup_maxiter=setup_maxiter, setup_tol=5e-6)
    mg = L.newMultigridQuda(C.byref(mgp))
    b = point_source(oracle.V)
    # outer solve: GCR preconditioned by MG (setInvertParam of the reference test)
    p = mg_inv_param(q, kappa, mu, precond=precond_prec)
    p.inv_type_precondition = q.QUDA_MG_INVERTER
    p.preconditioner = mg
    p.gcrNkrylov = 20
    p.tol = tol
    p.maxiter = 200
    p.reliable_delta = 1e-4
    x = np.zeros_like(b)
    L.invertQuda(vp(x), vp(b), C.byref(p))
    res = host_residual(oracle, g, x, b, kappa, mu)
    # same solve without the preconditioner
    p0 = mg_inv_param(q, kappa, mu, precond=precond_prec)
    p0.gcrNkrylov = 20; p0.tol = tol; p0.maxiter = 4000; p0.reliable_delta = 1e-4
    x0 = np.zeros_like(b)
    L.invertQuda(vp(x0), vp(b), C.byref(p0))
    devs = []
    for l in range(n_level - 1):
        dev = (C.c_double * 3)()
        L.mgVerifyQudaB200(mg, l, dev)
        devs.append(list(dev))
    L.destroyMultigridQuda(mg)
    return res, p.true_res, p.iter, p0.iter, host_residual(oracle, g, x0, b, kappa, mu), devs, mgp.secs, p.secs, p0.secs


def test_two_level_mg_gcr_solve(quda, oracle):
    """BASELINE config 4 in miniature: 2-level MG (4^4 aggregates), twisted mass, MR smoother, K-cycle."""
    res, true_res, it_mg, it_plain, res_plain, devs, t_setup, t_mg, t_plain = run_mg_solve(
        quda, oracle, (8, 8, 8, 16), ((4, 4, 4, 4),), (8,), 2, kappa=0.1245, mu=0.005, eps=0.25, tol=1e-8)
    print(f"2-level: MG iters {it_mg} vs plain GCR {it_plain}; host res {res:.2e} (plain {res_plain:.2e}); setup {t_setup:.2f}s solve {t_mg:.3f}s plain {t_plain:.3f}s")
    assert res < 5e-8 and abs(true_res - res) < 0.5 * res + 1e-10
    assert res_plain < 5e-8
    assert it_mg < it_plain / 3, (it_mg, it_plain)
    for d in devs:
        assert d[0] < 5e-6 and d[1] < 1e-4 and d[2] < 5e-5, d


def test_three_level_mg_gcr_solve(quda, oracle):
    """BASELINE config 5 in miniature: 3-level K-cycle, 4^4 then 2^4 aggregates, same true residual as the host check."""
    res, true_res, it_mg, it_plain, res_plain, devs, t_setup, t_mg, t_plain = run_mg_solve(
        quda, oracle, (16, 16, 16, 16), ((4, 4, 4, 4), (2, 2, 2, 2)), (8, 8), 3, kappa=0.1248, mu=0.004, eps=0.25, tol=1e-8)
    print(f"3-level: MG iters {it_mg} vs plain GCR {it_plain}; host res {res:.2e}; setup {t_setup:.2f}s solve {t_mg:.3f}s plain {t_plain:.3f}s")
    assert res < 5e-8 and abs(true_res - res) < 0.5 * res + 1e-10
    assert it_mg < it_plain / 3, (it_mg, it_plain)
    for d in devs:
        assert d[0] < 5e-6 and d[1] < 1e-4 and d[2] < 5e-5, d


@pytest.mark.parametrize("n_level,X,blocks,nvecs", [(2, (8, 8, 8, 16), ((4, 4, 4, 4),), (8,)),
                                                    (3, (16, 16, 16, 16), ((4, 4, 4, 4), (2, 2, 2, 2)), (8, 8))])
@pytest.mark.parametrize("mode", [3, 1])
def test_block_mg_multi_src_solve(quda, oracle, n_level, X, blocks, nvecs, mode, monkeypatch):
    """invertMultiSrcQuda on the block path: all sources through the K-cycle in lock-step, coarse levels on
    the multi-RHS tensor-core operator.  Every solution is checked with the host operator of the oracle and against the
    one-source-at-a-time path (same tolerance, comparable iteration count)."""
    q, L = quda, quda.lib()
    kappa, mu, tol = 0.1248, 0.004, 1e-8
    oracle.set_dims(X)
    g = oracle.weak_gauge(eps=0.25, antiperiodic=False, seed=4711)
    load_gauge(q, g, X, prec=8, sloppy=4, precond=4, recon=12)
    ip = mg_inv_param(q, kappa, mu)
    mgp = q.multigrid_param(ip, n_level=n_level, geo_block=blocks, n_vec=nvecs, nu_pre=2, nu_post=2, setup_maxiter=200, setup_tol=5e-6)
    mg = L.newMultigridQuda(C.byref(mgp))
    nsrc = 5
    rng = np.random.default_rng(11)
    bs = [point_source(oracle.V)] + [rng.standard_normal(oracle.V * 24) for _ in range(nsrc - 1)]

    def solve(block):
        monkeypatch.setenv("QB_BLOCK_MG", "1" if block else "0")
        monkeypatch.setenv("QB_BLOCK_MG_MODE", str(mode))
        p = mg_inv_param(q, kappa, mu)
        p.inv_type_precondition = q.QUDA_MG_INVERTER
        p.preconditioner = mg
        p.gcrNkrylov = 20; p.tol = tol; p.maxiter = 200; p.reliable_delta = 1e-4
        p.num_src = nsrc
        xs = [np.zeros(oracle.V * 24) for _ in range(nsrc)]
        L.invertMultiSrcQuda((C.c_void_p * nsrc)(*[a.ctypes.data for a in xs]), (C.c_void_p * nsrc)(*[a.ctypes.data for a in bs]), C.byref(p))
        return xs, p.iter, p.true_res, p.secs

    xb, it_b, tr_b, t_b = solve(True)
    xs, it_s, tr_s, t_s = solve(False)
    if mode == 3 and n_level == 2:
        # blocks of 2 sources: 5 sources = 2 + 2 + a left-over single one (ordinary path), counters accumulate
        monkeypatch.setenv("QB_BLOCK_MG_R", "2")
        xc, it_c, tr_c, _ = solve(True)
        monkeypatch.delenv("QB_BLOCK_MG_R")
        assert tr_c < 5e-8 and it_c >= it_b
        for a, b_ in zip(xc, xs):
            assert np.linalg.norm(a - b_) / np.linalg.norm(b_) < 1e-6
    L.destroyMultigridQuda(mg)
    worst = max(host_residual(oracle, g, x, b, kappa, mu) for x, b in zip(xb, bs))
    print(f"block MG ({n_level} levels, mode {mode}): {it_b} lock-step iterations in {t_b:.3f} s, sequential {it_s} iterations (sum over {nsrc}) in {t_s:.3f} s; "
          f"worst host residual {worst:.2e}, reported {tr_b:.2e}")
    assert worst < 5e-8 and tr_b < 5e-8 and tr_s < 5e-8
    assert it_b <= 1.5 * it_s / nsrc + 3, (it_b, it_s)    # lock-step count = that of the slowest source
    for a, b_ in zip(xb, xs):
        assert np.linalg.norm(a - b_) / np.linalg.norm(b_) < 1e-6


def test_null_vector_files_round_trip(quda, oracle, tmp_path):
    """vec_outfile / vec_infile (multigrid.cpp:607-691; SURVEY 8f.2: QIO-free container): a hierarchy rebuilt from the saved
    near-null vectors (compute_null_vector = NO) has the same vectors, skips the setup solves and solves in the same iterations."""
    q, L = quda, quda.lib()
    X, kappa, mu = (8, 8, 8, 16), 0.1245, 0.005
    oracle.set_dims(X)
    g = oracle.weak_gauge(eps=0.25, antiperiodic=False, seed=4711)
    load_gauge(q, g, X, prec=8, sloppy=4, precond=4, recon=12)
    base = str(tmp_path / "nullvec").encode()

    def build(compute):
        ip = mg_inv_param(q, kappa, mu)
        mgp = q.multigrid_param(ip, n_level=3, geo_block=((2, 2, 2, 4), (2, 2, 2, 2)), n_vec=(8, 8), setup_maxiter=100, setup_tol=5e-6)
        if compute:
            mgp.vec_outfile = base
        else:
            mgp.compute_null_vector = q.QUDA_COMPUTE_NULL_VECTOR_NO
            mgp.vec_infile = base
        mg = L.newMultigridQuda(C.byref(mgp))
        vecs = []
        info = (C.c_int * 8)()
        L.mgLevelInfoQudaB200(mg, 0, info)   # geometry of level 1 (the coarse side of the level-0 transfer)
        for lvl, n in ((0, oracle.V * 24), (1, int(np.prod(info[0:4])) * info[7] * 2)):
            v = np.zeros(n, dtype=np.float32)
            L.mgNullVectorQudaB200(mg, lvl, 3, v.ctypes.data)
            vecs.append(v)
        b = point_source(oracle.V); x = np.zeros_like(b)
        p = mg_inv_param(q, kappa, mu); p.inv_type_precondition = q.QUDA_MG_INVERTER; p.preconditioner = mg
        p.gcrNkrylov = 20; p.tol = 1e-8; p.maxiter = 200; p.reliable_delta = 1e-4
        L.invertQuda(vp(x), vp(b), C.byref(p))
        res = host_residual(oracle, g, x, b, kappa, mu)
        L.destroyMultigridQuda(mg)
        return vecs, p.iter, res, mgp.secs

    v1, it1, res1, t1 = build(True)
    import os
    assert os.path.exists(base.decode() + "_level_0") and os.path.exists(base.decode() + "_level_1")
    # documented layout: 128-byte header, then global lexicographic sites, [spin][colour][re,im] fp32
    raw = open(base.decode() + "_level_0", "rb").read()
    hdr = np.frombuffer(raw[8:8 + 10 * 4], dtype="<i4")
    assert raw[:8] == b"QB200VEC" and list(hdr) == [2, 0, 8, 4, 3, 8, 8, 8, 16, 4]
    vec3 = np.frombuffer(raw, dtype="<f4", offset=128 + 3 * oracle.V * 24 * 4, count=oracle.V * 24).reshape(oracle.V, 24)
    lex = np.arange(oracle.V)
    xs, ys, zs, ts = lex % 8, (lex // 8) % 8, (lex // 64) % 8, lex // 512
    eo = ((xs + ys + zs + ts) & 1) * oracle.Vh + (lex >> 1)
    assert np.array_equal(vec3, v1[0].reshape(oracle.V, 24)[eo])
    v2, it2, res2, t2 = build(False)
    print(f"null-vector files: setup {t1:.2f} s computing, {t2:.2f} s loading; MG-GCR {it1} / {it2} iterations, residuals {res1:.2e} / {res2:.2e}")
    for a, b_ in zip(v1, v2):
        assert np.array_equal(a, b_)
    assert it1 == it2 and res2 < 5e-8


def test_mg_fp16_preconditioner_storage(quda, oracle, monkeypatch):
    """QB_MG_HALF_STORAGE=1: V (prolongator / restrictor) and the coarse links of the single-RHS kernel stored as fp16, fp32 arithmetic.
    Only the preconditioner changes: the outer solve reaches the same true residual in (almost) the same number of iterations, and the
    transfer operators agree with their fp32 versions to fp16 rounding."""
    q, L = quda, quda.lib()
    X, kappa, mu, tol = (16, 16, 16, 16), 0.1248, 0.004, 1e-8
    oracle.set_dims(X)
    g = oracle.weak_gauge(eps=0.25, antiperiodic=False, seed=4711)
    load_gauge(q, g, X, prec=8, sloppy=4, precond=4, recon=12)
    rng = np.random.default_rng(3)
    out = {}
    for half in (0, 1):
        monkeypatch.setenv("QB_MG_HALF_STORAGE", str(half))
        ip = mg_inv_param(q, kappa, mu)
        mgp = q.multigrid_param(ip, n_level=3, geo_block=((4, 4, 4, 4), (2, 2, 2, 2)), n_vec=(8, 8), setup_maxiter=200, setup_tol=5e-6)
        mg = L.newMultigridQuda(C.byref(mgp))
        info = (C.c_int * 8)()
        L.mgLevelInfoQudaB200(mg, 0, info)
        nc = int(np.prod(info[0:4])) * info[7] * 2
        rng = np.random.default_rng(3)
        fine = rng.standard_normal(oracle.V * 24).astype(np.float32)
        coarse = rng.standard_normal(nc).astype(np.float32)
        pf = np.zeros(oracle.V * 24, dtype=np.float32); rc = np.zeros(nc, dtype=np.float32)
        fp = C.POINTER(C.c_float)
        L.mgProlongQudaB200(mg, 0, pf.ctypes.data_as(fp), coarse.ctypes.data_as(fp))
        L.mgRestrictQudaB200(mg, 0, rc.ctypes.data_as(fp), fine.ctypes.data_as(fp))
        b = point_source(oracle.V); x = np.zeros_like(b)
        p = mg_inv_param(q, kappa, mu); p.inv_type_precondition = q.QUDA_MG_INVERTER; p.preconditioner = mg
        p.gcrNkrylov = 20; p.tol = tol; p.maxiter = 200; p.reliable_delta = 1e-4
        L.invertQuda(vp(x), vp(b), C.byref(p))
        out[half] = (pf, rc, p.iter, host_residual(oracle, g, x, b, kappa, mu))
        L.destroyMultigridQuda(mg)
    dp = np.linalg.norm(out[1][0] - out[0][0]) / np.linalg.norm(out[0][0])
    dr = np.linalg.norm(out[1][1] - out[0][1]) / np.linalg.norm(out[0][1])
    print(f"fp16 storage: P deviates by {dp:.2e}, R by {dr:.2e}; MG-GCR iterations {out[0][2]} (fp32) / {out[1][2]} (fp16), residuals {out[0][3]:.2e} / {out[1][3]:.2e}")
    assert 1e-6 < dp < 2e-3 and 1e-6 < dr < 2e-3            # fp16 rounding of V, and really a different storage
    assert out[1][3] < 5e-8 and abs(out[1][2] - out[0][2]) <= 2


def test_mg_half_precision_smoother(quda, oracle):
    """cuda_prec_precondition = half on the fine level (int16 links and smoother mat-vec), as the reference allows."""
    res, true_res, it_mg, it_plain, *_ = run_mg_solve(
        quda, oracle, (8, 8, 8, 16), ((4, 4, 4, 4),), (8,), 2, kappa=0.1245, mu=0.005, eps=0.25, tol=1e-8, precond_prec=2)
    assert res < 5e-8 and it_mg < it_plain / 3


@pytest.mark.parametrize("mask", [8, 12, 15])
def test_mg_on_partitioned_lattice_self_exchange(mask):
    """Coarse-level halo path on one GPU (the reference's --partition trick): V ghost slices and ghost links in the
    coarse-link build, coarse-spinor halos in the coarse Dslash.  Checks the Galerkin identity R M P = M_c on both
    levels (fine M with halos is verified against the oracle elsewhere) and a 3-level MG-GCR solve with the host residual."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = f"""
import sys, ctypes as C, numpy as np
sys.path.insert(0, {root!r})
import quda_b200 as q
from tests import oracle_util as ou
from tests.test_multigrid_gpu import load_gauge, mg_inv_param, host_residual, point_source, vp
o = ou.load_oracle(); X=(8,8,8,16); o.set_dims(X)
kappa, mu = 0.1245, 0.005
g = o.weak_gauge(eps=0.25, antiperiodic=True, seed=4711)
L = q.lib(); L.initQuda(0); L.commDimPartitionedSetQudaB200({mask})
load_gauge(q, g, X, prec=8, sloppy=4, precond=4, recon=12, antiperiodic=True)
ip = mg_inv_param(q, kappa, mu)
mgp = q.multigrid_param(ip, n_level=3, geo_block=((2,2,2,4),(2,2,2,2)), n_vec=(8,8), setup_maxiter=100, setup_tol=5e-6)
mg = L.newMultigridQuda(C.byref(mgp))
worst = 0.0
for lvl in (0, 1):
    dev = (C.c_double*3)(); L.mgVerifyQudaB200(mg, lvl, dev)
    print("DEV", lvl, list(dev))
    assert dev[0] < 5e-6 and dev[1] < 1e-4 and dev[2] < 5e-5, list(dev)
b = point_source(o.V); x = np.zeros_like(b)
p = mg_inv_param(q, kappa, mu); p.inv_type_precondition = q.QUDA_MG_INVERTER; p.preconditioner = mg
p.gcrNkrylov = 20; p.tol = 1e-8; p.maxiter = 200; p.reliable_delta = 1e-4
L.invertQuda(vp(x), vp(b), C.byref(p))
res = host_residual(o, g, x, b, kappa, mu)
L.destroyMultigridQuda(mg); L.endQuda()
print("RESULT", res, p.iter)
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    res, it = r.stdout.strip().split("RESULT")[-1].split()
    assert float(res) < 5e-8 and int(it) < 60, r.stdout[-1000:]


@pytest.mark.parametrize("X,bs", [((16, 8, 8, 8), (4, 4, 4, 4)), ((32, 4, 4, 8), (8, 2, 2, 4)), ((16, 4, 4, 4), (2, 2, 2, 2))])
def test_restrictor_is_adjoint_of_prolongator_row_major_kernel(quda, oracle, X, bs):
    """Lattices with X/2 a multiple of 8 take the row-major restrictor (restrict_rows_kernel): <R f, c> = <f, P c> for random
    f, c (R = P^dag, lib/restrictor.cu:90-125 vs lib/prolongator.cu:41-56), and R agrees with the aggregate-major kernel."""
    import os
    q, L = quda, quda.lib()
    oracle.set_dims(X)
    g = oracle.weak_gauge(eps=0.3, antiperiodic=False, seed=3)
    load_gauge(q, g, X)
    ip = mg_inv_param(q, 0.124, 0.02)
    nvec = 8
    mgp = q.multigrid_param(ip, n_level=2, geo_block=(bs,), n_vec=(nvec,), setup_maxiter=5, setup_tol=1e-1, run_verify=False)
    mg = L.newMultigridQuda(C.byref(mgp))
    info = (C.c_int * 8)()
    L.mgLevelInfoQudaB200(mg, 0, info)
    Vf, Vc, N = int(np.prod(X)), int(np.prod(info[0:4])), info[7]
    rng = np.random.default_rng(2)
    f = rng.standard_normal(2 * Vf * 12).astype(np.float32)
    c = rng.standard_normal(2 * Vc * N).astype(np.float32)
    Rf = np.zeros_like(c); Pc = np.zeros_like(f)
    L.mgRestrictQudaB200(mg, 0, vp(Rf), vp(f))
    L.mgProlongQudaB200(mg, 0, vp(Pc), vp(c))
    lhs = np.vdot(as_c(Rf.astype(np.float64)), as_c(c.astype(np.float64)))
    rhs = np.vdot(as_c(f.astype(np.float64)), as_c(Pc.astype(np.float64)))
    assert abs(lhs - rhs) <= 2e-6 * np.linalg.norm(f) * np.linalg.norm(Pc)
    os.environ["QB_RESTRICT_OLD"] = "1"
    Rf_old = np.zeros_like(c)
    L.mgRestrictQudaB200(mg, 0, vp(Rf_old), vp(f))
    os.environ.pop("QB_RESTRICT_OLD")
    assert rel_l2(Rf.astype(np.float64), Rf_old.astype(np.float64)) < 1e-6
    L.destroyMultigridQuda(mg)
