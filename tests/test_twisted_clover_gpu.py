"""GPU parity tests of the twisted-clover (and Wilson-clover) operator through the C ABI: loadCloverQuda + dslashQuda / MatQuda /
invertQuda with dslash_type = QUDA_TWISTED_CLOVER_DSLASH, modelled on tests/dslash_test.cpp (--dslash_type twisted-clover,
test types 0-4).  Oracle: oracle/tm_oracle_impl.h restatement of tests/clover_reference.cpp (tmc_dslash / tmc_mat / tmc_matpc),
pinned bit for bit to the reference's own objects by tests/test_oracle.py.  Tolerances: fp64 1e-13, fp32 1e-6, half 1e-3
relative L2 (north_star)."""
import ctypes as C

import numpy as np
import pytest

from tests.oracle_util import rel_l2

pytestmark = pytest.mark.gpu

KAPPA, MU = 0.1, 0.05
TOL = {8: 1e-13, 4: 1e-6, 2: 1e-3}


def vp(a):
    return a.ctypes.data_as(C.c_void_p)


class Ctx:
    def __init__(self, q, o, X, prec, recon=18, dslash_type=None):
        self.q, self.o, self.X, self.prec = q, o, X, prec
        self.dslash_type = q.QUDA_TWISTED_CLOVER_DSLASH if dslash_type is None else dslash_type
        o.set_dims(X)
        self.g = o.gauge(kind=1, antiperiodic=True, seed=137)
        self.sp = o.drand(2 * o.Vh * 24, seed=137) - 0.5
        self.Vh = o.Vh
        self.even = self.sp[: self.Vh * 24].copy()
        self.c = o.clover(norm=0.1, diag=1.0, seed=4242)
        self.mu = MU if self.dslash_type == q.QUDA_TWISTED_CLOVER_DSLASH else 0.0
        self.cinv = o.clover_inverse(self.c, KAPPA, self.mu)
        L = q.lib()
        gp = q.gauge_param(X, cuda_prec=prec, reconstruct=recon)
        L.loadGaugeQuda((C.c_void_p * 4)(*[a.ctypes.data for a in self.g]), C.byref(gp))
        p = self.param()
        got_inv = np.zeros_like(self.c)
        p.return_clover_inverse = 1 if self.dslash_type == q.QUDA_TWISTED_CLOVER_DSLASH else 0
        L.loadCloverQuda(vp(self.c), vp(got_inv), C.byref(p))
        self.returned_inverse = got_inv

    def param(self, **kw):
        p = self.q.invert_param(kappa=KAPPA, mu=self.mu, cuda_prec=self.prec, dslash_type=self.dslash_type, **kw)
        p.clover_cpu_prec = 8
        p.clover_cuda_prec = p.clover_cuda_prec_sloppy = p.clover_cuda_prec_precondition = self.prec
        p.clover_order = self.q.QUDA_PACKED_CLOVER_ORDER
        p.clover_coeff = 1.0
        p.compute_clover = p.compute_clover_inverse = p.return_clover = p.return_clover_inverse = 0
        return p


@pytest.mark.parametrize("prec", [8, 4, 2])
def test_tmc_dslash_all_variants(quda, oracle, prec):
    """dslash_test --test 0: the preconditioned hop A^-1 D (or D A^-1 daggered) with A = C + i a gamma5."""
    c = Ctx(quda, oracle, (8, 8, 8, 8), prec)
    L = quda.lib()
    if prec == 8:
        # the "inverse" field handed back by loadCloverQuda is (C^2 + (2 kappa mu)^2)^-1 (lib/clover_invert.cu:56-90)
        assert rel_l2(c.returned_inverse, c.cinv) < 1e-13
    worst = 0.0
    for flavor in (1, -1):
        for parity in (0, 1):
            for matpc in (0, 2):
                for dagger in (0, 1):
                    p = c.param(flavor=flavor, matpc=matpc, dagger=dagger)
                    out = np.zeros(c.Vh * 24)
                    L.dslashQuda(vp(out), vp(c.even), C.byref(p), parity)
                    oracle.set_dims(c.X)
                    if dagger and matpc == 2:
                        # asymmetric + dagger: the device operator is A^-dag D^dag (dirac_twisted_clover.cpp:207-214); the host
                        # tmc_dslash additionally twists its input, so compose the expected result from the pinned primitives
                        t = oracle.wil_dslash(c.g, c.even, parity, 1)
                        ref = np.zeros_like(t)
                        oracle.L.orc_twist_clover_gamma5_d(vp(ref), vp(t), vp(c.c), vp(c.cinv), 1, KAPPA, MU, flavor, parity, 1)
                    else:
                        ref = oracle.tmc_dslash(c.g, c.even, c.c, c.cinv, KAPPA, MU, flavor, parity, matpc, dagger)
                    worst = max(worst, rel_l2(out, ref))
    print(f"twisted-clover dslash prec {prec}: worst rel-L2 {worst:.2e}")
    assert worst <= TOL[prec]


@pytest.mark.parametrize("prec", [8, 4])
def test_tmc_matpc_and_mat(quda, oracle, prec):
    """dslash_test --test 1 / 2 (+ dagger, all matpc types, both flavours)."""
    c = Ctx(quda, oracle, (8, 8, 8, 8), prec, recon=12)
    L = quda.lib()
    worst = 0.0
    for flavor in (1, -1):
        for matpc in range(4):
            for dagger in (0, 1):
                p = c.param(flavor=flavor, matpc=matpc, dagger=dagger, solution_type=quda.QUDA_MATPC_SOLUTION)
                out = np.zeros(c.Vh * 24)
                L.MatQuda(vp(out), vp(c.even), C.byref(p))
                ref = oracle.tmc_matpc(c.g, c.even, c.c, c.cinv, KAPPA, MU, flavor, matpc, dagger)
                worst = max(worst, rel_l2(out, ref))
        for dagger in (0, 1):
            p = c.param(flavor=flavor, dagger=dagger, solution_type=quda.QUDA_MAT_SOLUTION)
            out = np.zeros(2 * c.Vh * 24)
            L.MatQuda(vp(out), vp(c.sp), C.byref(p))
            ref = oracle.tmc_mat(c.g, c.sp, c.c, KAPPA, MU, flavor, dagger)
            worst = max(worst, rel_l2(out, ref))
    print(f"twisted-clover matpc / mat prec {prec}: worst rel-L2 {worst:.2e}")
    assert worst <= TOL[prec]


def test_wilson_clover_is_the_zero_twist_limit(quda, oracle):
    """dslash_type = QUDA_CLOVER_WILSON_DSLASH: same code with a = 0 (clover_reference.cpp clover_matpc / clover_mat)."""
    c = Ctx(quda, oracle, (4, 6, 4, 8), 8, dslash_type=quda.QUDA_CLOVER_WILSON_DSLASH)
    L = quda.lib()
    for matpc in range(4):
        p = c.param(matpc=matpc, solution_type=quda.QUDA_MATPC_SOLUTION)
        out = np.zeros(c.Vh * 24)
        L.MatQuda(vp(out), vp(c.even), C.byref(p))
        ref = oracle.tmc_matpc(c.g, c.even, c.c, c.cinv, KAPPA, 0.0, 1, matpc, 0)
        assert rel_l2(out, ref) <= 1e-13
    p = c.param(solution_type=quda.QUDA_MAT_SOLUTION)
    out = np.zeros(2 * c.Vh * 24)
    L.MatQuda(vp(out), vp(c.sp), C.byref(p))
    assert rel_l2(out, oracle.tmc_mat(c.g, c.sp, c.c, KAPPA, 0.0, 1, 0)) <= 1e-13


@pytest.mark.parametrize("solve_type,matpc", [("direct_pc", 0), ("direct_pc", 3), ("direct", 0)])
def test_tmc_solve_reaches_host_residual(quda, oracle, solve_type, matpc):
    """invert_test with --dslash_type twisted-clover: GCR solve of M x = b, residual checked with the host operator."""
    q, L = quda, quda.lib()
    c = Ctx(q, oracle, (8, 8, 8, 8), 8, recon=12)
    p = c.param(flavor=1, matpc=matpc, solution_type=q.QUDA_MAT_SOLUTION)
    p.solve_type = q.QUDA_DIRECT_PC_SOLVE if solve_type == "direct_pc" else q.QUDA_DIRECT_SOLVE
    p.inv_type = q.QUDA_GCR_INVERTER
    p.gcrNkrylov = 20
    p.tol = 1e-10
    p.maxiter = 2000
    p.reliable_delta = 1e-4
    b = c.sp.copy()
    x = np.zeros_like(b)
    L.invertQuda(vp(x), vp(b), C.byref(p))
    res = np.linalg.norm(b - oracle.tmc_mat(c.g, x, c.c, KAPPA, MU, 1, 0)) / np.linalg.norm(b)
    print(f"twisted-clover GCR ({solve_type}, matpc {matpc}): {p.iter} iterations, host residual {res:.2e}, reported {p.true_res:.2e}")
    assert res < 5e-10 and abs(p.true_res - res) < 0.5 * res + 1e-12


@pytest.mark.parametrize("nvec", [4, 8])
def test_tmc_multigrid_coarsening_and_solve(quda, oracle, nvec):
    """Multigrid on the twisted-clover operator (the coarse site-diagonal block is V^dag (C + i a gamma5) V:
    lib/coarse_op.cuh computeTMCAV / COMPUTE_COARSE_CLOVER): the Galerkin identity R M P = M_c checked inside the library
    (n_vec 4: CUDA-core build, n_vec 8: tensor-core build) and an MG-preconditioned GCR solve checked with the host operator."""
    q, L = quda, quda.lib()
    X = (8, 8, 8, 8)
    oracle.set_dims(X)
    g = oracle.weak_gauge(eps=0.25, antiperiodic=False, seed=21)
    cl = oracle.clover(norm=0.05, diag=1.0, seed=99)
    kappa, mu = 0.122, 0.01
    gp = q.gauge_param(X, cuda_prec=8, reconstruct=18, cuda_prec_sloppy=4, cuda_prec_precondition=4, t_boundary=q.QUDA_PERIODIC_T)
    L.loadGaugeQuda((C.c_void_p * 4)(*[a.ctypes.data for a in g]), C.byref(gp))

    def param():
        p = q.invert_param(kappa=kappa, mu=mu, cuda_prec=8, dslash_type=q.QUDA_TWISTED_CLOVER_DSLASH, solution_type=q.QUDA_MAT_SOLUTION)
        p.cuda_prec_sloppy = 4; p.cuda_prec_precondition = 4
        p.clover_cpu_prec = 8
        p.clover_cuda_prec = 8; p.clover_cuda_prec_sloppy = 4; p.clover_cuda_prec_precondition = 4
        p.clover_order = q.QUDA_PACKED_CLOVER_ORDER
        p.clover_coeff = 1.0
        p.compute_clover = p.compute_clover_inverse = p.return_clover = p.return_clover_inverse = 0
        p.solve_type = q.QUDA_DIRECT_SOLVE
        p.inv_type = q.QUDA_GCR_INVERTER
        p.gcrNkrylov = 20; p.tol = 1e-9; p.maxiter = 2000; p.reliable_delta = 1e-4
        return p

    ip = param()
    L.loadCloverQuda(vp(cl), None, C.byref(ip))
    mgp = q.multigrid_param(ip, n_level=2, geo_block=((4, 4, 4, 4),), n_vec=(nvec,), setup_maxiter=200, setup_tol=5e-6, run_verify=False)
    mg = L.newMultigridQuda(C.byref(mgp))
    dev = (C.c_double * 3)()
    L.mgVerifyQudaB200(mg, 0, dev)
    print(f"twisted-clover MG n_vec={nvec}: verify deviations {list(dev)}")
    assert dev[0] < 5e-6 and dev[2] < 5e-5, list(dev)
    b = np.zeros(oracle.V * 24); b[0] = 1.0; b[2] = 1.0
    x = np.zeros_like(b)
    p = param()
    p.inv_type_precondition = q.QUDA_MG_INVERTER
    p.preconditioner = mg
    L.invertQuda(vp(x), vp(b), C.byref(p))
    res = np.linalg.norm(b - oracle.tmc_mat(g, x, cl, kappa, mu, 1, 0)) / np.linalg.norm(b)
    x0 = np.zeros_like(b)
    p0 = param()
    L.invertQuda(vp(x0), vp(b), C.byref(p0))
    print(f"twisted-clover MG-GCR: {p.iter} iterations (plain GCR {p0.iter}), host residual {res:.2e}")
    assert res < 5e-9 and p.iter < p0.iter / 2
    L.destroyMultigridQuda(mg)


def test_invert_multi_src(quda, oracle):
    """invertMultiSrcQuda: several sources through one resident operator; every solution checked with the host operator."""
    q, L = quda, quda.lib()
    c = Ctx(q, oracle, (8, 8, 8, 8), 8, recon=12)
    p = c.param(flavor=1, solution_type=q.QUDA_MAT_SOLUTION)
    p.solve_type = q.QUDA_DIRECT_PC_SOLVE
    p.inv_type = q.QUDA_GCR_INVERTER
    p.gcrNkrylov = 20; p.tol = 1e-9; p.maxiter = 2000; p.reliable_delta = 1e-4
    nsrc = 3
    p.num_src = nsrc
    rng = np.random.default_rng(8)
    bs = [rng.standard_normal(oracle.V * 24) for _ in range(nsrc)]
    xs = [np.zeros(oracle.V * 24) for _ in range(nsrc)]
    L.invertMultiSrcQuda((C.c_void_p * nsrc)(*[a.ctypes.data for a in xs]), (C.c_void_p * nsrc)(*[a.ctypes.data for a in bs]), C.byref(p))
    for x, b in zip(xs, bs):
        res = np.linalg.norm(b - oracle.tmc_mat(c.g, x, c.c, KAPPA, MU, 1, 0)) / np.linalg.norm(b)
        assert res < 5e-9
    assert p.iter > 0 and p.true_res < 5e-9


@pytest.mark.parametrize("nvec", [4, 8])
def test_tmc_multigrid_on_the_even_odd_system(quda, oracle, nvec):
    """The reference's default solve type with twisted clover (DiracTwistedCloverPC::createCoarseOp, computeTMCAV lib/coarse_op.cuh:384-456):
    coarse_grid_solution_type = MATPC coarsens A^-1 M with the site-dependent A = C + i a gamma5.  Dense check of the coarse operator against
    P^dag A^-1 (tmc_mat) P with the oracle (A^-1 from the oracle's clover blocks in numpy), the library's own identity, and an
    even-odd MG-GCR solve checked with the host operator."""
    from tests.test_multigrid_gpu import as_c
    q, L = quda, quda.lib()
    X = (4, 4, 4, 8)
    bs = (2, 2, 2, 2)
    oracle.set_dims(X)
    g = oracle.weak_gauge(eps=0.25, antiperiodic=False, seed=21)
    cl = oracle.clover(norm=0.05, diag=1.0, seed=99)
    kappa, mu = 0.122, 0.03
    a = 2 * kappa * mu
    gp = q.gauge_param(X, cuda_prec=8, reconstruct=18, cuda_prec_sloppy=4, cuda_prec_precondition=4, t_boundary=q.QUDA_PERIODIC_T)
    L.loadGaugeQuda((C.c_void_p * 4)(*[x_.ctypes.data for x_ in g]), C.byref(gp))

    def param():
        p = q.invert_param(kappa=kappa, mu=mu, cuda_prec=8, dslash_type=q.QUDA_TWISTED_CLOVER_DSLASH, solution_type=q.QUDA_MAT_SOLUTION)
        p.cuda_prec_sloppy = 4; p.cuda_prec_precondition = 4
        p.clover_cpu_prec = 8
        p.clover_cuda_prec = 8; p.clover_cuda_prec_sloppy = 4; p.clover_cuda_prec_precondition = 4
        p.clover_order = q.QUDA_PACKED_CLOVER_ORDER
        p.clover_coeff = 1.0
        p.compute_clover = p.compute_clover_inverse = p.return_clover = p.return_clover_inverse = 0
        p.solve_type = q.QUDA_DIRECT_SOLVE
        p.inv_type = q.QUDA_GCR_INVERTER
        p.gcrNkrylov = 20; p.tol = 1e-9; p.maxiter = 2000; p.reliable_delta = 1e-4
        return p

    ip = param()
    L.loadCloverQuda(vp(cl), None, C.byref(ip))
    mgp = q.multigrid_param(ip, n_level=2, geo_block=(bs,), n_vec=(nvec,), setup_maxiter=100, setup_tol=1e-4, solve_type=q.QUDA_DIRECT_PC_SOLVE)
    mg = L.newMultigridQuda(C.byref(mgp))
    dev = (C.c_double * 3)()
    L.mgVerifyQudaB200(mg, 0, dev)
    assert dev[0] < 5e-6 and dev[2] < 5e-5, list(dev)
    info = (C.c_int * 8)()
    L.mgLevelInfoQudaB200(mg, 0, info)
    N = info[7]
    Vf, Vc = int(np.prod(X)), int(np.prod(info[0:4]))
    nf, nc = Vf * 12, Vc * N
    P = np.zeros((nf, nc), dtype=np.complex128); Mc = np.zeros((nc, nc), dtype=np.complex128)
    e = np.zeros(2 * nc, dtype=np.float32); out = np.zeros(2 * nf, dtype=np.float32); oc = np.zeros(2 * nc, dtype=np.float32)
    for i in range(nc):
        e[:] = 0; e[2 * i] = 1
        L.mgProlongQudaB200(mg, 0, vp(out), vp(e)); P[:, i] = as_c(out.astype(np.float64))
        L.mgMatQudaB200(mg, 1, 0, vp(oc), vp(e)); Mc[:, i] = as_c(oc.astype(np.float64))
    # A^-1 site by site from the packed Hermitian blocks (two 6 x 6 blocks per site: chirality +, -)
    blocks = oracle.clover_unpack(cl).reshape(Vf, 2, 6, 6)
    Ainv = np.zeros_like(blocks)
    for chi, sgn in ((0, 1.0), (1, -1.0)):
        Ainv[:, chi] = np.linalg.inv(blocks[:, chi] + 1j * sgn * a * np.eye(6))

    def apply_ainv(z):
        v = z.reshape(Vf, 2, 6)
        return np.einsum("xcij,xcj->xci", Ainv, v).reshape(-1)

    def to_reals(z):
        r = np.zeros(2 * z.size); r[0::2] = z.real; r[1::2] = z.imag
        return r

    AMP = np.stack([apply_ainv(as_c(oracle.tmc_mat(g, to_reals(P[:, i]), cl, kappa, mu, 1, 0))) for i in range(nc)], axis=1)
    ref = P.conj().T @ AMP
    err = np.abs(Mc - ref).max() / np.abs(ref).max()
    print(f"twisted clover, preconditioned coarsening, n_vec {nvec}: |M_c - P^dag A^-1 M P| = {err:.2e}")
    assert err < 1e-5
    b = np.zeros(oracle.V * 24); b[0] = 1.0; b[2] = 1.0
    x = np.zeros_like(b)
    p = param()
    p.solve_type = q.QUDA_DIRECT_PC_SOLVE
    p.inv_type_precondition = q.QUDA_MG_INVERTER
    p.preconditioner = mg
    L.invertQuda(vp(x), vp(b), C.byref(p))
    res = np.linalg.norm(b - oracle.tmc_mat(g, x, cl, kappa, mu, 1, 0)) / np.linalg.norm(b)
    p0 = param(); p0.solve_type = q.QUDA_DIRECT_PC_SOLVE
    x0 = np.zeros_like(b)
    L.invertQuda(vp(x0), vp(b), C.byref(p0))
    print(f"twisted-clover even-odd MG-GCR: {p.iter} iterations (plain even-odd GCR {p0.iter}), host residual {res:.2e}")
    assert res < 5e-9 and p.iter < p0.iter
    L.destroyMultigridQuda(mg)
