"""invertQuda interface semantics that the reference defines in lib/interface_quda.cpp:2276-2543: initial guesses, resident
solutions (make_resident_solution, :2493-2508), the fields written back."""
import ctypes as C
import math

import numpy as np
import pytest

from tests.test_multigrid_gpu import host_residual, load_gauge, mg_inv_param, vp

pytestmark = pytest.mark.gpu


def setup(q, oracle, X=(8, 8, 8, 8), sloppy=8):
    oracle.set_dims(X)
    g = oracle.weak_gauge(eps=0.3, antiperiodic=True, seed=7)
    load_gauge(q, g, X, prec=8, sloppy=sloppy, precond=sloppy, antiperiodic=True)
    return g, oracle.drand(oracle.V * 24, seed=5)


@pytest.mark.parametrize("sloppy", [8, 4])
def test_gcr_restarted_from_a_converged_initial_guess_keeps_it(quda, oracle, sloppy):
    """use_init_guess = YES with a guess that already satisfies the tolerance: no iteration runs and the guess must come back
    unchanged (uniform precision used to return zero: the guess is moved to the accumulator and must be restored)"""
    q, L = quda, quda.lib()
    kappa, mu = 0.12, 0.1
    g, b = setup(q, oracle, sloppy=sloppy)
    p = mg_inv_param(q, kappa, mu, sloppy=sloppy, precond=sloppy)
    p.solve_type = q.QUDA_DIRECT_PC_SOLVE
    p.tol = 1e-9; p.maxiter = 2000; p.gcrNkrylov = 16; p.reliable_delta = 1e-4
    x = np.zeros_like(b)
    L.invertQuda(vp(x), vp(b), C.byref(p))
    assert host_residual(oracle, g, x, b, kappa, mu) < 5e-9 and p.iter > 0
    p2 = mg_inv_param(q, kappa, mu, sloppy=sloppy, precond=sloppy)
    p2.solve_type = q.QUDA_DIRECT_PC_SOLVE
    p2.tol = 1e-8; p2.maxiter = 2000; p2.gcrNkrylov = 16; p2.reliable_delta = 1e-4
    p2.use_init_guess = q.QUDA_USE_INIT_GUESS_YES
    x2 = x.copy()
    L.invertQuda(vp(x2), vp(b), C.byref(p2))
    assert p2.iter == 0 and p2.true_res < 1e-8
    assert np.linalg.norm(x2 - x) / np.linalg.norm(x) < 1e-12
    assert host_residual(oracle, g, x2, b, kappa, mu) < 5e-9
    assert math.isnan(p2.true_res_hq)   # not computed: must not read as a converged heavy-quark residual


def test_make_resident_solution_keeps_the_solution_on_the_device(quda, oracle):
    q, L = quda, quda.lib()
    kappa, mu = 0.12, 0.1
    g, b = setup(q, oracle)
    p = mg_inv_param(q, kappa, mu, sloppy=8, precond=8)
    p.tol = 1e-9; p.maxiter = 2000; p.gcrNkrylov = 16; p.reliable_delta = 1e-4
    p.make_resident_solution = 1
    x = np.full_like(b, 7.0)
    L.invertQuda(vp(x), vp(b), C.byref(p))
    assert (x == 7.0).all()            # h_x is not written (interface_quda.cpp:2493-2497)
    f = L.residentSolutionQudaB200()
    assert f
    L.saveSpinorQudaB200(vp(x), f, C.byref(p))
    assert host_residual(oracle, g, x, b, kappa, mu) < 5e-9
    # the resident field is a full field of the outer precision: apply M to it on the device and compare with b
    out = L.newSpinorQudaB200(q.QUDA_FULL_SITE_SUBSET, q.QUDA_DOUBLE_PRECISION)
    L.matResidentQudaB200(out, f, C.byref(p))
    mb = np.zeros_like(b)
    L.saveSpinorQudaB200(vp(mb), out, C.byref(p))
    L.freeSpinorQudaB200(out)
    assert np.linalg.norm(mb - b) / np.linalg.norm(b) < 5e-9
    # the next ordinary solve writes h_x again and leaves the resident field alone
    p.make_resident_solution = 0
    x3 = np.zeros_like(b)
    L.invertQuda(vp(x3), vp(b), C.byref(p))
    assert host_residual(oracle, g, x3, b, kappa, mu) < 5e-9 and L.residentSolutionQudaB200() == f


@pytest.mark.parametrize("inv", ["gcr", "bicgstab", "cg"])
def test_half_precision_sloppy_solver_vectors(quda, oracle, inv):
    """cuda_prec_sloppy = half: int16 + norm Krylov vectors and an int16 sloppy operator inside reliable updates / defect correction in the
    outer precision (the reference's mixed-precision mode, lib/inv_gcr_quda.cpp, lib/blas_core.h:12-52).  The solution must still reach
    the requested residual of the fp64 operator."""
    q, L = quda, quda.lib()
    kappa, mu = 0.12, 0.1
    g, b = setup(q, oracle, sloppy=2)
    p = mg_inv_param(q, kappa, mu, sloppy=2, precond=2)
    if inv == "cg":
        p.solve_type = q.QUDA_NORMOP_PC_SOLVE
        p.inv_type = q.QUDA_CG_INVERTER
        p.tol = 1e-9; p.maxiter = 8000; p.reliable_delta = 0.1
    else:
        p.solve_type = q.QUDA_DIRECT_PC_SOLVE
        p.inv_type = q.QUDA_GCR_INVERTER if inv == "gcr" else q.QUDA_BICGSTAB_INVERTER
        p.tol = 1e-8; p.maxiter = 4000; p.gcrNkrylov = 16; p.reliable_delta = 1e-2
    x = np.zeros_like(b)
    L.invertQuda(vp(x), vp(b), C.byref(p))
    res = host_residual(oracle, g, x, b, kappa, mu)
    print(f"{inv} with half-precision sloppy vectors: {p.iter} iterations, host residual {res:.2e}")
    assert res < (5e-8 if inv != "cg" else 5e-7) and p.iter > 0


def test_mg_gcr_with_half_precision_krylov_space(quda, oracle):
    """outer GCR in int16 vectors around the multigrid preconditioner (fp32 inside): converges to the same residual in about the same
    number of iterations as with fp32 Krylov vectors"""
    from tests.test_multigrid_gpu import point_source
    q, L = quda, quda.lib()
    X, kappa, mu = (8, 8, 8, 16), 0.1245, 0.005
    oracle.set_dims(X)
    g = oracle.weak_gauge(eps=0.25, antiperiodic=False, seed=4711)
    iters = {}
    for sloppy in (4, 2):
        load_gauge(q, g, X, prec=8, sloppy=sloppy, precond=4, recon=12)
        ip = mg_inv_param(q, kappa, mu, sloppy=4)
        mgp = q.multigrid_param(ip, n_level=2, geo_block=((4, 4, 4, 4),), n_vec=(8,), setup_maxiter=200, setup_tol=5e-6)
        mg = L.newMultigridQuda(C.byref(mgp))
        b = point_source(oracle.V); x = np.zeros_like(b)
        p = mg_inv_param(q, kappa, mu, sloppy=sloppy)
        p.inv_type_precondition = q.QUDA_MG_INVERTER; p.preconditioner = mg
        p.gcrNkrylov = 20; p.tol = 1e-8; p.maxiter = 300; p.reliable_delta = 1e-2
        L.invertQuda(vp(x), vp(b), C.byref(p))
        res = host_residual(oracle, g, x, b, kappa, mu)
        L.destroyMultigridQuda(mg)
        iters[sloppy] = p.iter
        assert res < 5e-8, (sloppy, res)
    print(f"MG-GCR iterations with fp32 / int16 Krylov vectors: {iters[4]} / {iters[2]}")
    assert iters[2] <= 2 * iters[4] + 4
