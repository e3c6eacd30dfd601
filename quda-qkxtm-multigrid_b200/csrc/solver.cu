// GCR / MR / BiCGStab on resident fields.
#include <chrono>
#include <algorithm>
#include <cmath>
#include "solver.h"

namespace qb {

static int g_half_vectors = -1;
bool half_vectors_enabled() {
  if (g_half_vectors < 0) { const char *e = getenv("QB_HALF_VECTORS"); g_half_vectors = (e && atoi(e) == 0) ? 0 : 1; }
  return g_half_vectors == 1;
}
void set_half_vectors(bool on) { g_half_vectors = on ? 1 : 0; }

using blas::Complex;

SpinorField *new_like(const SpinorField &a, Prec prec) { return new SpinorField(a.Vh, a.nparity, prec, a.nspin, a.ncolor, 1, a.nflavor); }

static void ensure(std::unique_ptr<SpinorField> &f, const SpinorField &like, Prec prec) {
  if (!f || f->Vh != like.Vh || f->nparity != like.nparity || f->ncomplex != like.ncomplex || f->nflavor != like.nflavor || f->prec != prec) f.reset(new_like(like, prec));
}

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

Solver *Solver::create(SolverParam &param, const DiracMatrix &mat, const DiracMatrix &matSloppy, const DiracMatrix &matPrecon, Solver *K) {
  switch (param.inv_type) {
    case INV_GCR: return new GCR(mat, matSloppy, matPrecon, param, K);
    case INV_MR: return new MR(mat, matSloppy, param);
    case INV_BICGSTAB: return new BiCGStab(mat, matSloppy, param);
    case INV_CG: return new CG(mat, matSloppy, param);
    default: QB_ERROR("Invalid solver type %d (this build provides CG, GCR, MR and BiCGStab)", (int)param.inv_type);
  }
}

// -------------------------------------------------------------------------------------------------
// CG (lib/inv_cg_quda.cpp:37-330): iteration in the sloppy precision, solution accumulated in the outer
// precision; a reliable update (true residual from the outer-precision operator, sloppy accumulator folded
// into y) whenever the iterated residual has dropped by `delta` relative to its maximum since the last
// update (:178-240).
// -------------------------------------------------------------------------------------------------
void CG::operator()(SpinorField &x, SpinorField &b) {
  const Prec px = x.prec, ps = blas_prec(param.precision_sloppy);
  const double t0 = now_s();
  ensure(r, x, px); ensure(y, x, px); ensure(p, x, ps); ensure(Ap, x, ps);
  SpinorField *rs = r.get(), *xs = &x;
  const bool mixed = px != ps;
  if (mixed) { ensure(rS, x, ps); ensure(xS, x, ps); ensure(tmp, x, px); rs = rS.get(); xs = xS.get(); }
  const double b2 = blas::norm2(b);
  if (b2 == 0.0) { blas::zero(x); param.true_res = 0.0; return; }
  double r2;
  if (param.use_init_guess) {
    mat(*r, x);
    r2 = blas::xmyNorm(b, *r);
    blas::copy(*y, x);
  } else {
    blas::copy(*r, b);
    r2 = b2;
    blas::zero(*y);
  }
  blas::zero(*xs);   // xs == &x when not mixed: x accumulates the correction, y holds the initial guess
  if (mixed) blas::copy(*rS, *r);
  blas::copy(*p, *rs);
  const double stop = param.tol * param.tol * b2;
  const double delta = param.delta > 0.0 && param.delta < 1.0 ? param.delta : 1e-1;
  double rNorm = sqrt(r2), r0Norm = rNorm, maxrx = rNorm, maxrr = rNorm;
  int k = 0, updates = 0;
  while (r2 > stop && k < param.maxiter) {
    matSloppy(*Ap, *p);
    const double pAp = blas::reDotProduct(*p, *Ap);
    if (!(pAp > 0.0)) { log_msg(1, "CG: <p, A p> = %e is not positive: the operator is not Hermitian positive definite (use a NORMOP solve type)\n", pAp); break; }
    const double alpha = r2 / pAp;
    const double r2_old = r2;
    r2 = blas::axpyNorm(-alpha, *Ap, *rs);
    blas::axpy(alpha, *p, *xs);
    rNorm = sqrt(r2);
    if (rNorm > maxrx) maxrx = rNorm;
    if (rNorm > maxrr) maxrr = rNorm;
    const bool updateX = rNorm < delta * r0Norm && r0Norm <= maxrx;
    const bool updateR = (rNorm < delta * maxrr && r0Norm <= maxrr) || updateX;
    k++;
    if (mixed && (updateR || !(r2 > stop))) {
      // reliable update: y += xS, r = b - A y in the outer precision, restart the sloppy accumulator
      blas::copy(*tmp, *xS);
      blas::xpy(*tmp, *y);
      mat(*r, *y);
      r2 = blas::xmyNorm(b, *r);
      blas::copy(*rS, *r);
      blas::zero(*xS);
      rNorm = sqrt(r2); maxrr = rNorm; maxrx = rNorm; r0Norm = rNorm;
      updates++;
    }
    blas::xpay(*rs, r2 / r2_old, *p);   // p = r + beta p
    if (param.verbosity >= 2) log_msg(2, "CG: %d iterations, <r,r> = %e, |r|/|b| = %e\n", k, r2, sqrt(r2 / b2));
  }
  // x = y + accumulated correction
  if (mixed) { blas::copy(*tmp, *xS); blas::xpy(*tmp, *y); blas::copy(x, *y); }
  else blas::xpy(*y, x);
  param.iter += k;
  param.secs += now_s() - t0;
  if (param.compute_true_res) {
    mat(*r, x);
    param.true_res = sqrt(blas::xmyNorm(b, *r) / b2);
  } else param.true_res = sqrt(r2 / b2);
  if (k == param.maxiter && r2 > stop) log_msg(1, "CG: exceeded maximum iterations %d\n", param.maxiter);
  log_msg(1, "CG: Convergence at %d iterations (%d reliable updates), L2 relative residual: iterated = %e, true = %e\n", k, updates, sqrt(r2 / b2), param.true_res);
}

// -------------------------------------------------------------------------------------------------
// MR: fixed-iteration minimal residual with relaxation omega (lib/inv_mr_quda.cpp:37-198)
//   alpha = <Ar, r> / <Ar, Ar>;  x += omega alpha r;  r -= omega alpha Ar
// The residual is normalised by |r0| up front ("domain-wise normalisation", :95-99) and scaled back.
// -------------------------------------------------------------------------------------------------
void MR::operator()(SpinorField &x, SpinorField &b) {
  blas::set_global_reduction(param.global_reduction);
  const Prec ps = blas_prec(param.precision_sloppy);
  ensure(r, x, ps); ensure(Ar, x, ps); ensure(y, x, ps);
  const double t0 = param.is_preconditioner ? 0.0 : now_s();

  // Smoother fast path (fp32 vectors, fixed iteration count, source preserved): the same iteration without the |r0| normalisation
  // passes (alpha = <Ar,r>/<Ar,Ar> does not depend on the scale of r; the reference normalises for its int16 vectors,
  // inv_mr_quda.cpp:95-99), started from b itself instead of a copy, and with the residual update of the last step dropped:
  // 11 field streams instead of 24 for MR(2) from a zero guess, 15 instead of 26 with an initial guess.
  if (param.is_preconditioner && param.preserve_source && x.prec == ps && b.prec == ps && ps != PREC_DOUBLE && !getenv("QB_MR_PLAIN")) {
    const SpinorField *rc = &b;
    bool x_valid = param.use_init_guess;
    if (param.use_init_guess) {
      matSloppy(*r, x);
      blas::axpby(1.0, b, -1.0, *r);  // r = b - A x0
      rc = r.get();
    }
    int k = 0;
    residual_valid = false;
    while (k < param.maxiter) {
      matSloppy(*Ar, *rc);
      const blas::double3_ d = blas::cDotProductNormA(*Ar, *rc);
      if (d.z == 0.0) break;
      const Complex alpha = param.omega * Complex(d.x, d.y) / d.z;
      const bool last = k == param.maxiter - 1;
      if (last && !keep_residual) {
        if (x_valid) blas::caxpy(alpha, *rc, x);
        else blas::cax(alpha, *rc, x);
      } else if (rc == &b) {
        blas::mrFirstStep(alpha, b, *Ar, x, *r, x_valid);
        rc = r.get();
      } else {
        blas::caxpyXmaz(alpha, *r, x, *Ar);  // x += alpha r; r -= alpha A r
      }
      x_valid = true;
      k++;
      if (last && keep_residual) residual_valid = true;
    }
    if (!x_valid) blas::zero(x);
    param.iter += k;
    blas::set_global_reduction(true);
    return;
  }

  double r2;
  if (param.use_init_guess) {
    const SpinorField *x0 = &x;
    if (x.prec != ps) { ensure(xs, x, ps); blas::copy(*xs, x); x0 = xs.get(); }
    matSloppy(*r, *x0);
    blas::copy(*y, b);
    r2 = blas::xmyNorm(*y, *r);  // r = b - A x0
  } else {
    blas::copy(*r, b);
    r2 = blas::norm2(*r);
    blas::zero(x);
  }
  blas::zero(*y);
  const double b2 = param.is_preconditioner ? r2 : blas::norm2(b);
  const double c2 = r2;
  if (c2 > 0.0) { blas::ax(1.0 / sqrt(c2), *r); r2 = 1.0; }

  int k = 0;
  while (k < param.maxiter && r2 > 0.0) {
    matSloppy(*Ar, *r);
    const blas::double3_ d = blas::cDotProductNormA(*Ar, *r);
    if (d.z == 0.0) break;
    const Complex alpha = Complex(d.x, d.y) / d.z;
    blas::caxpyXmaz(param.omega * alpha, *r, *y, *Ar);
    k++;
    if (param.verbosity >= 3) log_msg(3, "MR: %d iterations, <r|A|r> = (%e, %e)\n", k, d.x, d.y);
  }

  const double scale = c2 > 0.0 ? sqrt(c2) : 1.0;
  if (x.prec == y->prec) {
    if (param.use_init_guess) blas::axpy(scale, *y, x);
    else blas::axpby(scale, *y, 0.0, x);
  } else {
    blas::ax(scale, *y);
    ensure(yx, x, x.prec);
    blas::copy(*yx, *y);
    if (param.use_init_guess) blas::xpy(*yx, x);
    else blas::copy(x, *yx);
  }
  if (!param.preserve_source && b.prec == r->prec) {
    if (c2 > 0.0) blas::axpby(scale, *r, 0.0, b);  // hand the residual back in the source (GCR inner solve)
  }
  param.iter += k;
  if (!param.is_preconditioner) {
    param.secs += now_s() - t0;
    if (param.compute_true_res) {
      std::unique_ptr<SpinorField> t(new_like(x, x.prec));
      mat(*t, x);
      std::unique_ptr<SpinorField> bb(new_like(x, x.prec));
      blas::copy(*bb, b);
      param.true_res = sqrt(blas::xmyNorm(*bb, *t) / b2);
      log_msg(1, "MR: Converged after %d iterations, relative residual: true = %e\n", k, param.true_res);
    }
  }
  blas::set_global_reduction(true);
}

// -------------------------------------------------------------------------------------------------
// GCR(nKrylov), flexible (right-preconditioned by K), mixed precision with reliable restarts
// (lib/inv_gcr_quda.cpp:235-516).  Orthogonalisation of A p_k against the previous directions is done
// as ONE block dot + ONE block caxpy (classical Gram-Schmidt in a single pass over memory each).
// -------------------------------------------------------------------------------------------------
GCR::GCR(const DiracMatrix &mat_, const DiracMatrix &matSloppy_, const DiracMatrix &matPrecon_, SolverParam &p_, Solver *K_)
    : Solver(p_), mat(mat_), matSloppy(matSloppy_), matPrecon(matPrecon_), K(K_), nKrylov(p_.Nkrylov) {
  if (nKrylov < 1) QB_ERROR("Parameter gcrNkrylov undefined");
  if (nKrylov > 32) QB_ERROR("gcrNkrylov = %d exceeds the supported maximum of 32", nKrylov);
  if (!K && param.inv_type_precondition == INV_MR) {
    // inner MR as in fillInnerSolveParam (inv_gcr_quda.cpp:22-51)
    Kparam = param;
    Kparam.inv_type = INV_MR;
    Kparam.inv_type_precondition = INV_NONE;
    Kparam.is_preconditioner = true;
    Kparam.global_reduction = false;
    Kparam.use_init_guess = false;
    Kparam.preserve_source = true;
    Kparam.precision = Kparam.precision_sloppy = param.precision_precondition;
    K = new MR(matPrecon, matPrecon, Kparam);
    own_K = true;
  }
  p.resize(nKrylov); Ap.resize(nKrylov);
  alpha.resize(nKrylov); beta.resize((size_t)nKrylov * nKrylov); gamma.resize(nKrylov);
}

GCR::~GCR() { if (own_K) delete K; }

void GCR::operator()(SpinorField &x, SpinorField &b) {
  const Prec pp = blas_prec(param.precision), ps = blas_prec(param.precision_sloppy);
  const bool mixed = (pp != ps) || x.prec != pp;
  const double t0 = param.is_preconditioner ? 0.0 : now_s();
  if (x.prec != b.prec) QB_ERROR("GCR: x and b precision differ");
  const Prec px = x.prec;
  ensure(r, x, px); ensure(y, x, px); ensure(tmp, x, ps);
  for (int i = 0; i < nKrylov; i++) { ensure(p[i], x, ps); ensure(Ap[i], x, ps); }
  SpinorField *xS = &x, *rS = r.get();
  if (px != ps) { ensure(x_sloppy, x, ps); ensure(r_sloppy, x, ps); xS = x_sloppy.get(); rS = r_sloppy.get(); }
  (void)mixed;

  blas::zero(*y);
  const double b2 = blas::norm2(b);
  double r2;
  if (param.use_init_guess) {
    mat(*r, x);
    r2 = blas::xmyNorm(b, *r);
    blas::copy(*y, x);
    if (xS == &x) blas::zero(x);
  } else {
    blas::copy(*r, b);
    r2 = b2;
    blas::zero(x);
  }
  if (xS != &x) blas::zero(*xS);
  if (b2 == 0.0) {
    blas::zero(x);
    param.true_res = 0.0;
    return;
  }
  const double stop = param.tol * param.tol * b2;  // L2 relative residual
  if (rS != r.get()) blas::copy(*rS, *r);

  int total_iter = 0, restart = 0, k = 0, resIncrease = 0, resIncreaseTotal = 0;
  double r2_old = r2;
  bool l2_converge = false;
  if (param.verbosity >= 2) log_msg(2, "GCR%s: %d iterations, <r,r> = %e, |r|/|b| = %e\n", param.name, total_iter, r2, sqrt(r2 / b2));

  while (r2 > stop && total_iter < param.maxiter) {
    // p_k = K r  (flexible preconditioning) or p_k = r
    if (K) {
      if (p[k]->prec == rS->prec) (*K)(*p[k], *rS);
      else QB_ERROR("GCR: preconditioner precision handling expects sloppy-precision vectors");
    } else {
      blas::copy(*p[k], *rS);
    }
    matSloppy(*Ap[k], *p[k]);

    // orthogonalise A p_k against A p_0..k-1 : beta_i = <Ap_i, Ap_k>; Ap_k -= sum beta_i Ap_i
    if (k > 0) {
      std::vector<SpinorField *> prev(k);
      for (int i = 0; i < k; i++) prev[i] = Ap[i].get();
      std::vector<Complex> bt(k);
      blas::cDotProduct(bt.data(), prev, *Ap[k]);
      for (int i = 0; i < k; i++) { beta[(size_t)i * nKrylov + k] = bt[i]; bt[i] = -bt[i]; }
      blas::caxpy(bt.data(), prev, *Ap[k]);
    }
    const blas::double3_ Apr = blas::cDotProductNormA(*Ap[k], *rS);
    gamma[k] = sqrt(Apr.z);
    if (gamma[k] == 0.0) QB_ERROR("GCR breakdown");
    alpha[k] = Complex(Apr.x, Apr.y) / gamma[k];
    // Ap_k /= gamma_k ; r -= alpha_k Ap_k
    r2 = blas::cabxpyAxNorm(1.0 / gamma[k], -alpha[k], *Ap[k], *rS);
    k++;
    total_iter++;
    if (param.verbosity >= 2) log_msg(2, "GCR%s: %d iterations, <r,r> = %e, |r|/|b| = %e\n", param.name, total_iter, r2, sqrt(r2 / b2));

    if (k == nKrylov || total_iter == param.maxiter || (r2 < stop && !l2_converge) || sqrt(r2 / r2_old) < param.delta) {
      // back substitution for the solution coefficients, then x += sum delta_i p_i
      std::vector<Complex> delta(k);
      for (int i = k - 1; i >= 0; i--) {
        delta[i] = alpha[i];
        for (int j = i + 1; j < k; j++) delta[i] -= beta[(size_t)i * nKrylov + j] * delta[j];
        delta[i] /= gamma[i];
      }
      std::vector<SpinorField *> P(k);
      for (int i = 0; i < k; i++) P[i] = p[i].get();
      blas::caxpy(delta.data(), P, *xS);
      // reliable update in the outer precision
      if (xS != &x) blas::copy(x, *xS);
      blas::xpy(x, *y);
      if (param.is_preconditioner && !param.compute_true_res && (r2 < stop || total_iter == param.maxiter)) {
        // inner solves skip the final true-residual mat-vec
        k = 0;
        break;
      }
      mat(*r, *y);
      r2 = blas::xmyNorm(b, *r);
      if (r2 > r2_old) {
        resIncrease++; resIncreaseTotal++;
        log_msg(2, "GCR: new reliable residual norm %e is greater than previous reliable residual norm %e (total #inc %i)\n", sqrt(r2), sqrt(r2_old), resIncreaseTotal);
        if (resIncrease > param.max_res_increase || resIncreaseTotal > param.max_res_increase_total) {
          log_msg(1, "GCR: solver exiting due to too many true residual norm increases\n");
          k = 0;
          break;
        }
      } else resIncrease = 0;
      k = 0;
      if (r2 > stop) {
        restart++;
        if (rS != r.get()) blas::copy(*rS, *r);
        blas::zero(*xS);
        if (r2 < stop) l2_converge = true;
      }
      r2_old = r2;
    }
  }
  // y carries the accumulated solution (the initial guess was moved there and x zeroed): restore it also when the guess already
  // satisfied the tolerance and no iteration ran
  if (total_iter > 0 || param.use_init_guess) blas::copy(x, *y);
  param.iter += total_iter;
  if (!param.is_preconditioner) {
    param.secs += now_s() - t0;
    if (param.compute_true_res) {
      mat(*r, x);
      param.true_res = sqrt(blas::xmyNorm(b, *r) / b2);
    }
    log_msg(1, "GCR%s: Convergence at %d iterations (%d restarts), L2 relative residual: iterated = %e, true = %e\n", param.name, total_iter,
            restart, sqrt(r2 / b2), param.true_res);
  }
}

// -------------------------------------------------------------------------------------------------
// BiCGStab (lib/inv_bicgstab_quda.cpp).  Used for null-vector generation: with compute_null_vector the
// right-hand side is zero and `x` holds the random initial guess (b2 := |r0|^2 of the guess, :95-125).
// -------------------------------------------------------------------------------------------------
void BiCGStab::operator()(SpinorField &x, SpinorField &b) {
  const Prec ps = blas_prec(param.precision_sloppy);
  if (!param.compute_null_vector && x.prec != ps) {
    // mixed precision: defect correction around the sloppy-precision iteration (the reference's reliable
    // updates, inv_bicgstab_quda.cpp, play the same role): r = b - A x in the outer precision, A e = r sloppy, x += e
    ensure(y, x, x.prec); ensure(xs, x, ps); ensure(rs, x, ps);
    std::unique_ptr<SpinorField> rp(new_like(x, x.prec)), es(new_like(x, ps));
    const double t0 = now_s();
    const double b2 = blas::norm2(b);
    if (b2 == 0.0) { blas::zero(x); return; }
    if (!param.use_init_guess) blas::zero(x);
    const double stop = param.tol * param.tol * b2;
    SolverParam inner = param;
    inner.precision = ps;
    inner.use_init_guess = false;
    inner.compute_null_vector = false;
    inner.iter = 0;
    BiCGStab in_solver(matSloppy, matSloppy, inner);
    double r2 = b2;
    for (int cycle = 0; cycle < 50; cycle++) {
      mat(*rp, x);
      r2 = blas::xmyNorm(b, *rp);  // rp = b - A x
      if (r2 <= stop || inner.iter >= param.maxiter) break;
      inner.tol = std::max(0.5 * sqrt(stop / r2), param.delta > 0 && param.delta < 1 ? param.delta : 1e-3);
      inner.maxiter = param.maxiter - inner.iter;
      blas::copy(*rs, *rp);
      in_solver(*es, *rs);
      blas::copy(*y, *es);
      blas::xpy(*y, x);
    }
    param.iter += inner.iter;
    param.secs += now_s() - t0;
    param.true_res = sqrt(r2 / b2);
    log_msg(1, "BiCGstab: Convergence at %d iterations, L2 relative residual: true = %e\n", inner.iter, param.true_res);
    return;
  }
  ensure(r, x, ps); ensure(r0, x, ps); ensure(p, x, ps); ensure(v, x, ps); ensure(t, x, ps); ensure(xs, x, ps);
  const double t0 = now_s();
  blas::copy(*xs, x);
  double b2;
  if (param.compute_null_vector) {
    matSloppy(*r, *xs);
    blas::ax(-1.0, *r);  // r = 0 - A x0
    b2 = blas::norm2(*r);
  } else if (param.use_init_guess) {
    ensure(rs, x, ps);
    blas::copy(*rs, b);
    matSloppy(*r, *xs);
    blas::xmyNorm(*rs, *r);
    b2 = blas::norm2(b);
  } else {
    blas::copy(*r, b);
    blas::zero(*xs);
    b2 = blas::norm2(b);
  }
  double r2 = blas::norm2(*r);
  if (b2 == 0.0) { blas::zero(x); return; }
  const double stop = param.tol * param.tol * b2;
  blas::copy(*r0, *r);
  blas::zero(*p); blas::zero(*v);
  Complex rho(1.0, 0.0), rho0(1.0, 0.0), alpha(1.0, 0.0), omega(1.0, 0.0);
  int k = 0;
  Complex rho_next = blas::cDotProduct(*r0, *r);
  while (r2 > stop && k < param.maxiter) {
    rho0 = rho;
    rho = rho_next;
    if (std::abs(rho0) == 0.0 || std::abs(omega) == 0.0) break;
    const Complex beta = (rho / rho0) * (alpha / omega);
    // p = r + beta (p - omega v)
    blas::cxpaypbz(*r, -beta * omega, *v, beta, *p);
    matSloppy(*v, *p);
    const Complex r0v = blas::cDotProduct(*r0, *v);
    if (std::abs(r0v) == 0.0) break;
    alpha = rho / r0v;
    blas::caxpy(-alpha, *v, *r);  // s = r - alpha v (kept in r)
    matSloppy(*t, *r);
    const blas::double3_ ts = blas::cDotProductNormA(*t, *r);
    if (ts.z == 0.0) { blas::caxpy(alpha, *p, *xs); r2 = blas::norm2(*r); k++; break; }
    omega = Complex(ts.x, ts.y) / ts.z;
    // x += alpha p + omega s ; r = s - omega t ; <r0, r> of the next iteration and |r|^2 in the same pass
    const blas::double3_ up = blas::bicgstabUpdate(alpha, *p, omega, *r, *t, *xs, *r0);
    rho_next = Complex(up.x, up.y);
    r2 = up.z;
    k++;
    if (param.verbosity >= 3) log_msg(3, "BiCGstab: %d iterations, <r,r> = %e, |r|/|b| = %e\n", k, r2, sqrt(r2 / b2));
  }
  blas::copy(x, *xs);
  param.iter += k;
  param.secs += now_s() - t0;
  param.true_res = sqrt(r2 / b2);
  log_msg(2, "BiCGstab: %d iterations, iterated relative residual %e\n", k, param.true_res);
}

}  // namespace qb
