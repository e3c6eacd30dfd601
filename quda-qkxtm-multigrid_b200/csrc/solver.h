// Krylov solvers used on the multigrid path: GCR (outer, coarsest, K-cycle), MR (smoother),
// BiCGStab (null-vector setup).  Same roles and parameter meaning as the reference's
// lib/inv_gcr_quda.cpp, lib/inv_mr_quda.cpp, lib/inv_bicgstab_quda.cpp, lib/solver.cpp:13-80.
#pragma once
#include <memory>
#include <vector>
#include "blas.h"
#include "dirac.h"

namespace qb {

enum InverterType { INV_CG = 0, INV_BICGSTAB = 1, INV_GCR = 2, INV_MR = 3, INV_MG = 15, INV_NONE = -1 };

struct SolverParam {
  InverterType inv_type = INV_GCR;
  InverterType inv_type_precondition = INV_NONE;
  double tol = 1e-7;
  int maxiter = 1000;
  int Nkrylov = 20;
  double delta = 1e-3;          // reliable_delta
  double omega = 1.0;           // MR relaxation
  Prec precision = PREC_DOUBLE, precision_sloppy = PREC_DOUBLE, precision_precondition = PREC_DOUBLE;
  bool use_init_guess = false;
  bool preserve_source = true;
  bool is_preconditioner = false;
  bool global_reduction = true;
  bool compute_true_res = true;
  bool compute_null_vector = false;  // BiCGStab: solve M x = 0 from a random initial guess
  int pipeline = 0;
  int precondition_cycle = 1;
  int max_res_increase = 1, max_res_increase_total = 10;
  int verbosity = 0;
  const char *name = "";
  // written back
  double true_res = 0.0;
  int iter = 0;
  double secs = 0.0, gflops = 0.0;
};

class Solver {
 public:
  SolverParam &param;
  explicit Solver(SolverParam &p) : param(p) {}
  virtual ~Solver() {}
  virtual void operator()(SpinorField &x, SpinorField &b) = 0;
  virtual long long flops() const { return 0; }
  // factory (lib/solver.cpp:13-80); K may be null
  static Solver *create(SolverParam &param, const DiracMatrix &mat, const DiracMatrix &matSloppy, const DiracMatrix &matPrecon, Solver *K = nullptr);
};

class MR : public Solver {
  DiracMatrix mat, matSloppy;
  std::unique_ptr<SpinorField> r, Ar, y, xs, yx;
 public:
  MR(const DiracMatrix &mat_, const DiracMatrix &matSloppy_, SolverParam &p) : Solver(p), mat(mat_), matSloppy(matSloppy_) {}
  void operator()(SpinorField &x, SpinorField &b) override;
  // Smoother use (the multigrid's use_solver_residual, lib/multigrid.cpp:536-546): with keep_residual set the fixed-iteration fast path
  // also carries out the residual update of its last step (one BLAS pass instead of the operator application the caller would need to
  // recompute b - A x) and residual() returns b - A x of the solution just produced; nullptr when the path taken did not keep it.
  bool keep_residual = false;
  const SpinorField *residual() const { return residual_valid ? r.get() : nullptr; }
 private:
  bool residual_valid = false;
};

class GCR : public Solver {
  DiracMatrix mat, matSloppy, matPrecon;
  Solver *K;              // preconditioner (multigrid V/K-cycle, MR, ...), not owned unless own_K
  bool own_K = false;
  SolverParam Kparam;
  int nKrylov;
  std::vector<std::unique_ptr<SpinorField>> p, Ap;
  std::unique_ptr<SpinorField> r, y, x_sloppy, r_sloppy, r_pre, p_pre, tmp;
  std::vector<blas::Complex> alpha, beta;  // beta[i * nKrylov + k]
  std::vector<double> gamma;
 public:
  GCR(const DiracMatrix &mat_, const DiracMatrix &matSloppy_, const DiracMatrix &matPrecon_, SolverParam &p, Solver *K_ = nullptr);
  ~GCR();
  void operator()(SpinorField &x, SpinorField &b) override;
  // iteration count / tolerance of the internally created MR preconditioner (maxiter_precondition, tol_precondition)
  void set_inner(int maxiter, double tol) { if (own_K) { Kparam.maxiter = maxiter; Kparam.tol = tol; } }
};

// Conjugate gradient on a Hermitian positive definite operator (the normal operator M^dag M of the NORMOP solve types),
// mixed precision with reliable updates (lib/inv_cg_quda.cpp)
class CG : public Solver {
  DiracMatrix mat, matSloppy;
  std::unique_ptr<SpinorField> r, y, p, Ap, rS, xS, tmp;
 public:
  CG(const DiracMatrix &mat_, const DiracMatrix &matSloppy_, SolverParam &p_) : Solver(p_), mat(mat_), matSloppy(matSloppy_) {}
  void operator()(SpinorField &x, SpinorField &b) override;
};

class BiCGStab : public Solver {
  DiracMatrix mat, matSloppy;
  std::unique_ptr<SpinorField> r, r0, p, v, t, y, xs, rs;
 public:
  BiCGStab(const DiracMatrix &mat_, const DiracMatrix &matSloppy_, SolverParam &p) : Solver(p), mat(mat_), matSloppy(matSloppy_) {}
  void operator()(SpinorField &x, SpinorField &b) override;
};

SpinorField *new_like(const SpinorField &a, Prec prec);
// Precision of the solver's vectors for an operator precision p.  int16 + norm vectors (the reference's half precision, lib/blas_core.h:12-52)
// are used for single fine-grid fields; doublet fields and QB_HALF_VECTORS=0 keep fp32 vectors in front of a half-precision operator.
bool half_vectors_enabled();
void set_half_vectors(bool on);
inline Prec blas_prec(Prec p) { return (p == PREC_HALF && !half_vectors_enabled()) ? PREC_SINGLE : p; }

}  // namespace qb
