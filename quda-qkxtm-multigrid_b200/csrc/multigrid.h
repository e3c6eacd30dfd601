// Adaptive multigrid preconditioner (replaces /root/reference/lib/multigrid.cpp, include/multigrid.h).
// Level construction (:11-295), V / K cycle (:488-604), null-vector generation (:693-779) and the
// consistency checks of MG::verify (:372-486, disabled in the reference fork, enabled here for tests).
#pragma once
#include <memory>
#include <string>
#include <vector>
#include "coarse.h"
#include "solver.h"
#include "transfer.h"

namespace qb {

struct MGLevelParam {
  int geo_bs[4] = {4, 4, 4, 4};
  int spin_bs = 2;
  int nvec = 24;
  InverterType smoother = INV_MR;
  bool smoother_pc = true;          // smoother_solve_type == QUDA_DIRECT_PC_SOLVE
  // coarse_grid_solution_type == QUDA_MATPC_SOLUTION: this level hands a single-parity residual of the even-odd system to the next level,
  // whose operator is then built by preconditioned coarsening (multigrid.cpp:145-155); needs smoother_pc
  bool coarse_pc = false;
  int nu_pre = 2, nu_post = 2;
  double smoother_tol = 0.25;
  double omega = 0.85;
  bool recursive = true;            // cycle_type == QUDA_MG_CYCLE_RECURSIVE (K-cycle)
  bool global_reduction = true;
};

struct MGParam {
  int n_level = 2;
  MGLevelParam level[4];
  int setup_maxiter = 500;
  double setup_tol = 5e-6;
  bool compute_null_vector = true;
  bool generate_all_levels = true;
  int verbosity = 1;
  std::string vec_infile, vec_outfile;   // near-null vector files (QudaMultigridParam::vec_infile / vec_outfile), "" = none
  // 16-bit storage of the preconditioner's data (V for P / R, coarse links for the single-RHS coarse kernel), fp32 arithmetic:
  // on when cuda_prec_precondition = half (or QB_MG_HALF_STORAGE=1); the outer solver's accuracy is unaffected
  bool half_storage = false;
  bool keep_null_vectors = true;  // false: free the near-null vectors of a level once V and the coarse vectors exist (24 x 96 B/site)
};

class MG : public Solver {
 public:
  MGParam &mp;
  int level;
  const Dirac *matResidual;   // full operator of this level (fp32 vectors)
  const Dirac *matSmooth;     // even-odd preconditioned (or full) operator used by the smoother
  SolverParam dummy;
  // next level
  std::unique_ptr<Transfer> transfer;
  std::shared_ptr<CoarseOperator> coarse_op;
  std::unique_ptr<DiracCoarse> coarseResidual, coarseSmooth;
  std::unique_ptr<MG> coarse;
  std::unique_ptr<Solver> coarse_solver_gcr;  // K-cycle wrapper
  std::unique_ptr<Solver> coarse_solver_pc;   // K-cycle on the next level's even-odd system: prepare -> coarse_solver_gcr -> reconstruct
  bool pc_coarsen = false;    // the next level's operator is the Galerkin product of S^-1 M (preconditioned coarsening)
  int pc_parity = 0;          // parity the even-odd system of this level lives on (matpc_type of the smoother operator)
  SolverParam param_coarse_solver, param_presmooth, param_postsmooth;
  std::unique_ptr<Solver> presmoother, postsmoother;
  std::vector<std::unique_ptr<SpinorField>> B;   // near-null vectors of this level
  std::unique_ptr<SpinorField> r, r_coarse, x_coarse, b_copy, x32, b32;
  double setup_secs = 0.0;
  double t_prof[6] = {0, 0, 0, 0, 0, 0};
  long ncycle = 0;
  void print_profile();

  MG(MGParam &mp, int level, const Dirac *matResidual, const Dirac *matSmooth, std::vector<std::unique_ptr<SpinorField>> *B_in);
  void operator()(SpinorField &x, SpinorField &b) override;
  // deviations of the identities checked by MG::verify: {|P^dag P eta - eta|, max_k |P R v_k - v_k|, |R M P eta - M_c eta|} (relative)
  void verify(double *dev3);
  long long flops() const override;

  void smooth(Solver &s, SpinorField &x, SpinorField &b);

 private:
  void generate_null_vectors();
  void load_vectors();
  void save_vectors();
  void cycle(SpinorField &x, SpinorField &b);
  void cycle_pc(SpinorField &x, SpinorField &b);   // single-parity fields of the even-odd system (multigrid.cpp:494-560 with MATPC types)
};

// Solver on the even-odd system wrapped in prepare / reconstruct so that it takes full fields (PreconditionedSolver, include/invert_quda.h:598)
class PreconditionedSolver : public Solver {
  std::unique_ptr<Solver> solver;
  const Dirac *dirac;
 public:
  PreconditionedSolver(Solver *s, const Dirac *d, SolverParam &p) : Solver(p), solver(s), dirac(d) {}
  void operator()(SpinorField &x, SpinorField &b) override {
    SpinorField src, sol;
    dirac->prepare(src, sol, x, b, SOL_MAT);
    (*solver)(sol, src);
    dirac->reconstruct(x, b, SOL_MAT);
  }
};

// handle returned by newMultigridQuda (multigrid_solver of interface_quda.cpp:2161-2255)
struct MultigridSolver {
  MGParam mp;
  std::unique_ptr<DiracTM> dirac, diracSmooth;
  std::unique_ptr<MG> mg;
};

void random_fill(SpinorField &f, unsigned long long seed);
void mg_profile_enable(bool on);   // section timers of the cycle (stream-synchronised, so they slow the solve down)
// batched null-vector generation on coarse levels through the multi-RHS tensor-core operator (block_solver.cu)
bool block_null_vectors_supported(const Dirac *matSmooth, int nvec);
int block_null_vectors(const Dirac *matSmooth, std::vector<SpinorField *> &x, int maxiter, double tol);
// block multigrid (block_solver.cu): R right-hand sides through the K-cycle in lock-step, coarse levels on the multi-RHS tensor-core operator
bool block_mg_supported(const MG &mg, int R, int mode);
int block_mg_gcr_solve(MG &mg, const DiracMatrix &mat, const DiracMatrix &matSloppy, std::vector<SpinorField *> &x, std::vector<SpinorField *> &b,
                       const SolverParam &sp, int mode, std::vector<double> &true_res);

}  // namespace qb
