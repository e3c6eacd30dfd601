// Multigrid level construction and cycle.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <unistd.h>
#include "comm.h"
#include "multigrid.h"

namespace qb {

using blas::Complex;

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// deterministic counter-based uniform numbers in [0,1): value depends only on (seed, parity, site, component)
__global__ void random_fill_kernel(float4 *v, long Vh, int nplanes, int nparity, unsigned long long seed) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)nparity * nplanes * Vh) return;
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(4 * t + 1);
  float r[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    z += 0x9E3779B97F4A7C15ull;
    unsigned long long x = z;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    x ^= x >> 31;
    r[k] = (float)(x >> 40) * (1.0f / 16777216.0f);
  }
  v[t] = make_float4(r[0], r[1], r[2], r[3]);
}

void random_fill(SpinorField &f, unsigned long long seed) {
  if (f.prec != PREC_SINGLE) QB_ERROR("random_fill: single precision fields only");
  const long n = (long)f.nparity * f.planes() * f.Vh;
  random_fill_kernel<<<div_up(n, 256), 256, 0, rt().compute>>>((float4 *)f.v, f.Vh, f.planes(), f.nparity, seed);
  QB_CHECK_LAUNCH();
}

static SpinorField *new_full(const Dirac &d) {
  SpinorField *p = d.new_parity_field(PREC_SINGLE);
  SpinorField *f = new SpinorField(p->Vh, 2, PREC_SINGLE, p->nspin, p->ncolor);
  delete p;
  return f;
}

MG::MG(MGParam &mp_, int level_, const Dirac *matResidual_, const Dirac *matSmooth_, std::vector<std::unique_ptr<SpinorField>> *B_in)
    : Solver(dummy), mp(mp_), level(level_), matResidual(matResidual_), matSmooth(matSmooth_) {
  const double t0 = now_s();
  const MGLevelParam &lp = mp.level[level];
  const bool last = level == mp.n_level - 1;
  log_msg(1, "MG level %d: creating level %d of %d\n", level + 1, level + 1, mp.n_level);
  if (matResidual->is_pc()) QB_ERROR("MG: the residual operator of a level must be the full (unpreconditioned) operator");

  // ---- smoothers (multigrid.cpp:47-92) ----
  param_presmooth.inv_type = lp.smoother;
  param_presmooth.inv_type_precondition = INV_NONE;
  param_presmooth.is_preconditioner = true;
  param_presmooth.preserve_source = true;
  param_presmooth.use_init_guess = false;
  param_presmooth.maxiter = lp.nu_pre;
  param_presmooth.Nkrylov = 4;
  param_presmooth.tol = lp.smoother_tol;
  param_presmooth.omega = lp.omega;
  param_presmooth.global_reduction = lp.global_reduction;
  param_presmooth.precision = param_presmooth.precision_sloppy = param_presmooth.precision_precondition = PREC_SINGLE;
  param_presmooth.compute_true_res = false;
  param_presmooth.verbosity = 0;
  if (last) {  // coarsest grid: GCR(20) to smoother_tol (multigrid.cpp:63-70)
    param_presmooth.inv_type = lp.smoother == INV_MR ? INV_GCR : lp.smoother;
    param_presmooth.Nkrylov = 20;
    param_presmooth.maxiter = 1000;
    param_presmooth.delta = 1e-8;
  }
  DiracMatrix ms(matSmooth);
  presmoother.reset(Solver::create(param_presmooth, ms, ms, ms));
  // the residual after pre-smoothing comes out of the smoother itself where that is possible (use_solver_residual, lib/multigrid.cpp:536-546)
  if (!last && !(getenv("QB_MG_SOLVER_RESIDUAL") && atoi(getenv("QB_MG_SOLVER_RESIDUAL")) == 0))
    if (MR *mr = dynamic_cast<MR *>(presmoother.get())) mr->keep_residual = true;
  if (!last) {
    param_postsmooth = param_presmooth;
    param_postsmooth.use_init_guess = true;
    param_postsmooth.maxiter = lp.nu_post;
    postsmoother.reset(Solver::create(param_postsmooth, ms, ms, ms));
  }
  r.reset(new_full(*matResidual));
  pc_parity = (matSmooth->matpc() == MATPC_EVEN_EVEN || matSmooth->matpc() == MATPC_EVEN_EVEN_ASYM) ? 0 : 1;

  if (!last) {
    // ---- near-null vectors ----
    if (B_in) B = std::move(*B_in);
    if ((int)B.size() < lp.nvec) {
      // multigrid.cpp:29-38: compute, or load from vec_infile_level_<l>; multigrid.cpp:776: save after computing
      if (mp.compute_null_vector) {
        generate_null_vectors();
        if (!mp.vec_outfile.empty()) save_vectors();
      } else if (!mp.vec_infile.empty()) {
        load_vectors();
      } else {
        QB_ERROR("MG level %d: %d null vectors needed but compute_null_vector is off and vec_infile is empty", level + 1, lp.nvec);
      }
    }
    std::vector<SpinorField *> Bp;
    for (int i = 0; i < lp.nvec; i++) Bp.push_back(B[i].get());

    // ---- transfer operator and Galerkin coarse operator ----
    int bs[4] = {lp.geo_bs[0], lp.geo_bs[1], lp.geo_bs[2], lp.geo_bs[3]};
    int fineX[4];
    if (level == 0) {
      const DiracTM *d = dynamic_cast<const DiracTM *>(matResidual);
      if (!d) QB_ERROR("MG: level-0 operator must be Wilson / twisted mass");
      for (int k = 0; k < 4; k++) fineX[k] = d->lat->geom.X[k];
    } else {
      const DiracCoarse *d = dynamic_cast<const DiracCoarse *>(matResidual);
      for (int k = 0; k < 4; k++) fineX[k] = d->op->geom.X[k];
    }
    transfer.reset(new Transfer(Bp, lp.nvec, bs, lp.spin_bs, fineX));
    for (int k = 0; k < 4; k++) mp.level[level].geo_bs[k] = bs[k];  // written back like multigrid.cpp:120-122
    const MGLevelParam &cp = mp.level[level + 1];

    // ---- coarse-level null vectors: restricted fine ones unless every level generates its own ----
    std::vector<std::unique_ptr<SpinorField>> Bc;
    if (level + 1 < mp.n_level - 1 && !mp.generate_all_levels) {
      if (cp.nvec > lp.nvec) QB_ERROR("MG: n_vec[%d] = %d > n_vec[%d] = %d requires generate_all_levels", level + 1, cp.nvec, level, lp.nvec);
      for (int i = 0; i < cp.nvec; i++) {
        Bc.emplace_back(transfer->new_coarse_field());
        transfer->R(*Bc.back(), *B[i]);
      }
    }
    // V holds everything the cycle needs; the near-null vectors (n_vec x 96 B/site on the fine grid) are only kept for verify()
    if (!mp.keep_null_vectors) B.clear();

    // coarsening of the even-odd preconditioned system (multigrid.cpp:145-155): coarse_grid_solution_type = MATPC with an even-odd smoother
    pc_coarsen = lp.coarse_pc && lp.smoother_pc;
    if (lp.coarse_pc && !lp.smoother_pc) QB_ERROR("MG level %d: coarse_grid_solution_type = QUDA_MATPC_SOLUTION needs smoother_solve_type = QUDA_DIRECT_PC_SOLVE", level + 1);
    if (pc_coarsen && (matSmooth->matpc() == MATPC_EVEN_EVEN_ASYM || matSmooth->matpc() == MATPC_ODD_ODD_ASYM))
      QB_ERROR("Unsupported coarsening of matpc = %d (the preconditioned coarsening needs the symmetric even-odd operator, lib/coarse_op.cuh:1317)", matSmooth->matpc());
    coarse_op.reset(new CoarseOperator());
    const double tc0 = now_s();
    if (pc_coarsen) log_msg(1, "MG level %d: coarsening the even-odd preconditioned operator (bi-directional links of S^-1 M)\n", level + 1);
    matResidual->create_coarse_op(*coarse_op, *transfer, pc_coarsen);
    if (cp.smoother_pc) {
      coarse_op->compute_xinv();
      // Yhat = Xinv Y (createYpreconditioned, lib/coarse_op.cuh:1217-1283): the even-odd coarse operator 1 - Yhat_pq Yhat_qp in two launches
      // instead of four (QB_MG_YHAT=0: keep the Xinv launches, saves one copy of the links)
      if (!(getenv("QB_MG_YHAT") && atoi(getenv("QB_MG_YHAT")) == 0)) coarse_op->compute_yhat();
    }
    if (mp.half_storage) { transfer->enable_half_v(); coarse_op->enable_half_links(); }
    QB_CUDA(cudaStreamSynchronize(rt().compute));
    log_msg(1, "MG level %d: coarse operator %d x %d x %d x %d, N = %d built in %.3f s\n", level + 1, coarse_op->geom.X[0], coarse_op->geom.X[1],
            coarse_op->geom.X[2], coarse_op->geom.X[3], coarse_op->N, now_s() - tc0);
    coarseResidual.reset(new DiracCoarse(coarse_op, false, MATPC_EVEN_EVEN));
    // the even-odd systems of all levels live on the parity of the outer matpc_type (multigrid.cpp:300-309 sets the transfer's parity from it)
    coarseSmooth.reset(new DiracCoarse(coarse_op, cp.smoother_pc, pc_parity == 0 ? MATPC_EVEN_EVEN : MATPC_ODD_ODD));
    r_coarse.reset(transfer->new_coarse_field());
    x_coarse.reset(transfer->new_coarse_field());
    coarse.reset(new MG(mp, level + 1, coarseResidual.get(), coarseSmooth.get(), Bc.empty() ? nullptr : &Bc));

    // ---- coarse solver: the next level's cycle, wrapped in GCR(10) for a K-cycle (multigrid.cpp:225-275) ----
    if (lp.recursive && level + 1 < mp.n_level - 1) {
      param_coarse_solver.inv_type = INV_GCR;
      param_coarse_solver.inv_type_precondition = INV_MG;
      param_coarse_solver.is_preconditioner = true;
      param_coarse_solver.preserve_source = true;
      param_coarse_solver.use_init_guess = false;
      param_coarse_solver.maxiter = 11;
      param_coarse_solver.Nkrylov = 10;
      param_coarse_solver.tol = cp.smoother_tol;
      param_coarse_solver.global_reduction = true;
      param_coarse_solver.compute_true_res = false;
      param_coarse_solver.delta = 1e-8;
      param_coarse_solver.precision = param_coarse_solver.precision_sloppy = param_coarse_solver.precision_precondition = PREC_SINGLE;
      param_coarse_solver.verbosity = 0;
      if (cp.coarse_pc && cp.smoother_pc) {
        // the next level injects single-parity fields into its coarse grid: its K-cycle GCR runs on the even-odd operator and the
        // cycle below it sees single-parity fields (multigrid.cpp:250-261)
        DiracMatrix mc(coarseSmooth.get());
        coarse_solver_pc.reset(new PreconditionedSolver(new GCR(mc, mc, mc, param_coarse_solver, coarse.get()), coarseSmooth.get(), param_coarse_solver));
      } else {
        DiracMatrix mc(coarseResidual.get());
        coarse_solver_gcr.reset(new GCR(mc, mc, mc, param_coarse_solver, coarse.get()));
      }
    }
  }
  setup_secs = now_s() - t0;
  log_msg(1, "MG level %d: setup completed in %.3f s\n", level + 1, setup_secs);
}

// BiCGStab on the smoother operator with zero source and random initial guess, then global Gram-Schmidt
// (multigrid.cpp:693-779)
void MG::generate_null_vectors() {
  const MGLevelParam &lp = mp.level[level];
  log_msg(1, "MG level %d: generating %d null vectors (BiCGStab, maxiter %d, tol %g)\n", level + 1, lp.nvec, mp.setup_maxiter, mp.setup_tol);
  SolverParam sp;
  sp.inv_type = INV_BICGSTAB;
  sp.maxiter = mp.setup_maxiter;
  sp.tol = mp.setup_tol;
  sp.use_init_guess = true;
  sp.compute_null_vector = true;
  sp.precision = sp.precision_sloppy = sp.precision_precondition = PREC_SINGLE;
  sp.verbosity = mp.verbosity >= 3 ? 3 : 0;
  DiracMatrix ms(matSmooth);
  std::unique_ptr<SpinorField> b(new_full(*matResidual));
  auto seed_of = [&](int i) { return 0x5EEDull + 7919ull * (unsigned long long)(level * 1000 + i) + 104729ull * (unsigned long long)rt().rank; };
  auto orthonormalise_and_keep = [&](std::unique_ptr<SpinorField> &x) {
    const int i = (int)B.size();
    for (int j = 0; j < i; j++) {
      const Complex a = blas::cDotProduct(*B[j], *x);
      blas::caxpy(-a, *B[j], *x);
    }
    const double n2 = blas::norm2(*x);
    if (!(n2 > 1e-16)) QB_ERROR("Cannot orthogonalize %d vector", i);
    blas::ax(1.0 / sqrt(n2), *x);
    B.push_back(std::move(x));
  };
  if (level >= 1 && B.empty() && block_null_vectors_supported(matSmooth, lp.nvec)) {
    // coarse level: all solves advance together on block fields, the operator runs on the tensor cores (block_solver.cu)
    std::vector<std::unique_ptr<SpinorField>> xs(lp.nvec);
    std::vector<SpinorField *> px(lp.nvec);
    for (int i = 0; i < lp.nvec; i++) {
      xs[i].reset(new_full(*matResidual));
      random_fill(*xs[i], seed_of(i));
      px[i] = xs[i].get();
    }
    const int it = block_null_vectors(matSmooth, px, mp.setup_maxiter, mp.setup_tol);
    log_msg(1, "MG level %d: %d null vectors from one batched BiCGStab on the multi-RHS tensor-core operator (%d iterations)\n", level + 1, lp.nvec, it);
    for (int i = 0; i < lp.nvec; i++) orthonormalise_and_keep(xs[i]);
    return;
  }
  while ((int)B.size() < lp.nvec) {
    const int i = (int)B.size();
    std::unique_ptr<SpinorField> x(new_full(*matResidual));
    random_fill(*x, seed_of(i));
    blas::zero(*b);
    BiCGStab solve(ms, ms, sp);
    SpinorField src, sol;
    matSmooth->prepare(src, sol, *x, *b, SOL_MAT);
    solve(sol, src);
    matSmooth->reconstruct(*x, *b, SOL_MAT);
    for (int j = 0; j < i; j++) {
      const Complex a = blas::cDotProduct(*B[j], *x);
      blas::caxpy(-a, *B[j], *x);
    }
    const double n2 = blas::norm2(*x);
    if (!(n2 > 1e-16)) QB_ERROR("Cannot orthogonalize %d vector", i);
    blas::ax(1.0 / sqrt(n2), *x);
    B.push_back(std::move(x));
    log_msg(2, "MG level %d: null vector %d done (%d BiCGStab iterations so far)\n", level + 1, i, sp.iter);
  }
}

// Near-null vector files (multigrid.cpp:607-691 uses QIO's read / write_spinor_field; QIO / LIME are not in this image, so files
// written by the reference cannot be read here and vice versa -- stated in INTEGRATION.md and in the error message below).  The
// container is ONE file per level, <name>_level_<l>, independent of the rank layout: a setup saved on N ranks loads on M ranks.
//   bytes 0..127  header, little-endian: char magic[8] = "QB200VEC"; int32 version = 2, level, nvec, nspin, ncolor, global X[4], prec (4)
//   then nvec vectors, each the GLOBAL field in lexicographic site order (x fastest, then y, z, t; NOT even-odd),
//   per site [spin][colour][re, im] float32, DeGrand-Rossi basis on the fine level (coarse levels: spin = chirality, colour = vector index)
// Every rank writes / reads the runs of sites it owns at their global offsets (pwrite / pread).
namespace {
constexpr int VEC_HEADER_BYTES = 128;
struct VecFileHeader {
  char magic[8];          // "QB200VEC"
  int version, level, nvec, nspin, ncolor, X[4], prec;
};
static_assert(sizeof(VecFileHeader) <= VEC_HEADER_BYTES, "header layout");
std::string vec_file_name(const std::string &base, int level) { return base + "_level_" + std::to_string(level); }
void level_dims(const Dirac *m, int *X) {
  if (const DiracTM *d = dynamic_cast<const DiracTM *>(m)) for (int k = 0; k < 4; k++) X[k] = d->lat->geom.X[k];
  else if (const DiracCoarse *c = dynamic_cast<const DiracCoarse *>(m)) for (int k = 0; k < 4; k++) X[k] = c->op->geom.X[k];
  else QB_ERROR("MG: unknown operator type");
}
// local host order [parity][x_cb][component] <-> local lexicographic order (x fastest)
void to_lex(std::vector<float> &lex, std::vector<float> &eo, const int *X, int ncomp2, bool back) {
  const long V = (long)X[0] * X[1] * X[2] * X[3], Vh = V / 2;
  for (long i = 0; i < V; i++) {
    const int x = (int)(i % X[0]), y = (int)((i / X[0]) % X[1]), z = (int)((i / ((long)X[0] * X[1])) % X[2]), t = (int)(i / ((long)X[0] * X[1] * X[2]));
    const long e = (long)((x + y + z + t) & 1) * Vh + (i >> 1);
    if (back) memcpy(eo.data() + e * ncomp2, lex.data() + i * ncomp2, sizeof(float) * ncomp2);
    else memcpy(lex.data() + i * ncomp2, eo.data() + e * ncomp2, sizeof(float) * ncomp2);
  }
}
// calls f(local lexicographic start site, run length in sites, global lexicographic start site) for every run of this rank's sites that
// is contiguous in the global file
template <typename F> void for_each_run(const int *X, F f) {
  const Runtime &r = rt();
  int G[4], off[4];
  for (int d = 0; d < 4; d++) { G[d] = X[d] * r.grid[d]; off[d] = r.coord[d] * X[d]; }
  // a run covers the dimensions 0..m-1 entirely, where all of grid[0..m-2] are 1 (the x extent of a rank is always contiguous)
  int m = 1;
  while (m < 4 && r.grid[m - 1] == 1) m++;
  long run = 1;
  for (int d = 0; d < m; d++) run *= X[d];
  long outer = 1;
  for (int d = m; d < 4; d++) outer *= X[d];
  for (long o = 0; o < outer; o++) {
    int c[4] = {0, 0, 0, 0};
    long rem = o;
    for (int d = m; d < 4; d++) { c[d] = (int)(rem % X[d]); rem /= X[d]; }
    const long lstart = (((long)c[3] * X[2] + c[2]) * X[1] + c[1]) * X[0] + c[0];
    const long gstart = ((((long)(c[3] + off[3]) * G[2] + (c[2] + off[2])) * G[1] + (c[1] + off[1])) * G[0]) + (c[0] + off[0]);
    f(lstart, run, gstart);
  }
}
}  // namespace

void MG::save_vectors() {
  const std::string name = vec_file_name(mp.vec_outfile, level);
  log_msg(1, "MG level %d: saving %d vectors to %s\n", level + 1, (int)B.size(), name.c_str());
  Runtime &r = rt();
  VecFileHeader h{};
  memcpy(h.magic, "QB200VEC", 8);
  h.version = 2; h.level = level; h.nvec = (int)B.size(); h.nspin = B[0]->nspin; h.ncolor = B[0]->ncolor; h.prec = 4;
  int X[4];
  level_dims(matResidual, X);
  long Vg = 1;
  for (int d = 0; d < 4; d++) { h.X[d] = X[d] * r.grid[d]; Vg *= h.X[d]; }
  const int nc2 = B[0]->ncomplex * 2;
  if (r.rank == 0) {
    FILE *f = fopen(name.c_str(), "wb");
    if (!f) QB_ERROR("cannot open %s for writing", name.c_str());
    char head[VEC_HEADER_BYTES] = {0};
    memcpy(head, &h, sizeof(h));
    if (fwrite(head, 1, VEC_HEADER_BYTES, f) != VEC_HEADER_BYTES) QB_ERROR("write error on %s", name.c_str());
    fclose(f);
  }
  comm_barrier();
  const int fd = open(name.c_str(), O_WRONLY);
  if (fd < 0) QB_ERROR("cannot open %s for writing", name.c_str());
  const long V = (long)X[0] * X[1] * X[2] * X[3];
  std::vector<float> host((size_t)V * nc2), lex((size_t)V * nc2);
  for (size_t k = 0; k < B.size(); k++) {
    export_generic(host.data(), *B[k], r.compute);
    to_lex(lex, host, X, nc2, false);
    for_each_run(X, [&](long ls, long n, long gs) {
      const off_t at = VEC_HEADER_BYTES + ((off_t)k * Vg + gs) * nc2 * (off_t)sizeof(float);
      const size_t bytes = (size_t)n * nc2 * sizeof(float);
      if (pwrite(fd, lex.data() + ls * nc2, bytes, at) != (ssize_t)bytes) QB_ERROR("write error on %s", name.c_str());
    });
  }
  close(fd);
  comm_barrier();
}

void MG::load_vectors() {
  const MGLevelParam &lp = mp.level[level];
  const std::string name = vec_file_name(mp.vec_infile, level);
  log_msg(1, "MG level %d: loading %d vectors from %s\n", level + 1, lp.nvec, name.c_str());
  Runtime &r = rt();
  const int fd = open(name.c_str(), O_RDONLY);
  if (fd < 0) QB_ERROR("cannot open %s for reading", name.c_str());
  char head[VEC_HEADER_BYTES];
  VecFileHeader h{};
  if (pread(fd, head, VEC_HEADER_BYTES, 0) != VEC_HEADER_BYTES) QB_ERROR("%s is truncated", name.c_str());
  memcpy(&h, head, sizeof(h));
  if (memcmp(h.magic, "QB200VEC", 8) != 0 || h.version != 2)
    QB_ERROR("%s is not a near-null vector file of this library (version 2 container, see INTEGRATION.md; the reference's QIO / LIME files are not readable here)", name.c_str());
  std::unique_ptr<SpinorField> like(new_full(*matResidual));
  int X[4];
  level_dims(matResidual, X);
  long Vg = 1;
  bool dims_ok = true;
  for (int d = 0; d < 4; d++) { dims_ok = dims_ok && h.X[d] == X[d] * r.grid[d]; Vg *= h.X[d]; }
  if (h.level != level || h.nvec < lp.nvec || h.nspin != like->nspin || h.ncolor != like->ncolor || !dims_ok)
    QB_ERROR("%s does not match this level (file: level %d, %d vectors, %d x %d components, global lattice %d x %d x %d x %d)", name.c_str(), h.level, h.nvec,
             h.nspin, h.ncolor, h.X[0], h.X[1], h.X[2], h.X[3]);
  const int nc2 = like->ncomplex * 2;
  const long V = (long)X[0] * X[1] * X[2] * X[3];
  std::vector<float> host((size_t)V * nc2), lex((size_t)V * nc2);
  B.clear();
  for (int k = 0; k < lp.nvec; k++) {
    for_each_run(X, [&](long ls, long n, long gs) {
      const off_t at = VEC_HEADER_BYTES + ((off_t)k * Vg + gs) * nc2 * (off_t)sizeof(float);
      const size_t bytes = (size_t)n * nc2 * sizeof(float);
      if (pread(fd, lex.data() + ls * nc2, bytes, at) != (ssize_t)bytes) QB_ERROR("%s is truncated", name.c_str());
    });
    to_lex(lex, host, X, nc2, true);
    std::unique_ptr<SpinorField> v(new_full(*matResidual));
    import_generic(*v, host.data(), r.compute);
    QB_CUDA(cudaStreamSynchronize(r.compute));
    B.push_back(std::move(v));
  }
  close(fd);
}

// x <- smoother applied to M x = b through the (possibly even-odd preconditioned) smoother operator
void MG::smooth(Solver &s, SpinorField &x, SpinorField &b) {
  SpinorField src, sol;
  matSmooth->prepare(src, sol, x, b, SOL_MAT);
  s(sol, src);
  matSmooth->reconstruct(x, b, SOL_MAT);
}

// optional wall-clock profile of the cycle (QUDA_B200_MG_PROFILE=1): sections are bracketed by stream syncs
static int g_mg_profile = -1;
void mg_profile_enable(bool on) { g_mg_profile = on ? 1 : 0; }
static bool mg_profile_on() {
  if (g_mg_profile < 0) { const char *e = getenv("QUDA_B200_MG_PROFILE"); g_mg_profile = (e && e[0] == '1') ? 1 : 0; }
  return g_mg_profile == 1;
}
struct Section {
  double *acc; double t0; bool on;
  Section(double *a) : acc(a), on(mg_profile_on()) { if (on) { cudaStreamSynchronize(rt().compute); t0 = now_s(); } }
  ~Section() { if (on) { cudaStreamSynchronize(rt().compute); *acc += now_s() - t0; } }
};

// The cycle on the even-odd system M_pc x = b of this level (single-parity fields; multigrid.cpp:494-560 with outer = inner =
// QUDA_MATPC_SOLUTION).  M_pc is the Schur complement of S^-1 M, whose Galerkin product is the next level's operator: the residual
// is restricted from this parity only, the coarse level solves its FULL system, and the correction is prolongated back to this parity.
void MG::cycle_pc(SpinorField &x, SpinorField &b) {
  const MGLevelParam &lp = mp.level[level];
  ncycle++;
  if (!matSmooth->is_pc()) QB_ERROR("MG level %d: single-parity fields need an even-odd preconditioned smoother operator", level + 1);
  if (level == mp.n_level - 1) {  // coarsest grid: solve the even-odd system directly
    Section s(&t_prof[0]);
    (*presmoother)(x, b);
    return;
  }
  if (!pc_coarsen) QB_ERROR("Unsupported solution type combination: single-parity fields on MG level %d need coarse_grid_solution_type = QUDA_MATPC_SOLUTION", level + 1);
  SpinorField rp;
  r->view_parity(rp, pc_parity);
  const SpinorField *rsrc = &rp;
  if (lp.nu_pre > 0) {
    { Section s(&t_prof[0]); (*presmoother)(x, b); }
    Section s(&t_prof[1]);
    const MR *mr = dynamic_cast<const MR *>(presmoother.get());
    if (mr && mr->residual()) rsrc = mr->residual();   // b - M_pc x left behind by the smoother's last step
    else {
      matSmooth->M(rp, x);
      blas::axpby(1.0, b, -1.0, rp);
    }
  } else {
    blas::zero(x);
    blas::copy(rp, b);
  }
  { Section s(&t_prof[2]); transfer->R(*r_coarse, *rsrc, pc_parity); }
  {
    Section s(&t_prof[3]);
    if (coarse_solver_pc) (*coarse_solver_pc)(*x_coarse, *r_coarse);
    else if (coarse_solver_gcr) (*coarse_solver_gcr)(*x_coarse, *r_coarse);
    else (*coarse)(*x_coarse, *r_coarse);
  }
  {
    Section s(&t_prof[4]);
    SpinorField *fo = &x;
    const SpinorField *ci = x_coarse.get();
    transfer->P_multi(&fo, &ci, 1, true, pc_parity);
  }
  if (lp.nu_post > 0) { Section s(&t_prof[5]); (*postsmoother)(x, b); }
}

void MG::cycle(SpinorField &x, SpinorField &b) {
  const MGLevelParam &lp = mp.level[level];
  if (x.nparity == 1) { cycle_pc(x, b); return; }
  if (level < mp.n_level - 1 && pc_coarsen) {
    // full fields on a level that coarsens its even-odd system (outer MAT, inner MATPC): reduce to the even-odd system exactly,
    // run the single-parity cycle on it, reconstruct the other parity
    SpinorField src, sol;
    matSmooth->prepare(src, sol, x, b, SOL_MAT);
    cycle_pc(sol, src);
    matSmooth->reconstruct(x, b, SOL_MAT);
    return;
  }
  ncycle++;
  if (level == mp.n_level - 1) {  // coarsest grid solve
    Section s(&t_prof[0]);
    smooth(*presmoother, x, b);
    return;
  }
  // pre-smoothing (zero initial guess) and residual
  int rpar = -1;   // >= 0: the residual lives on that parity only
  if (lp.nu_pre > 0) {
    { Section s(&t_prof[0]); smooth(*presmoother, x, b); }
    Section s(&t_prof[1]);
    const MR *mr = dynamic_cast<const MR *>(presmoother.get());
    if (mr && mr->residual() && matSmooth->is_pc()) {
      // x came out of the even-odd system with x_q reconstructed exactly, so b - M x vanishes on parity q and equals S (src - M_pc x_p) on
      // parity p (S = the site-diagonal term; for the asymmetric fine operator no S): one site-local kernel on the smoother's own residual
      // instead of a full operator application, and the restrictor reads one parity of V only
      rpar = pc_parity;
      SpinorField rp;
      r->view_parity(rp, rpar);
      const bool symmetric = matSmooth->matpc() == MATPC_EVEN_EVEN || matSmooth->matpc() == MATPC_ODD_ODD;
      if (symmetric) matSmooth->Diag(rp, *mr->residual(), rpar);
      else blas::copy(rp, *mr->residual());
    } else {
      matResidual->M(*r, x);
      blas::axpby(1.0, b, -1.0, *r);
    }
  } else {
    blas::zero(x);
    blas::copy(*r, b);
  }
  // coarse-grid correction
  { Section s(&t_prof[2]); transfer->R(*r_coarse, *r, rpar); }
  {
    Section s(&t_prof[3]);
    if (coarse_solver_pc) (*coarse_solver_pc)(*x_coarse, *r_coarse);
    else if (coarse_solver_gcr) (*coarse_solver_gcr)(*x_coarse, *r_coarse);
    else (*coarse)(*x_coarse, *r_coarse);
  }
  {  // x += P x_coarse in one pass (the prolongator adds into x: no temporary, no separate xpy)
    Section s(&t_prof[4]);
    SpinorField *fo = &x;
    const SpinorField *ci = x_coarse.get();
    transfer->P_multi(&fo, &ci, 1, true);
  }
  // post-smoothing with x as the initial guess
  if (lp.nu_post > 0) { Section s(&t_prof[5]); smooth(*postsmoother, x, b); }
}

void MG::print_profile() {
  if (!mg_profile_on()) return;
  log_msg(0, "MG level %d profile: %ld cycles, smooth-pre %.4f s, residual %.4f s, restrict %.4f s, coarse-solve %.4f s, prolong %.4f s, smooth-post %.4f s\n",
          level + 1, ncycle, t_prof[0], t_prof[1], t_prof[2], t_prof[3], t_prof[4], t_prof[5]);
  for (double &t : t_prof) t = 0;
  ncycle = 0;
  if (coarse) coarse->print_profile();
}

void MG::operator()(SpinorField &x, SpinorField &b) {
  if (x.nparity != b.nparity) QB_ERROR("MG: solution and source must both be full or both single-parity fields");
  if (x32 && x32->nparity != x.nparity) { x32.reset(); b32.reset(); }
  if (x.prec == PREC_SINGLE && b.prec == PREC_SINGLE) {
    cycle(x, b);  // b is only read (smoothers preserve their source)
    return;
  }
  // outer solver runs in another precision: the cycle itself is single precision (reference: cuda_prec_sloppy)
  if (!x32) { x32.reset(new_like(x, PREC_SINGLE)); b32.reset(new_like(b, PREC_SINGLE)); }
  blas::copy(*b32, b);
  cycle(*x32, *b32);
  blas::copy(x, *x32);
}

long long MG::flops() const { return 0; }

// MG::verify identities (multigrid.cpp:372-486)
void MG::verify(double *dev) {
  dev[0] = dev[1] = dev[2] = 0.0;
  if (!transfer) return;
  std::unique_ptr<SpinorField> eta(transfer->new_coarse_field()), t_c(transfer->new_coarse_field()), t_c2(transfer->new_coarse_field());
  std::unique_ptr<SpinorField> t_f(new_full(*matResidual)), t_f2(new_full(*matResidual));
  random_fill(*eta, 424242ull + level + 104729ull * (unsigned long long)rt().rank);
  // (1) R P eta = eta
  transfer->P(*t_f, *eta);
  transfer->R(*t_c, *t_f);
  dev[0] = sqrt(blas::xmyNorm(*eta, *t_c) / blas::norm2(*eta));
  // (2) P R v_k = v_k for the null vectors
  if (B.empty()) dev[1] = -1.0;  // null vectors were released (run_verify off)
  for (size_t k = 0; k < B.size() && (int)k < mp.level[level].nvec; k++) {
    transfer->R(*t_c, *B[k]);
    transfer->P(*t_f, *t_c);
    const double d = sqrt(blas::xmyNorm(*B[k], *t_f) / blas::norm2(*B[k]));
    if (d > dev[1]) dev[1] = d;
  }
  // (3) R M P eta = M_c eta
  transfer->P(*t_f, *eta);
  matResidual->M(*t_f2, *t_f);
  if (pc_coarsen) {  // the coarse operator is the Galerkin product of S^-1 M
    matResidual->DiagInv(*t_f, *t_f2);
    transfer->R(*t_c, *t_f);
  } else {
    transfer->R(*t_c, *t_f2);
  }
  coarseResidual->M(*t_c2, *eta);
  dev[2] = sqrt(blas::xmyNorm(*t_c2, *t_c) / blas::norm2(*t_c2));
}

}  // namespace qb
