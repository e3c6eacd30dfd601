// Drives libquda_b200.so through the `namespace quda` C++ facade (include/quda_cpp.h) with the call sequence of the QKXTM code in
// the reference library (lib/interface_quda.cpp:6285-6500: createDirac -> ColorSpinorField::Create -> cudaColorSpinorField ->
// DiracM -> prepare -> Solver::create -> solve -> reconstruct -> copy back).  Inputs (gauge, source) come from files written by
// tests/test_cpp_facade.py; the solution goes back to a file and is checked there against the oracle's host operator.
// Compiled with plain g++ (no CUDA headers): g++ -std=c++11 -I include facade_solve.cpp -L... -lquda_b200
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "quda_cpp.h"

using namespace quda;

static std::vector<double> read_file(const char *name, size_t n) {
  std::vector<double> v(n);
  FILE *f = fopen(name, "rb");
  if (!f || fread(v.data(), sizeof(double), n, f) != n) { fprintf(stderr, "cannot read %s\n", name); exit(2); }
  fclose(f);
  return v;
}

int main(int argc, char **argv) {
  if (argc < 11) { fprintf(stderr, "usage: facade_solve X Y Z T kappa mu pc(0|1) gauge.bin source.bin solution.bin\n"); return 2; }
  int X[4] = {atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4])};
  const double kappa = atof(argv[5]), mu = atof(argv[6]);
  const bool pc_solve = atoi(argv[7]) != 0;
  const size_t V = (size_t)X[0] * X[1] * X[2] * X[3];
  std::vector<double> gauge = read_file(argv[8], 4 * V * 18), src = read_file(argv[9], V * 24), sol(V * 24, 0.0);

  initQuda(0);
  QudaGaugeParam gp = newQudaGaugeParam();
  for (int d = 0; d < 4; d++) gp.X[d] = X[d];
  gp.anisotropy = 1.0; gp.type = QUDA_WILSON_LINKS; gp.gauge_order = QUDA_QDP_GAUGE_ORDER; gp.t_boundary = QUDA_PERIODIC_T;
  gp.cpu_prec = QUDA_DOUBLE_PRECISION; gp.cuda_prec = QUDA_DOUBLE_PRECISION; gp.reconstruct = QUDA_RECONSTRUCT_12;
  gp.cuda_prec_sloppy = QUDA_SINGLE_PRECISION; gp.reconstruct_sloppy = QUDA_RECONSTRUCT_12;
  gp.cuda_prec_precondition = QUDA_SINGLE_PRECISION; gp.reconstruct_precondition = QUDA_RECONSTRUCT_12;
  gp.gauge_fix = QUDA_GAUGE_FIXED_NO; gp.ga_pad = 0;
  void *links[4];
  for (int d = 0; d < 4; d++) links[d] = gauge.data() + (size_t)d * V * 18;
  loadGaugeQuda(links, &gp);

  QudaInvertParam inv = newQudaInvertParam();
  inv.kappa = kappa; inv.mu = mu; inv.epsilon = 0.0; inv.twist_flavor = QUDA_TWIST_PLUS; inv.dslash_type = QUDA_TWISTED_MASS_DSLASH;
  inv.matpc_type = QUDA_MATPC_EVEN_EVEN; inv.dagger = QUDA_DAG_NO;
  inv.cpu_prec = QUDA_DOUBLE_PRECISION; inv.cuda_prec = QUDA_DOUBLE_PRECISION; inv.cuda_prec_sloppy = QUDA_SINGLE_PRECISION;
  inv.cuda_prec_precondition = QUDA_SINGLE_PRECISION;
  inv.solution_type = QUDA_MAT_SOLUTION; inv.solve_type = pc_solve ? QUDA_DIRECT_PC_SOLVE : QUDA_DIRECT_SOLVE;
  inv.mass_normalization = QUDA_KAPPA_NORMALIZATION; inv.gamma_basis = QUDA_DEGRAND_ROSSI_GAMMA_BASIS; inv.dirac_order = QUDA_DIRAC_ORDER;
  inv.input_location = QUDA_CPU_FIELD_LOCATION; inv.output_location = QUDA_CPU_FIELD_LOCATION;
  inv.tune = QUDA_TUNE_NO; inv.sp_pad = 0; inv.cl_pad = 0; inv.verbosity = QUDA_SILENT; inv.Ls = 1; inv.mass = 0.0; inv.m5 = 0.0;
  inv.inv_type = QUDA_GCR_INVERTER; inv.inv_type_precondition = QUDA_INVALID_INVERTER; inv.preserve_source = QUDA_PRESERVE_SOURCE_YES;
  inv.gcrNkrylov = 20; inv.tol = 1e-9; inv.maxiter = 4000; inv.reliable_delta = 1e-4; inv.use_init_guess = QUDA_USE_INIT_GUESS_NO;

  // ---- from here on: the reference's internal C++ API ----
  TimeProfile profile("facade_solve");
  Dirac *d = NULL, *dSloppy = NULL, *dPre = NULL;
  createDirac(d, dSloppy, dPre, inv, pc_solve);
  Dirac &dirac = *d;

  const bool pc_solution = false;
  ColorSpinorParam cpuParam(src.data(), inv, X, pc_solution, QUDA_CPU_FIELD_LOCATION);
  ColorSpinorField *h_b = ColorSpinorField::Create(cpuParam);
  cpuParam.v = sol.data();
  ColorSpinorField *h_x = ColorSpinorField::Create(cpuParam);
  ColorSpinorParam cudaParam(cpuParam, inv);
  ColorSpinorField *b = new cudaColorSpinorField(*h_b, cudaParam);
  ColorSpinorField *x = new cudaColorSpinorField(cudaParam);
  ColorSpinorField *in = NULL, *out = NULL;

  const double nb = blas::norm2(*b);
  double nh = 0.0;
  for (size_t i = 0; i < src.size(); i++) nh += src[i] * src[i];
  printf("FACADE norm2 device %.15e host %.15e\n", nb, nh);
  massRescale(*static_cast<cudaColorSpinorField *>(b), inv);

  DiracM m(dirac), mSloppy(*dSloppy), mPre(*dPre);
  dirac.prepare(in, out, *x, *b, inv.solution_type);
  SolverParam solverParam(inv);
  Solver *solve = Solver::create(solverParam, m, mSloppy, mPre, profile);
  (*solve)(*out, *in);
  solverParam.updateInvertParam(inv);
  dirac.reconstruct(*x, *b, inv.solution_type);

  // residual through the facade's own operator: |b - M x| / |b| with the full (unpreconditioned) operator
  Dirac *dfull = NULL, *t1 = NULL, *t2 = NULL;
  createDirac(dfull, t1, t2, inv, false);
  ColorSpinorField *r = new cudaColorSpinorField(cudaParam);
  dfull->M(*r, *x);
  const double r2 = blas::xmyNorm(*b, *r);
  printf("FACADE iter %d true_res %.6e facade_res %.6e\n", inv.iter, inv.true_res, sqrt(r2 / blas::norm2(*b)));

  *h_x = *x;
  FILE *f = fopen(argv[10], "wb");
  fwrite(sol.data(), sizeof(double), sol.size(), f);
  fclose(f);

  delete solve; delete r; delete x; delete b; delete h_x; delete h_b;
  delete d; delete dSloppy; delete dPre; delete dfull; delete t1; delete t2;
  endQuda();
  return 0;
}
