// C-ABI entry points: the quda.h subset of SURVEY.md section 8(b) plus the resident-field extensions.
// Behavioural model: /root/reference/lib/interface_quda.cpp (line refs at each function).
#include <cfloat>
#include <climits>
#include <chrono>
#include <map>
#include <vector>
#include "../../include/quda.h"
#include "blas.h"
#include "comm.h"
#include "dirac.h"

using namespace qb;

#define INVALID_INT QUDA_INVALID_ENUM
#define INVALID_DOUBLE DBL_MIN

namespace {

struct GaugeSet {
  Lattice lat;
  QudaGaugeParam param;
  std::shared_ptr<GaugeField> precise, sloppy, precondition;
  bool loaded = false;
} G;

void require_init() {
  if (!rt().memory_ready) QB_ERROR("QUDA not initialized (call initQuda first)");
}
void require_gauge() {
  require_init();
  if (!G.loaded) QB_ERROR("Gauge field not allocated");
}

Prec to_prec(QudaPrecision p, const char *what) {
  if (p == QUDA_DOUBLE_PRECISION) return PREC_DOUBLE;
  if (p == QUDA_SINGLE_PRECISION) return PREC_SINGLE;
  if (p == QUDA_HALF_PRECISION) return PREC_HALF;
  QB_ERROR("Parameter %s undefined", what);
}

HostBasis to_basis(QudaGammaBasis b) {
  if (b == QUDA_DEGRAND_ROSSI_GAMMA_BASIS || b == QUDA_CHIRAL_GAMMA_BASIS) return BASIS_DEGRAND_ROSSI;
  if (b == QUDA_UKQCD_GAMMA_BASIS) return BASIS_UKQCD;
  QB_ERROR("Parameter gamma_basis undefined");
}
HostSpinorOrder to_order(QudaDiracFieldOrder o) {
  if (o == QUDA_DIRAC_ORDER) return ORDER_SPIN_COLOR;
  if (o == QUDA_QDP_DIRAC_ORDER) return ORDER_COLOR_SPIN;
  QB_ERROR("Dirac order %d not supported", (int)o);
}

void check_invert_param_operator(const QudaInvertParam *p) {
  if (p->dslash_type != QUDA_TWISTED_MASS_DSLASH && p->dslash_type != QUDA_WILSON_DSLASH)
    QB_ERROR("Unsupported dslash_type %d (this build covers Wilson and degenerate twisted mass)", (int)p->dslash_type);
  if (p->kappa == INVALID_DOUBLE) QB_ERROR("Parameter kappa undefined");
  if (p->dslash_type == QUDA_TWISTED_MASS_DSLASH) {
    if (p->mu == INVALID_DOUBLE) QB_ERROR("Parameter mu undefined");
    if (p->twist_flavor != QUDA_TWIST_PLUS && p->twist_flavor != QUDA_TWIST_MINUS)
      QB_ERROR("Twist flavor not set %d (only the degenerate +-1 flavours are supported)", (int)p->twist_flavor);
  }
  if (p->matpc_type == QUDA_MATPC_INVALID) QB_ERROR("Parameter matpc_type undefined");
  if (p->dagger == QUDA_DAG_INVALID) QB_ERROR("Parameter dagger undefined");
}

const GaugeField *pick_gauge(Prec prec) {
  if (G.precise && G.precise->prec == prec) return G.precise.get();
  if (G.sloppy && G.sloppy->prec == prec) return G.sloppy.get();
  if (G.precondition && G.precondition->prec == prec) return G.precondition.get();
  QB_ERROR("no resident gauge field in precision %d (load it via cuda_prec / cuda_prec_sloppy / cuda_prec_precondition)", (int)prec);
}

// setDiracParam + Dirac::create (interface_quda.cpp:1265-1340, :1386-1410)
DiracTM *make_dirac(const QudaInvertParam *p, bool pc, const GaugeField *gauge, double kappa_scale = 1.0, double mu_scale = 1.0) {
  check_invert_param_operator(p);
  const int flavor = p->dslash_type == QUDA_TWISTED_MASS_DSLASH ? (int)p->twist_flavor : 0;
  const double mu = p->dslash_type == QUDA_TWISTED_MASS_DSLASH ? p->mu * mu_scale : 0.0;
  return new DiracTM(&G.lat, gauge, p->kappa * kappa_scale, mu, flavor, pc, (int)p->matpc_type, p->dagger == QUDA_DAG_YES);
}

// small pool of resident work fields so that dslashQuda / MatQuda do not cudaMalloc per call
std::map<std::pair<int, int>, std::vector<SpinorField *>> pool;
SpinorField *pool_get(int nparity, Prec prec) {
  auto &v = pool[{nparity, (int)prec}];
  if (!v.empty() && v.back()->Vh == G.lat.geom.Vh) {
    SpinorField *f = v.back();
    v.pop_back();
    return f;
  }
  return new SpinorField(G.lat.geom.Vh, nparity, prec);
}
void pool_put(SpinorField *f) { pool[{f->nparity, (int)f->prec}].push_back(f); }
void pool_clear() {
  for (auto &kv : pool)
    for (auto *f : kv.second) delete f;
  pool.clear();
}

void mass_rescale_out(SpinorField &out, const QudaInvertParam *p, bool pc, bool normal) {
  // MatQuda / MatDagMatQuda output normalisation (interface_quda.cpp:1754-1766, :1832-1846)
  const double kappa = p->kappa;
  double s = 1.0;
  if (pc) {
    if (p->mass_normalization == QUDA_MASS_NORMALIZATION) s = normal ? 1.0 / pow(2.0 * kappa, 4) : 0.25 / (kappa * kappa);
    else if (p->mass_normalization == QUDA_ASYMMETRIC_MASS_NORMALIZATION) s = normal ? 0.25 / (kappa * kappa) : 0.5 / kappa;
  } else if (p->mass_normalization == QUDA_MASS_NORMALIZATION || p->mass_normalization == QUDA_ASYMMETRIC_MASS_NORMALIZATION) {
    s = normal ? 0.25 / (kappa * kappa) : 0.5 / kappa;
  }
  if (s != 1.0) blas::ax(s, out);
}

}  // namespace

// accessors used by solver / multigrid glue
namespace qb {
Lattice &global_lattice() { return G.lat; }
const GaugeField *global_gauge(Prec prec) { return pick_gauge(prec); }
}

extern "C" {

// ---- verbosity / init -----------------------------------------------------------------------
void setVerbosityQuda(QudaVerbosity verbosity, const char prefix[], FILE *outfile) {
  Runtime &r = rt();
  r.verbosity = (int)verbosity;
  if (prefix) { strncpy(r.prefix, prefix, sizeof(r.prefix) - 1); r.prefix[sizeof(r.prefix) - 1] = 0; }
  r.out = outfile;
}

void initCommsGridQuda(int nDim, const int *dims, QudaCommsMap func, void *fdata) {
  if (nDim != 4) QB_ERROR("Number of communication grid dimensions must be 4");
  comm_set_grid(dims, func, fdata);
}

void initQudaDevice(int dev) {
  Runtime &r = rt();
  if (r.device_ready) return;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) QB_ERROR("No CUDA devices found (%s): this library has no CPU fallback", cudaGetErrorString(e));
  if (dev < 0) dev = r.rank % ndev;  // one process per GPU
  if (dev >= ndev) QB_ERROR("Device %d does not exist (%d visible)", dev, ndev);
  cudaDeviceProp prop;
  QB_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major < 10) log_msg(1, "WARNING: device %s is sm_%d%d; kernels are built for sm_100a only\n", prop.name, prop.major, prop.minor);
  QB_CUDA(cudaSetDevice(dev));
  r.device = dev;
  r.num_sms = prop.multiProcessorCount;
  r.device_ready = true;
  log_msg(2, "Using device %d: %s (%d SMs)\n", dev, prop.name, r.num_sms);
}

void initQudaMemory(void) {
  Runtime &r = rt();
  if (!r.device_ready) QB_ERROR("initQudaDevice must be called before initQudaMemory");
  if (r.memory_ready) return;
  int lo, hi;
  QB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  QB_CUDA(cudaStreamCreateWithPriority(&r.compute, cudaStreamNonBlocking, lo));
  QB_CUDA(cudaStreamCreateWithPriority(&r.halo, cudaStreamNonBlocking, hi));  // pack + exchange pre-empt the interior kernel
  QB_CUDA(cudaEventCreateWithFlags(&r.ev_pack_ready, cudaEventDisableTiming));
  QB_CUDA(cudaEventCreateWithFlags(&r.ev_halo_done, cudaEventDisableTiming));
  QB_CUDA(cudaEventCreateWithFlags(&r.ev_in_ready, cudaEventDisableTiming));
  blas::init();
  r.memory_ready = true;
}

void initQuda(int dev) {
  initQudaDevice(dev);
  initQudaMemory();
}

void freeGaugeQuda(void);
void free_staging_buffers();

void endQuda(void) {
  Runtime &r = rt();
  if (!r.device_ready) return;
  if (r.memory_ready) {
    QB_CUDA(cudaDeviceSynchronize());
    freeGaugeQuda();
    blas::end();
    free_staging_buffers();
    comm_finalize();
    cudaEventDestroy(r.ev_pack_ready); cudaEventDestroy(r.ev_halo_done); cudaEventDestroy(r.ev_in_ready);
    cudaStreamDestroy(r.compute); cudaStreamDestroy(r.halo);
    r.compute = r.halo = nullptr;
  }
  r.memory_ready = false;
  r.device_ready = false;
}

// ---- parameter structs (lib/check_params.h) ---------------------------------------------------
QudaGaugeParam newQudaGaugeParam(void) {
  QudaGaugeParam p;
  memset(&p, 0, sizeof(p));
  p.location = QUDA_CPU_FIELD_LOCATION;
  for (int i = 0; i < 4; i++) p.X[i] = INVALID_INT;
  p.anisotropy = p.tadpole_coeff = p.scale = INVALID_DOUBLE;
  p.type = QUDA_INVALID_LINKS;
  p.gauge_order = QUDA_INVALID_GAUGE_ORDER;
  p.t_boundary = QUDA_INVALID_T_BOUNDARY;
  p.cpu_prec = p.cuda_prec = p.cuda_prec_sloppy = p.cuda_prec_precondition = QUDA_INVALID_PRECISION;
  p.reconstruct = p.reconstruct_sloppy = p.reconstruct_precondition = QUDA_RECONSTRUCT_INVALID;
  p.gauge_fix = QUDA_GAUGE_FIXED_INVALID;
  p.ga_pad = INVALID_INT;
  p.site_ga_pad = p.staple_pad = p.llfat_ga_pad = p.mom_ga_pad = INVALID_INT;
  p.gaugeGiB = 0.0;
  p.staggered_phase_type = QUDA_MILC_STAGGERED_PHASE;
  p.return_result_gauge = 1;
  p.return_result_mom = 1;
  return p;
}

QudaInvertParam newQudaInvertParam(void) {
  QudaInvertParam p;
  memset(&p, 0, sizeof(p));
  p.input_location = p.output_location = p.clover_location = QUDA_CPU_FIELD_LOCATION;
  p.dslash_type = QUDA_INVALID_DSLASH;
  p.inv_type = QUDA_INVALID_INVERTER;
  p.mass = p.kappa = p.m5 = p.mu = INVALID_DOUBLE;
  p.epsilon = INVALID_DOUBLE;
  p.Ls = INVALID_INT;
  p.twist_flavor = QUDA_TWIST_INVALID;
  p.tol = INVALID_DOUBLE;
  p.tol_restart = INVALID_DOUBLE;
  p.tol_hq = INVALID_DOUBLE;
  p.residual_type = QUDA_L2_RELATIVE_RESIDUAL;
  p.maxiter = INVALID_INT;
  p.reliable_delta = INVALID_DOUBLE;
  p.use_sloppy_partial_accumulator = 0;
  p.max_res_increase = 1;
  p.max_res_increase_total = 10;
  p.heavy_quark_check = 10;
  p.pipeline = 0;
  p.num_offset = 0;
  p.num_src = 1;
  p.overlap = 0;
  for (int i = 0; i < QUDA_MAX_MULTI_SHIFT; i++) {
    p.offset[i] = p.tol_offset[i] = p.tol_hq_offset[i] = p.true_res_offset[i] = p.iter_res_offset[i] = INVALID_DOUBLE;
  }
  p.solution_type = QUDA_INVALID_SOLUTION;
  p.solve_type = QUDA_INVALID_SOLVE;
  p.matpc_type = QUDA_MATPC_INVALID;
  p.dagger = QUDA_DAG_INVALID;
  p.mass_normalization = QUDA_INVALID_NORMALIZATION;
  p.solver_normalization = QUDA_DEFAULT_NORMALIZATION;
  p.preserve_source = QUDA_PRESERVE_SOURCE_INVALID;
  p.cpu_prec = p.cuda_prec = p.cuda_prec_sloppy = p.cuda_prec_precondition = QUDA_INVALID_PRECISION;
  p.gamma_basis = QUDA_INVALID_GAMMA_BASIS;
  p.dirac_order = QUDA_INVALID_DIRAC_ORDER;
  p.sp_pad = INVALID_INT;
  p.tune = QUDA_TUNE_INVALID;
  p.Nsteps = INVALID_INT;
  p.gcrNkrylov = INVALID_INT;
  p.inv_type_precondition = QUDA_INVALID_INVERTER;
  p.preconditioner = p.preconditionerUP = p.preconditionerDN = 0;
  p.tol_precondition = INVALID_DOUBLE;
  p.maxiter_precondition = INVALID_INT;
  p.verbosity_precondition = QUDA_INVALID_VERBOSITY;
  p.schwarz_type = QUDA_ADDITIVE_SCHWARZ;
  p.precondition_cycle = 1;
  p.use_init_guess = QUDA_USE_INIT_GUESS_NO;
  p.omega = 1.0;
  p.clover_cpu_prec = p.clover_cuda_prec = p.clover_cuda_prec_sloppy = p.clover_cuda_prec_precondition = QUDA_INVALID_PRECISION;
  p.clover_order = QUDA_INVALID_CLOVER_ORDER;
  p.cl_pad = INVALID_INT;
  p.clover_coeff = INVALID_DOUBLE;
  p.verbosity = QUDA_INVALID_VERBOSITY;
  p.cuda_prec_ritz = QUDA_INVALID_PRECISION;
  p.use_reduced_vector_set = 1;
  p.cg_iterref_tol = 5e-2;
  p.eigcg_max_restarts = 2;
  p.max_restart_num = 3;
  p.inc_tol = 1e-2;
  p.eigenval_tol = 1e-1;
  return p;
}

QudaMultigridParam newQudaMultigridParam(void) {
  QudaMultigridParam p;
  memset(&p, 0, sizeof(p));
  p.invert_param = nullptr;
  p.n_level = INVALID_INT;
  for (int i = 0; i < QUDA_MAX_MG_LEVEL; i++) {
    p.smoother[i] = QUDA_INVALID_INVERTER;
    p.smoother_solve_type[i] = QUDA_INVALID_SOLVE;
    for (int j = 0; j < QUDA_MAX_DIM; j++) p.geo_block_size[i][j] = INVALID_INT;
    p.spin_block_size[i] = INVALID_INT;
    p.n_vec[i] = INVALID_INT;
    p.cycle_type[i] = QUDA_MG_CYCLE_INVALID;
    p.nu_pre[i] = p.nu_post[i] = INVALID_INT;
    p.coarse_grid_solution_type[i] = QUDA_INVALID_SOLUTION;
    p.smoother_tol[i] = INVALID_DOUBLE;
    p.global_reduction[i] = QUDA_BOOLEAN_YES;
    p.omega[i] = INVALID_DOUBLE;
    p.location[i] = QUDA_INVALID_FIELD_LOCATION;
  }
  p.setup_maxiter = INVALID_INT;
  p.setup_tol = INVALID_DOUBLE;
  p.compute_null_vector = QUDA_COMPUTE_NULL_VECTOR_INVALID;
  p.generate_all_levels = QUDA_BOOLEAN_INVALID;
  p.run_verify = QUDA_BOOLEAN_INVALID;
  p.delta_muPR = p.delta_kappaPR = p.delta_cswPR = p.delta_muCG = p.delta_kappaCG = p.delta_cswCG = 1.0;
  return p;
}

QudaEigParam newQudaEigParam(void) {
  QudaEigParam p;
  memset(&p, 0, sizeof(p));
  p.RitzMat_lanczos = p.RitzMat_Convcheck = QUDA_INVALID_SOLUTION;
  p.eig_type = QUDA_INVALID_TYPE;
  return p;
}

#define PR_I(s, f) log_msg(0, #f " = %d\n", (int)(s)->f)
#define PR_D(s, f) log_msg(0, #f " = %g\n", (double)(s)->f)
void printQudaGaugeParam(QudaGaugeParam *p) {
  log_msg(0, "QUDA Gauge Parameters:\n");
  PR_I(p, location); for (int i = 0; i < 4; i++) log_msg(0, "X[%d] = %d\n", i, p->X[i]);
  PR_D(p, anisotropy); PR_I(p, type); PR_I(p, gauge_order); PR_I(p, t_boundary); PR_I(p, cpu_prec); PR_I(p, cuda_prec);
  PR_I(p, reconstruct); PR_I(p, cuda_prec_sloppy); PR_I(p, reconstruct_sloppy); PR_I(p, cuda_prec_precondition);
  PR_I(p, reconstruct_precondition); PR_I(p, gauge_fix); PR_I(p, ga_pad); PR_D(p, gaugeGiB);
}
void printQudaInvertParam(QudaInvertParam *p) {
  log_msg(0, "QUDA Inverter Parameters:\n");
  PR_I(p, dslash_type); PR_I(p, inv_type); PR_D(p, kappa); PR_D(p, mu); PR_I(p, twist_flavor); PR_D(p, tol); PR_I(p, maxiter);
  PR_D(p, reliable_delta); PR_I(p, solution_type); PR_I(p, solve_type); PR_I(p, matpc_type); PR_I(p, dagger); PR_I(p, mass_normalization);
  PR_I(p, cpu_prec); PR_I(p, cuda_prec); PR_I(p, cuda_prec_sloppy); PR_I(p, cuda_prec_precondition); PR_I(p, dirac_order);
  PR_I(p, gamma_basis); PR_I(p, gcrNkrylov); PR_I(p, inv_type_precondition); PR_D(p, omega); PR_I(p, verbosity);
  PR_D(p, true_res); PR_I(p, iter); PR_D(p, gflops); PR_D(p, secs);
}
void printQudaMultigridParam(QudaMultigridParam *p) {
  log_msg(0, "QUDA Multigrid Parameters:\n");
  PR_I(p, n_level);
  for (int i = 0; i < p->n_level && i < QUDA_MAX_MG_LEVEL; i++) {
    log_msg(0, "level %d: block %d %d %d %d, spin_block %d, n_vec %d, smoother %d, nu_pre %d, nu_post %d, tol %g, omega %g, cycle %d\n", i,
            p->geo_block_size[i][0], p->geo_block_size[i][1], p->geo_block_size[i][2], p->geo_block_size[i][3], p->spin_block_size[i],
            p->n_vec[i], (int)p->smoother[i], p->nu_pre[i], p->nu_post[i], p->smoother_tol[i], p->omega[i], (int)p->cycle_type[i]);
  }
  PR_I(p, setup_maxiter); PR_D(p, setup_tol); PR_D(p, delta_muPR); PR_D(p, delta_kappaPR);
}

}  // extern "C"

// ---- gauge field --------------------------------------------------------------------------------

namespace qb {
void gather_ghost_links(GaugeField &gf, const Geom &geom, int mu, void *dst, cudaStream_t s);
void free_staging();
}
extern "C" void free_staging_buffers() { qb::free_staging(); }

static void build_gauge_ghost(GaugeField &gf) {
  Runtime &r = rt();
  const Geom &g = G.lat.geom;
  for (int d = 0; d < 4; d++) {
    if (!g.part[d]) continue;
    const size_t bytes = (size_t)2 * gf.recon * gf.store_bytes() * g.faceVh[d];
    if (gf.ghost[d]) QB_CUDA(cudaFree(gf.ghost[d]));
    QB_CUDA(cudaMalloc(&gf.ghost[d], bytes));
    if (r.size == 1 || r.grid[d] == 1) {
      gather_ghost_links(gf, g, d, gf.ghost[d], r.compute);
    } else {
      void *tmp;
      QB_CUDA(cudaMalloc(&tmp, bytes));
      gather_ghost_links(gf, g, d, tmp, r.compute);
      // my last slice goes to the forward neighbour; I receive the backward neighbour's last slice
      comm_sendrecv(tmp, comm_neighbor_rank(d, 1), gf.ghost[d], comm_neighbor_rank(d, 0), bytes, r.compute);
      QB_CUDA(cudaStreamSynchronize(r.compute));
      QB_CUDA(cudaFree(tmp));
    }
  }
  QB_CUDA(cudaStreamSynchronize(r.compute));
}

extern "C" {

// interface_quda.cpp:521-692
void loadGaugeQuda(void *h_gauge, QudaGaugeParam *param) {
  require_init();
  if (!param) QB_ERROR("loadGaugeQuda: null parameter struct");
  for (int i = 0; i < 4; i++)
    if (param->X[i] == INVALID_INT) QB_ERROR("Parameter X[%d] undefined", i);
  if (param->type != QUDA_WILSON_LINKS) QB_ERROR("Only QUDA_WILSON_LINKS gauge fields are supported");
  if (param->anisotropy == INVALID_DOUBLE) QB_ERROR("Parameter anisotropy undefined");
  if (param->gauge_order != QUDA_QDP_GAUGE_ORDER) QB_ERROR("Gauge order %d not supported (QUDA_QDP_GAUGE_ORDER only)", (int)param->gauge_order);
  if (param->t_boundary == QUDA_INVALID_T_BOUNDARY) QB_ERROR("Parameter t_boundary undefined");
  if (param->location != QUDA_CPU_FIELD_LOCATION) QB_ERROR("loadGaugeQuda expects a host gauge field");
  const Prec cpu_prec = to_prec(param->cpu_prec, "cpu_prec");
  const Prec prec = to_prec(param->cuda_prec, "cuda_prec");
  if (param->reconstruct == QUDA_RECONSTRUCT_INVALID) QB_ERROR("Parameter reconstruct undefined");
  // sloppy / precondition default to the precise settings when left unset
  const Prec prec_s = param->cuda_prec_sloppy == QUDA_INVALID_PRECISION ? prec : to_prec(param->cuda_prec_sloppy, "cuda_prec_sloppy");
  const int rec_s = param->reconstruct_sloppy == QUDA_RECONSTRUCT_INVALID ? (int)param->reconstruct : (int)param->reconstruct_sloppy;
  const Prec prec_p = param->cuda_prec_precondition == QUDA_INVALID_PRECISION ? prec_s : to_prec(param->cuda_prec_precondition, "cuda_prec_precondition");
  const int rec_p = param->reconstruct_precondition == QUDA_RECONSTRUCT_INVALID ? rec_s : (int)param->reconstruct_precondition;

  freeGaugeQuda();
  Runtime &r = rt();
  G.lat.init(param->X, (int)param->t_boundary, param->anisotropy);
  G.param = *param;
  const long Vh = G.lat.geom.Vh;
  void *const *hg = (void *const *)h_gauge;

  G.precise.reset(new GaugeField(Vh, prec, (int)param->reconstruct));
  import_gauge(*G.precise, hg, cpu_prec, G.lat.geom, r.compute);
  build_gauge_ghost(*G.precise);
  double gib = (double)G.precise->bytes() / (1 << 30);
  if (prec_s == prec && rec_s == (int)param->reconstruct) G.sloppy = G.precise;
  else {
    G.sloppy.reset(new GaugeField(Vh, prec_s, rec_s));
    import_gauge(*G.sloppy, hg, cpu_prec, G.lat.geom, r.compute);
    build_gauge_ghost(*G.sloppy);
    gib += (double)G.sloppy->bytes() / (1 << 30);
  }
  if (prec_p == prec_s && rec_p == rec_s) G.precondition = G.sloppy;
  else if (prec_p == prec && rec_p == (int)param->reconstruct) G.precondition = G.precise;
  else {
    G.precondition.reset(new GaugeField(Vh, prec_p, rec_p));
    import_gauge(*G.precondition, hg, cpu_prec, G.lat.geom, r.compute);
    build_gauge_ghost(*G.precondition);
    gib += (double)G.precondition->bytes() / (1 << 30);
  }
  param->gaugeGiB = gib;
  G.loaded = true;
}

void freeGaugeQuda(void) {
  if (!G.loaded) return;
  QB_CUDA(cudaDeviceSynchronize());
  pool_clear();
  G.precise.reset(); G.sloppy.reset(); G.precondition.reset();
  G.lat.release();
  G.loaded = false;
}

void saveGaugeQuda(void *h_gauge, QudaGaugeParam *param) {
  require_gauge();
  if (param->gauge_order != QUDA_QDP_GAUGE_ORDER) QB_ERROR("Gauge order %d not supported", (int)param->gauge_order);
  export_gauge((void *const *)h_gauge, *G.precise, to_prec(param->cpu_prec, "cpu_prec"), G.lat.geom, rt().compute);
}

// ---- operator application -----------------------------------------------------------------------
static void load_host_spinor(SpinorField &f, const void *h, const QudaInvertParam *p) {
  if (p->input_location != QUDA_CPU_FIELD_LOCATION) QB_ERROR("input_location must be QUDA_CPU_FIELD_LOCATION (use the *ResidentQudaB200 entry points for device fields)");
  import_spinor(f, h, to_prec(p->cpu_prec, "cpu_prec"), to_basis(p->gamma_basis), to_order(p->dirac_order), rt().compute);
}
static void save_host_spinor(void *h, const SpinorField &f, const QudaInvertParam *p) {
  if (p->output_location != QUDA_CPU_FIELD_LOCATION) QB_ERROR("output_location must be QUDA_CPU_FIELD_LOCATION");
  export_spinor(h, f, to_prec(p->cpu_prec, "cpu_prec"), to_basis(p->gamma_basis), to_order(p->dirac_order), rt().compute);
}

static void dslash_fields(SpinorField &out, SpinorField &in, QudaInvertParam *p, QudaParity parity) {
  if (parity != QUDA_EVEN_PARITY && parity != QUDA_ODD_PARITY) QB_ERROR("invalid parity %d", (int)parity);
  std::unique_ptr<DiracTM> d(make_dirac(p, true, pick_gauge(in.prec)));
  d->Dslash(out, in, (int)parity);
}
static void mat_fields(SpinorField &out, SpinorField &in, QudaInvertParam *p, bool normal) {
  const bool pc = (p->solution_type == QUDA_MATPC_SOLUTION || p->solution_type == QUDA_MATPCDAG_MATPC_SOLUTION);
  if ((pc ? 1 : 2) != in.nparity) QB_ERROR("solution_type %d needs a %s field", (int)p->solution_type, pc ? "single-parity" : "full");
  std::unique_ptr<DiracTM> d(make_dirac(p, pc, pick_gauge(in.prec)));
  if (normal) d->MdagM(out, in);
  else d->M(out, in);
  mass_rescale_out(out, p, pc, normal);
}

// interface_quda.cpp:1496-1569
void dslashQuda(void *h_out, void *h_in, QudaInvertParam *p, QudaParity parity) {
  require_gauge();
  const Prec prec = to_prec(p->cuda_prec, "cuda_prec");
  SpinorField *in = pool_get(1, prec), *out = pool_get(1, prec);
  load_host_spinor(*in, h_in, p);
  dslash_fields(*out, *in, p, parity);
  save_host_spinor(h_out, *out, p);
  pool_put(in); pool_put(out);
}

// interface_quda.cpp:1716-1784
void MatQuda(void *h_out, void *h_in, QudaInvertParam *p) {
  require_gauge();
  const bool pc = (p->solution_type == QUDA_MATPC_SOLUTION || p->solution_type == QUDA_MATPCDAG_MATPC_SOLUTION);
  const Prec prec = to_prec(p->cuda_prec, "cuda_prec");
  SpinorField *in = pool_get(pc ? 1 : 2, prec), *out = pool_get(pc ? 1 : 2, prec);
  load_host_spinor(*in, h_in, p);
  mat_fields(*out, *in, p, false);
  save_host_spinor(h_out, *out, p);
  pool_put(in); pool_put(out);
}

// interface_quda.cpp:1786-1863
void MatDagMatQuda(void *h_out, void *h_in, QudaInvertParam *p) {
  require_gauge();
  const bool pc = (p->solution_type == QUDA_MATPC_SOLUTION || p->solution_type == QUDA_MATPCDAG_MATPC_SOLUTION);
  const Prec prec = to_prec(p->cuda_prec, "cuda_prec");
  SpinorField *in = pool_get(pc ? 1 : 2, prec), *out = pool_get(pc ? 1 : 2, prec);
  load_host_spinor(*in, h_in, p);
  mat_fields(*out, *in, p, true);
  save_host_spinor(h_out, *out, p);
  pool_put(in); pool_put(out);
}

// ---- resident-field extensions (include/quda_b200_ext.h) ---------------------------------------
void *newSpinorQudaB200(QudaSiteSubset site_subset, QudaPrecision precision) {
  require_gauge();
  if (site_subset != QUDA_PARITY_SITE_SUBSET && site_subset != QUDA_FULL_SITE_SUBSET) QB_ERROR("invalid site subset %d", (int)site_subset);
  return new SpinorField(G.lat.geom.Vh, (int)site_subset, to_prec(precision, "precision"));
}
void freeSpinorQudaB200(void *f) { delete (SpinorField *)f; }
void loadSpinorQudaB200(void *f, const void *h_in, QudaInvertParam *p) {
  require_gauge();
  load_host_spinor(*(SpinorField *)f, h_in, p);
  QB_CUDA(cudaStreamSynchronize(rt().compute));
}
void saveSpinorQudaB200(void *h_out, const void *f, QudaInvertParam *p) {
  require_gauge();
  save_host_spinor(h_out, *(const SpinorField *)f, p);
}
void dslashResidentQudaB200(void *out, void *in, QudaInvertParam *p, QudaParity parity) {
  require_gauge();
  dslash_fields(*(SpinorField *)out, *(SpinorField *)in, p, parity);
}
void matResidentQudaB200(void *out, void *in, QudaInvertParam *p) {
  require_gauge();
  mat_fields(*(SpinorField *)out, *(SpinorField *)in, p, false);
}
void matDagMatResidentQudaB200(void *out, void *in, QudaInvertParam *p) {
  require_gauge();
  mat_fields(*(SpinorField *)out, *(SpinorField *)in, p, true);
}

double timeDslashQudaB200(void *out, void *in, QudaInvertParam *p, QudaParity parity, int niter, float *per_iter_ms) {
  require_gauge();
  if (niter < 1) QB_ERROR("niter must be >= 1");
  Runtime &r = rt();
  SpinorField &o = *(SpinorField *)out, &i = *(SpinorField *)in;
  std::unique_ptr<DiracTM> d(make_dirac(p, true, pick_gauge(i.prec)));
  std::vector<cudaEvent_t> ev(niter + 1);
  for (auto &e : ev) QB_CUDA(cudaEventCreate(&e));
  QB_CUDA(cudaStreamSynchronize(r.compute));
  QB_CUDA(cudaEventRecord(ev[0], r.compute));
  for (int k = 0; k < niter; k++) {
    d->Dslash(o, i, (int)parity);
    QB_CUDA(cudaEventRecord(ev[k + 1], r.compute));
  }
  QB_CUDA(cudaStreamSynchronize(r.compute));
  float total = 0;
  QB_CUDA(cudaEventElapsedTime(&total, ev[0], ev[niter]));
  if (per_iter_ms)
    for (int k = 0; k < niter; k++) QB_CUDA(cudaEventElapsedTime(&per_iter_ms[k], ev[k], ev[k + 1]));
  for (auto &e : ev) cudaEventDestroy(e);
  return (double)total / niter;
}

long long kernelLaunchCountQudaB200(void) { return rt().launches; }
void *computeStreamQudaB200(void) { return (void *)rt().compute; }
void syncQudaB200(void) { QB_CUDA(cudaDeviceSynchronize()); }

void ncclUniqueIdQudaB200(void *out) { comm_unique_id(out); }
void commsBootstrapQudaB200(int rank, int size, const void *id) { comm_bootstrap(rank, size, id); }
void commDimPartitionedSetQudaB200(int mask) {
  if (G.loaded) QB_ERROR("commDimPartitionedSetQudaB200 must be called before loadGaugeQuda");
  rt().part_mask |= (mask & 15);
}

// ---- outside the hot path: exported for link compatibility ------------------------------------
void loadCloverQuda(void *, void *, QudaInvertParam *) { QB_ERROR("loadCloverQuda: clover operators are outside this build's scope (SURVEY.md section 8f.1)"); }
void freeCloverQuda(void) {}
void invertMultiSrcQuda(void **, void **, QudaInvertParam *) { QB_ERROR("invertMultiSrcQuda is not implemented (SURVEY.md section 8f.4)"); }
void invertMultiShiftQuda(void **, void *, QudaInvertParam *) { QB_ERROR("invertMultiShiftQuda is outside this build's scope"); }
void cloverQuda(void *, void *, QudaInvertParam *, QudaParity *, int) { QB_ERROR("cloverQuda is outside this build's scope"); }

}  // extern "C"
