// C-ABI entry points: the quda.h subset of SURVEY.md section 8(b) plus the resident-field extensions.
// Behavioural model: /root/reference/lib/interface_quda.cpp (line refs at each function).
#include <algorithm>
#include <cctype>
#include <sched.h>
#include <cfloat>
#include <climits>
#include <cstring>
#include <chrono>
#include <limits>
#include <map>
#include <string>
#include <vector>
#include "../../include/quda.h"
#include "blas.h"
#include "comm.h"
#include "dirac.h"
#include "multigrid.h"
#include "solver.h"

using namespace qb;

#define INVALID_INT QUDA_INVALID_ENUM
#define INVALID_DOUBLE DBL_MIN

namespace {

struct GaugeSet {
  Lattice lat;
  QudaGaugeParam param;
  std::shared_ptr<GaugeField> precise, sloppy, precondition;
  bool loaded = false;
  CloverSet clover;  // loadCloverQuda (process-global like the gauge field, interface_quda.cpp:723-900)
  std::unique_ptr<SpinorField> solution_resident;  // make_resident_solution (solutionResident, interface_quda.cpp:2500-2508)
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
HostGaugeOrder host_gauge_order(QudaGaugeFieldOrder o) {
  if (o == QUDA_QDP_GAUGE_ORDER) return GAUGE_ORDER_QDP;
  if (o == QUDA_MILC_GAUGE_ORDER) return GAUGE_ORDER_MILC;
  if (o == QUDA_CPS_WILSON_GAUGE_ORDER) return GAUGE_ORDER_CPS;
  QB_ERROR("Gauge order %d not supported (QUDA_QDP_GAUGE_ORDER, QUDA_MILC_GAUGE_ORDER, QUDA_CPS_WILSON_GAUGE_ORDER)", (int)o);
}
HostSpinorOrder to_order(QudaDiracFieldOrder o) {
  if (o == QUDA_DIRAC_ORDER) return ORDER_SPIN_COLOR;
  if (o == QUDA_QDP_DIRAC_ORDER) return ORDER_COLOR_SPIN;
  QB_ERROR("Dirac order %d not supported", (int)o);
}

void check_invert_param_operator(const QudaInvertParam *p) {
  if (p->dslash_type != QUDA_TWISTED_MASS_DSLASH && p->dslash_type != QUDA_WILSON_DSLASH && p->dslash_type != QUDA_TWISTED_CLOVER_DSLASH &&
      p->dslash_type != QUDA_CLOVER_WILSON_DSLASH)
    QB_ERROR("Unsupported dslash_type %d (this build covers Wilson, Wilson-clover, degenerate twisted mass and twisted clover)", (int)p->dslash_type);
  if (p->kappa == INVALID_DOUBLE) QB_ERROR("Parameter kappa undefined");
  if ((p->dslash_type == QUDA_TWISTED_CLOVER_DSLASH || p->dslash_type == QUDA_CLOVER_WILSON_DSLASH) && !G.clover.loaded)
    QB_ERROR("Clover field not allocated (call loadCloverQuda)");
  if (p->dslash_type == QUDA_TWISTED_MASS_DSLASH || p->dslash_type == QUDA_TWISTED_CLOVER_DSLASH) {
    if (p->mu == INVALID_DOUBLE) QB_ERROR("Parameter mu undefined");
    if (p->twist_flavor != QUDA_TWIST_PLUS && p->twist_flavor != QUDA_TWIST_MINUS && p->twist_flavor != QUDA_TWIST_NONDEG_DOUBLET)
      QB_ERROR("Twist flavor not set %d (QUDA_TWIST_PLUS / MINUS or QUDA_TWIST_NONDEG_DOUBLET)", (int)p->twist_flavor);
    if (p->twist_flavor == QUDA_TWIST_NONDEG_DOUBLET) {
      if (p->dslash_type != QUDA_TWISTED_MASS_DSLASH) QB_ERROR("The non-degenerate doublet is implemented for QUDA_TWISTED_MASS_DSLASH only");
      if (p->epsilon == INVALID_DOUBLE) QB_ERROR("Parameter epsilon undefined");
    }
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
  const bool twisted = p->dslash_type == QUDA_TWISTED_MASS_DSLASH || p->dslash_type == QUDA_TWISTED_CLOVER_DSLASH;
  const int flavor = twisted ? (int)p->twist_flavor : 0;
  const double mu = twisted ? p->mu * mu_scale : 0.0;
  DiracTM *d = new DiracTM(&G.lat, gauge, p->kappa * kappa_scale, mu, flavor, pc, (int)p->matpc_type, p->dagger == QUDA_DAG_YES);
  if (flavor == 2) d->epsilon = p->epsilon;
  if (p->dslash_type == QUDA_TWISTED_CLOVER_DSLASH || p->dslash_type == QUDA_CLOVER_WILSON_DSLASH) d->clover = &G.clover;
  return d;
}

// small pool of resident work fields so that dslashQuda / MatQuda do not cudaMalloc per call
std::map<std::pair<int, int>, std::vector<SpinorField *>> pool;
SpinorField *pool_get(int nparity, Prec prec, int nflavor = 1) {
  auto &v = pool[{nparity + 16 * nflavor, (int)prec}];
  if (!v.empty() && v.back()->Vh == G.lat.geom.Vh) {
    SpinorField *f = v.back();
    v.pop_back();
    return f;
  }
  return new SpinorField(G.lat.geom.Vh, nparity, prec, 4, 3, 1, nflavor);
}
void pool_put(SpinorField *f) { pool[{f->nparity + 16 * f->nflavor, (int)f->prec}].push_back(f); }
// flavours per field of the operator selected by the parameters (2: non-degenerate twisted-mass doublet, host fields [parity][flavour][site])
int nflavor_of(const QudaInvertParam *p) {
  return (p->dslash_type == QUDA_TWISTED_MASS_DSLASH && p->twist_flavor == QUDA_TWIST_NONDEG_DOUBLET) ? 2 : 1;
}
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
void set_halo_only(bool on);  // dslash_host.cu
void set_hop_phase(int mask); // dslash_host.cu: 1 pack + exchange, 2 boundary sites, 4 interior (7 = a whole hop)
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

// CPU / memory affinity of the process to the NUMA node its GPU hangs off (the reference: setNumaAffinityNVML from initQudaDevice,
// lib/interface_quda.cpp:429-434, lib/numa_affinity.cpp, switched off by QUDA_ENABLE_NUMA=0).  Here without NVML: PCI bus id -> sysfs numa_node
// -> cpulist -> sched_setaffinity; pinned host buffers allocated afterwards land on that node (first touch), which is what keeps the
// H2D / D2H copies of 8 ranks on one host from all crossing the socket interconnect.  Only when several ranks share the host.
static void set_numa_affinity(int dev) {
  const char *env = getenv("QUDA_ENABLE_NUMA");
  if (env && strcmp(env, "0") == 0) return;
  if (rt().size <= 1 && !(env && strcmp(env, "1") == 0)) return;
  char bus[32] = "";
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), dev) != cudaSuccess) { cudaGetLastError(); return; }
  for (char *c = bus; *c; c++) *c = (char)tolower(*c);
  char path[128];
  snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
  FILE *f = fopen(path, "r");
  if (!f) return;
  int node = -1;
  if (fscanf(f, "%d", &node) != 1) node = -1;
  fclose(f);
  if (node < 0) return;
  snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
  f = fopen(path, "r");
  if (!f) return;
  char list[4096] = "";
  if (!fgets(list, sizeof(list), f)) list[0] = 0;
  fclose(f);
  cpu_set_t set;
  CPU_ZERO(&set);
  int n = 0;
  for (char *tok = strtok(list, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
    int a = 0, b = 0;
    if (sscanf(tok, "%d-%d", &a, &b) == 2) { for (int c = a; c <= b; c++) { CPU_SET(c, &set); n++; } }
    else if (sscanf(tok, "%d", &a) == 1) { CPU_SET(a, &set); n++; }
  }
  if (n == 0) return;
  if (sched_setaffinity(0, sizeof(set), &set) == 0) log_msg(2, "Set NUMA affinity for device %d: node %d, %d cpus\n", dev, node, n);
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
  set_numa_affinity(dev);
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
void pipe_cleanup_c();

void endQuda(void) {
  Runtime &r = rt();
  if (!r.device_ready) return;
  if (r.memory_ready) {
    QB_CUDA(cudaDeviceSynchronize());
    freeGaugeQuda();
    blas::end();
    pipe_cleanup_c();
    free_staging_buffers();
    pool_release_all();
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
  if (param->gauge_order != QUDA_QDP_GAUGE_ORDER && param->gauge_order != QUDA_MILC_GAUGE_ORDER && param->gauge_order != QUDA_CPS_WILSON_GAUGE_ORDER)
    QB_ERROR("Gauge order %d not supported (QUDA_QDP_GAUGE_ORDER, QUDA_MILC_GAUGE_ORDER, QUDA_CPS_WILSON_GAUGE_ORDER)", (int)param->gauge_order);
  if (param->t_boundary == QUDA_INVALID_T_BOUNDARY) QB_ERROR("Parameter t_boundary undefined");
  if (param->location != QUDA_CPU_FIELD_LOCATION) QB_ERROR("loadGaugeQuda expects a host gauge field");
  // interface_quda.cpp:582-589: use_resident_gauge takes the device field a gauge-update routine left behind (make_resident_gauge of
  // the force / update entry points, outside this build's scope), so there is never one to use
  if (param->use_resident_gauge != 0 && param->use_resident_gauge != INVALID_INT)
    QB_ERROR("No resident gauge field (use_resident_gauge = %d: this build has no gauge-update routine that could have left one; pass the host field)", param->use_resident_gauge);
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
  const HostGaugeOrder gorder = host_gauge_order(param->gauge_order);

  G.precise.reset(new GaugeField(Vh, prec, (int)param->reconstruct));
  import_gauge(*G.precise, hg, cpu_prec, G.lat.geom, r.compute, gorder);
  build_gauge_ghost(*G.precise);
  double gib = (double)G.precise->bytes() / (1 << 30);
  if (prec_s == prec && rec_s == (int)param->reconstruct) G.sloppy = G.precise;
  else {
    G.sloppy.reset(new GaugeField(Vh, prec_s, rec_s));
    import_gauge(*G.sloppy, hg, cpu_prec, G.lat.geom, r.compute, gorder);
    build_gauge_ghost(*G.sloppy);
    gib += (double)G.sloppy->bytes() / (1 << 30);
  }
  if (prec_p == prec_s && rec_p == rec_s) G.precondition = G.sloppy;
  else if (prec_p == prec && rec_p == (int)param->reconstruct) G.precondition = G.precise;
  else {
    G.precondition.reset(new GaugeField(Vh, prec_p, rec_p));
    import_gauge(*G.precondition, hg, cpu_prec, G.lat.geom, r.compute, gorder);
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
  G.clover.release();
  G.solution_resident.reset();
  G.precise.reset(); G.sloppy.reset(); G.precondition.reset();
  G.lat.release();
  pool_release_all();
  G.loaded = false;
}

void saveGaugeQuda(void *h_gauge, QudaGaugeParam *param) {
  require_gauge();
  export_gauge((void *const *)h_gauge, *G.precise, to_prec(param->cpu_prec, "cpu_prec"), G.lat.geom, rt().compute, host_gauge_order(param->gauge_order));
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

// Pipelined host path of dslashQuda: the parity field is cut into T-slabs; slab c is converted and multiplied as soon
// as slabs c-1, c, c+1 have landed, and its result starts its way back while later slabs are still arriving, so that
// H2D, the hop and D2H overlap on the two copy engines (the reference moves the whole field in, computes, moves it out:
// interface_quda.cpp:1509-1557).  Unpartitioned lattices only.
struct PipeState {
  cudaStream_t h2d = nullptr, d2h = nullptr;
  std::vector<cudaEvent_t> ev_in, ev_out;
  cudaEvent_t ev_done = nullptr;
};
static PipeState pipe_state;

void pipe_cleanup_c() {
  if (pipe_state.h2d) { cudaStreamDestroy(pipe_state.h2d); cudaStreamDestroy(pipe_state.d2h); cudaEventDestroy(pipe_state.ev_done); }
  for (auto e : pipe_state.ev_in) cudaEventDestroy(e);
  for (auto e : pipe_state.ev_out) cudaEventDestroy(e);
  pipe_state.h2d = pipe_state.d2h = nullptr; pipe_state.ev_done = nullptr; pipe_state.ev_in.clear(); pipe_state.ev_out.clear();
}

static bool dslash_pipelined(void *h_out, void *h_in, QudaInvertParam *p, QudaParity parity, SpinorField &in, SpinorField &out) {
  const Geom &g = G.lat.geom;
  // partitioned lattices: T-only splits (the default decomposition up to 4 ranks, and the weak-scaling layout of bench.py).  The faces are
  // the time slices 0 and T-1: they travel first, the halo exchange starts as soon as they have landed and overlaps the remaining copies;
  // the interior is multiplied slab by slab as on one GPU, the two boundary slices at the end.
  if (g.part[0] || g.part[1] || g.part[2]) return false;
  const bool part_t = g.part[3] != 0;
  if (p->dslash_type == QUDA_TWISTED_CLOVER_DSLASH || p->dslash_type == QUDA_CLOVER_WILSON_DSLASH || in.nflavor != 1) return false;  // two kernels per hop: plain path
  if (p->input_location != QUDA_CPU_FIELD_LOCATION || p->output_location != QUDA_CPU_FIELD_LOCATION) return false;
  const int T = g.X[3];
  // Slabs of whole time slices.  A slab can be multiplied once it and its two t-neighbours have landed, so whatever depends on the LAST
  // slab to arrive (that slab and its two neighbours) leaves only after the H2D stream has finished: the exposed tail.  The very last slab
  // is sent FIRST (slab 0 needs it across the periodic boundary).  Uniform slabs on B200 (32^3x64, fp32): 16 slabs 2.88 ms, 32 slabs 3.28 ms.
  // Measured with CUDA events on B200 (QB_PIPE_TRACE=1, 32^3x64 fp32, link floor 2.04 ms): every extra copy costs the H2D stream ~30 us,
  // the first D2H cannot start before the first three slabs are in, and the D2H stream runs one slab behind the H2D stream.  So: few
  // slabs, thin ones at the head (the first results leave early) and tapering ones at the end (little is left when the last input lands).
  const int want = getenv("QB_PIPE_CHUNKS") ? atoi(getenv("QB_PIPE_CHUNKS")) : 13;   // 9 / 11 / 13 slabs: 3.06 / 2.82 / 2.73 ms (uniform 16: 2.88 ms)
  const bool uniform = getenv("QB_PIPE_UNIFORM") && atoi(getenv("QB_PIPE_UNIFORM"));   // two getenv per call (microseconds against a millisecond call): the tests switch them per case
  const bool tapered = !uniform && T >= 32 && want >= 8;
  std::vector<int> tslices;   // time slices per slab
  if (tapered) {
    const int head = std::max(1, T / 32);
    // T partition: the last slab is exactly the face slice T-1 plus its neighbour (the boundary kernel reads T-2 as well)
    const int tail[4] = {std::max(1, T / 16), std::max(1, T / 32), 1, part_t ? 2 : 1};
    const int rest = T - 2 * head - (tail[0] + tail[1] + tail[2] + tail[3]);
    const int nbig = std::max(1, std::min(want - 6, rest));
    tslices.push_back(head); tslices.push_back(head);
    for (int c = 0; c < nbig; c++) tslices.push_back(rest / nbig + (c < rest % nbig ? 1 : 0));
    for (int c = 0; c < 4; c++) tslices.push_back(tail[c]);
  } else {
    int n = 0;
    for (int c : {32, 16, 8, 4}) if (c <= want && T % c == 0 && T / c >= 2) { n = c; break; }
    for (int c = 0; c < n; c++) tslices.push_back(T / n);
  }
  const int nchunk = (int)tslices.size();
  if (nchunk < 4 || (long)g.Vh * 24 * 4 < (4l << 20)) return false;  // small fields: latency dominates, keep it simple
  Runtime &r = rt();
  if (!pipe_state.h2d) {
    QB_CUDA(cudaStreamCreateWithFlags(&pipe_state.h2d, cudaStreamNonBlocking));
    QB_CUDA(cudaStreamCreateWithFlags(&pipe_state.d2h, cudaStreamNonBlocking));
    QB_CUDA(cudaEventCreateWithFlags(&pipe_state.ev_done, cudaEventDisableTiming));
  }
  while ((int)pipe_state.ev_in.size() < nchunk) {
    cudaEvent_t a, b;
    QB_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    QB_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    pipe_state.ev_in.push_back(a); pipe_state.ev_out.push_back(b);
  }
  const Prec hp = to_prec(p->cpu_prec, "cpu_prec");
  const HostBasis basis = to_basis(p->gamma_basis);
  const HostSpinorOrder order = to_order(p->dirac_order);
  const size_t site_bytes = 24 * (size_t)hp;
  const size_t hb = site_bytes * g.Vh;
  char *stage_in = (char *)staging(2 * hb), *stage_out = stage_in + hb;
  const long slice_sites = g.Vh / T;   // checkerboard sites per time slice (cb index is t-slowest)
  std::vector<long> begin(nchunk), count(nchunk);
  for (int c = 0, t0 = 0; c < nchunk; t0 += tslices[c], c++) { begin[c] = t0 * slice_sites; count[c] = tslices[c] * slice_sites; }
  std::unique_ptr<DiracTM> d(make_dirac(p, true, pick_gauge(in.prec)));
  // everything previously queued on the compute stream must be done before the staging buffers are reused
  QB_CUDA(cudaEventRecord(pipe_state.ev_done, r.compute));
  QB_CUDA(cudaStreamWaitEvent(pipe_state.h2d, pipe_state.ev_done, 0));
  // H2D stream: nothing but back-to-back copies (the reorder kernels run on the compute stream), last slab first
  static const bool trace = getenv("QB_PIPE_TRACE") && atoi(getenv("QB_PIPE_TRACE"));
  cudaEvent_t tr[4] = {nullptr, nullptr, nullptr, nullptr};   // h2d begin / end, d2h first copy begin / last copy end
  if (trace) for (auto &e : tr) QB_CUDA(cudaEventCreate(&e));
  if (trace) QB_CUDA(cudaEventRecord(tr[0], pipe_state.h2d));
  std::vector<int> arrival;
  arrival.push_back(nchunk - 1);
  for (int c = 0; c < nchunk - 1; c++) arrival.push_back(c);
  for (int c : arrival) {
    QB_CUDA(cudaMemcpyAsync(stage_in + begin[c] * site_bytes, (const char *)h_in + begin[c] * site_bytes, count[c] * site_bytes, cudaMemcpyHostToDevice, pipe_state.h2d));
    QB_CUDA(cudaEventRecord(pipe_state.ev_in[c], pipe_state.h2d));
  }
  if (trace) QB_CUDA(cudaEventRecord(tr[1], pipe_state.h2d));
  bool first_out = true;
  auto ship = [&](long b0, long n, cudaEvent_t ev) {   // reorder [b0, b0 + n) of `out` for the host and start its way back
    if (n <= 0) return;
    export_spinor_range(stage_out, out, hp, basis, order, b0, n, r.compute);
    QB_CUDA(cudaEventRecord(ev, r.compute));
    QB_CUDA(cudaStreamWaitEvent(pipe_state.d2h, ev, 0));
    if (trace && first_out) { QB_CUDA(cudaEventRecord(tr[2], pipe_state.d2h)); first_out = false; }
    QB_CUDA(cudaMemcpyAsync((char *)h_out + b0 * site_bytes, stage_out + b0 * site_bytes, n * site_bytes, cudaMemcpyDeviceToHost, pipe_state.d2h));
  };
  auto process = [&](int k) {
    d->DslashRange(out, in, (int)parity, (int)begin[k], (int)count[k], r.compute);   // partitioned: clipped to the interior slices
    long b0 = begin[k], n = count[k];
    if (part_t && k == 0) { b0 += slice_sites; n -= slice_sites; }   // the face slices leave after the boundary launch
    if (part_t && k == nchunk - 1) n -= slice_sites;
    ship(b0, n, pipe_state.ev_out[k]);
  };
  // convert every slab as it lands; slab k is multiplied once k-1, k, k+1 (periodic) are there: arrival order n-1, 0, 1, ... => after
  // slab c >= 1 has landed, slab c-1 is complete; the tail is n-2 and n-1
  for (size_t a = 0; a < arrival.size(); a++) {
    const int c = arrival[a];
    QB_CUDA(cudaStreamWaitEvent(r.compute, pipe_state.ev_in[c], 0));
    import_spinor_range(in, stage_in, hp, basis, order, begin[c], count[c], r.compute);
    if (part_t && a == 1) {
      // both face slices are on the device: pack and exchange them now (halo stream), under the copies still to come
      set_hop_phase(1);
      d->Dslash(out, in, (int)parity);
      set_hop_phase(7);
    }
    if (a >= 2) process(c - 1);
  }
  process(nchunk - 2);
  process(nchunk - 1);
  if (part_t) {
    // boundary sites (time slices 0 and T-1) from the ghost zone, then their way back
    set_hop_phase(2);
    d->Dslash(out, in, (int)parity);
    set_hop_phase(7);
    while ((int)pipe_state.ev_out.size() < nchunk + 2) {
      cudaEvent_t e;
      QB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      pipe_state.ev_out.push_back(e);
    }
    ship(0, slice_sites, pipe_state.ev_out[nchunk]);
    ship((long)(T - 1) * slice_sites, slice_sites, pipe_state.ev_out[nchunk + 1]);
  }
  if (trace) {
    QB_CUDA(cudaEventRecord(tr[3], pipe_state.d2h));
    QB_CUDA(cudaEventSynchronize(tr[3]));
    float h2d = 0, first = 0, total = 0, tail = 0;
    cudaEventElapsedTime(&h2d, tr[0], tr[1]); cudaEventElapsedTime(&first, tr[0], tr[2]); cudaEventElapsedTime(&total, tr[0], tr[3]); cudaEventElapsedTime(&tail, tr[1], tr[3]);
    fprintf(stderr, "dslashQuda pipeline: %d slabs, H2D %.3f ms, first D2H starts at %.3f ms, last D2H ends at %.3f ms (%.3f ms after the last H2D)\n", nchunk, h2d, first, total, tail);
    for (auto &e : tr) cudaEventDestroy(e);
  }
  QB_CUDA(cudaStreamSynchronize(pipe_state.d2h));
  QB_CUDA(cudaStreamSynchronize(r.compute));
  return true;
}

// interface_quda.cpp:1496-1569
void dslashQuda(void *h_out, void *h_in, QudaInvertParam *p, QudaParity parity) {
  require_gauge();
  const Prec prec = to_prec(p->cuda_prec, "cuda_prec");
  SpinorField *in = pool_get(1, prec, nflavor_of(p)), *out = pool_get(1, prec, nflavor_of(p));
  if (parity != QUDA_EVEN_PARITY && parity != QUDA_ODD_PARITY) QB_ERROR("invalid parity %d", (int)parity);
  if (dslash_pipelined(h_out, h_in, p, parity, *in, *out)) {
    pool_put(in); pool_put(out);
    return;
  }
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
  SpinorField *in = pool_get(pc ? 1 : 2, prec, nflavor_of(p)), *out = pool_get(pc ? 1 : 2, prec, nflavor_of(p));
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
  SpinorField *in = pool_get(pc ? 1 : 2, prec, nflavor_of(p)), *out = pool_get(pc ? 1 : 2, prec, nflavor_of(p));
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
// solution left on the device by the last invertQuda with make_resident_solution = 1 (owned by the library until the next such solve,
// freeGaugeQuda or endQuda); NULL if there is none.  Usable with saveSpinorQudaB200 / the *Resident* operators.
void *residentSolutionQudaB200(void) { return G.solution_resident.get(); }
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

// mean device time (ms) of the halo part of one hop alone (face pack kernel + NCCL exchange of every partitioned face, no interior /
// boundary kernel): what the NVLink path delivers when nothing hides it.  bytes_out (may be NULL) = bytes this rank sends per hop.
double timeHaloQudaB200(void *out, void *in, QudaInvertParam *p, QudaParity parity, int niter, double *bytes_out) {
  require_gauge();
  Runtime &r = rt();
  SpinorField &o = *(SpinorField *)out, &i = *(SpinorField *)in;
  const Geom &g = G.lat.geom;
  double bytes = 0;
  for (int d = 0; d < 4; d++)
    if (g.part[d]) bytes += 2.0 * g.faceVh[d] * (12.0 * (i.prec == PREC_HALF ? 2 : (int)i.prec) + (i.prec == PREC_HALF ? 4 : 0));
  if (bytes_out) *bytes_out = bytes;
  if (bytes == 0) return 0.0;
  std::unique_ptr<DiracTM> d(make_dirac(p, true, pick_gauge(i.prec)));
  set_halo_only(true);
  d->Dslash(o, i, (int)parity);
  cudaEvent_t e0, e1;
  QB_CUDA(cudaEventCreate(&e0)); QB_CUDA(cudaEventCreate(&e1));
  QB_CUDA(cudaStreamSynchronize(r.compute));
  QB_CUDA(cudaEventRecord(e0, r.compute));
  for (int k = 0; k < niter; k++) d->Dslash(o, i, (int)parity);
  QB_CUDA(cudaEventRecord(e1, r.compute));
  QB_CUDA(cudaStreamSynchronize(r.compute));
  set_halo_only(false);
  float total = 0;
  QB_CUDA(cudaEventElapsedTime(&total, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return (double)total / niter;
}

// mean device time (ms) of one batched hop on `nbatch` resident random fp32 parity fields (one launch: links read once for all members)
double timeDslashBatchQudaB200(QudaInvertParam *p, QudaParity parity, int nbatch, int niter, double *max_dev) {
  require_gauge();
  Runtime &r = rt();
  SpinorField in(G.lat.geom.Vh, 1, PREC_SINGLE, 4, 3, nbatch), out(G.lat.geom.Vh, 1, PREC_SINGLE, 4, 3, nbatch);
  for (int c = 0; c < nbatch; c++) { SpinorField m; in.member(m, c); random_fill(m, 1234 + c); }
  std::unique_ptr<DiracTM> d(make_dirac(p, true, pick_gauge(PREC_SINGLE)));
  for (int k = 0; k < 3; k++) d->Dslash(out, in, (int)parity);
  cudaEvent_t e0, e1;
  QB_CUDA(cudaEventCreate(&e0)); QB_CUDA(cudaEventCreate(&e1));
  QB_CUDA(cudaEventRecord(e0, r.compute));
  for (int k = 0; k < niter; k++) d->Dslash(out, in, (int)parity);
  QB_CUDA(cudaEventRecord(e1, r.compute));
  QB_CUDA(cudaStreamSynchronize(r.compute));
  float ms = 0;
  QB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (max_dev) {  // every member against the single-field kernel on the same input: same arithmetic, so the deviation must be exactly zero
    *max_dev = 0.0;
    SpinorField ref(G.lat.geom.Vh, 1, PREC_SINGLE);
    for (int c = 0; c < nbatch; c++) {
      SpinorField mi, mo;
      in.member(mi, c); out.member(mo, c);
      d->Dslash(ref, mi, (int)parity);
      const double n2 = blas::norm2(ref);
      const double dev = sqrt(blas::xmyNorm(mo, ref) / n2);
      if (!(dev <= *max_dev)) *max_dev = dev;
    }
  }
  return (double)ms / niter;
}

void setDslashBlockSizeQudaB200(int block) {
  if (block != 0 && (block < 32 || block > 128 || (block & 31))) QB_ERROR("dslash block size %d must be 0 (default) or a multiple of 32 in [32, 128]", block);
  G.lat.block_size = block;
}
long long kernelLaunchCountQudaB200(void) { return rt().launches; }
void *computeStreamQudaB200(void) { return (void *)rt().compute; }
void syncQudaB200(void) { QB_CUDA(cudaDeviceSynchronize()); }

void faceIndexMapQudaB200(int dim, int face_num, int parity, int *h_cb_out) {
  require_gauge();
  if (dim < 0 || dim > 3 || (face_num != 0 && face_num != 1) || (parity != 0 && parity != 1)) QB_ERROR("faceIndexMapQudaB200: bad arguments");
  face_index_map(G.lat, dim, face_num, parity, h_cb_out);
}
void commRankInfoQudaB200(int *info10) {
  Runtime &r = rt();
  info10[0] = r.rank; info10[1] = r.size;
  for (int d = 0; d < 4; d++) { info10[2 + d] = r.coord[d]; info10[6 + d] = r.grid[d]; }
}
// 1 when global reductions run their all-reduce inside the reduction kernel over the NVLink peer mailboxes (comm.h), 0 when they use
// ncclAllReduce (single rank: 0)
int commPeerReduceActiveQudaB200(void) { return comm_peer_reduce_ready() ? 1 : 0; }
// mean host-visible latency in microseconds of one global norm2 of an fp32 vector of n_reals reals (kernel + all-reduce over the ranks +
// the stream synchronisation after which the host holds the sum), over niter calls.  use_peer = 0 forces the ncclAllReduce path.
double timeReduceQudaB200(long n_reals, int niter, int use_peer) {
  require_init();
  if (n_reals < 24) n_reals = 24;
  SpinorField f(n_reals / 24, 1, PREC_SINGLE);
  random_fill(f, 12345ull + rt().rank);
  const bool had = comm_peer_reduce_ready();
  if (!use_peer) comm_peer_reduce_enable(false);
  double sink = 0;
  for (int i = 0; i < 5; i++) sink += blas::norm2(f);
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < niter; i++) sink += blas::norm2(f);
  const double us = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() * 1e6 / niter;
  comm_peer_reduce_enable(had);
  if (sink == -1.0) log_msg(0, "");
  return us;
}
void ncclUniqueIdQudaB200(void *out) { comm_unique_id(out); }
void commsBootstrapQudaB200(int rank, int size, const void *id) {
  comm_bootstrap(rank, size, id);
  if (size > 1 && rt().device_ready) set_numa_affinity(rt().device);   // the device was chosen before the ranks were known
}
void commDimPartitionedSetQudaB200(int mask) {
  if (G.loaded) QB_ERROR("commDimPartitionedSetQudaB200 must be called before loadGaugeQuda");
  rt().part_mask |= (mask & 15);
}

// ---- outside the hot path: exported for link compatibility ------------------------------------
// interface_quda.cpp:723-900.  The clover term itself must come from the host (QUDA_PACKED_CLOVER_ORDER); its inverses are
// always computed on the device.  For a twisted-clover operator return_clover_inverse hands back (C^2 + (2 kappa mu)^2)^-1,
// the field the reference's host verification code expects (tests/dslash_test.cpp:339-349, lib/clover_invert.cu:56-90).
void loadCloverQuda(void *h_clover, void *h_clovinv, QudaInvertParam *p) {
  require_gauge();
  if (!h_clover || p->compute_clover) QB_ERROR("loadCloverQuda: computing the clover term from the gauge field is not implemented, pass h_clover");
  if (p->clover_order != QUDA_PACKED_CLOVER_ORDER) QB_ERROR("loadCloverQuda: only QUDA_PACKED_CLOVER_ORDER is supported (got %d)", (int)p->clover_order);
  const Prec hp = to_prec(p->clover_cpu_prec, "clover_cpu_prec");
  G.clover.load(h_clover, hp, G.lat.geom.Vh);
  if (h_clovinv && p->return_clover_inverse) {
    const bool twisted = p->dslash_type == QUDA_TWISTED_CLOVER_DSLASH;
    const double a = twisted ? 2.0 * p->kappa * p->mu : 0.0;
    if (twisted) G.clover.inverse_to_host(h_clovinv, hp, a * a);
    else QB_ERROR("loadCloverQuda: return_clover_inverse is implemented for twisted clover only");
  }
}
void freeCloverQuda(void) { G.clover.release(); }
// interface_quda.cpp invertMultiSrcQuda: param->num_src right-hand sides with the same operator, solver and preconditioner.
void invertMultiShiftQuda(void **, void *, QudaInvertParam *) { QB_ERROR("invertMultiShiftQuda is outside this build's scope"); }
void cloverQuda(void *, void *, QudaInvertParam *, QudaParity *, int) { QB_ERROR("cloverQuda is outside this build's scope"); }

}  // extern "C"

// =================================================================================================
// Solvers and multigrid
// =================================================================================================
namespace {

InverterType to_inverter(QudaInverterType t) {
  switch (t) {
    case QUDA_CG_INVERTER: return INV_CG;
    case QUDA_GCR_INVERTER: return INV_GCR;
    case QUDA_MR_INVERTER: return INV_MR;
    case QUDA_BICGSTAB_INVERTER: return INV_BICGSTAB;
    case QUDA_MG_INVERTER: return INV_MG;
    case QUDA_INVALID_INVERTER: return INV_NONE;
    default: QB_ERROR("Invalid solver type %d (this build provides CG, GCR, MR, BiCGStab and the MG preconditioner)", (int)t);
  }
}

// SolverParam(QudaInvertParam&) of include/invert_quda.h:201-240
void fill_solver_param(SolverParam &s, const QudaInvertParam *p) {
  s.inv_type = to_inverter(p->inv_type);
  s.inv_type_precondition = to_inverter(p->inv_type_precondition);
  if (p->tol == INVALID_DOUBLE) QB_ERROR("Parameter tol undefined");
  if (p->maxiter == INVALID_INT) QB_ERROR("Parameter maxiter undefined");
  s.tol = p->tol;
  s.maxiter = p->maxiter;
  s.Nkrylov = p->gcrNkrylov == INVALID_INT ? 20 : p->gcrNkrylov;
  s.delta = p->reliable_delta == INVALID_DOUBLE ? 1e-3 : p->reliable_delta;
  s.omega = p->omega == INVALID_DOUBLE ? 1.0 : p->omega;
  s.precision = to_prec(p->cuda_prec, "cuda_prec");
  s.precision_sloppy = p->cuda_prec_sloppy == QUDA_INVALID_PRECISION ? s.precision : to_prec(p->cuda_prec_sloppy, "cuda_prec_sloppy");
  s.precision_precondition = p->cuda_prec_precondition == QUDA_INVALID_PRECISION ? s.precision_sloppy : to_prec(p->cuda_prec_precondition, "cuda_prec_precondition");
  s.use_init_guess = p->use_init_guess == QUDA_USE_INIT_GUESS_YES;
  s.preserve_source = true;
  s.is_preconditioner = false;
  s.global_reduction = true;
  s.compute_true_res = true;
  s.pipeline = p->pipeline;
  s.precondition_cycle = p->precondition_cycle;
  s.max_res_increase = p->max_res_increase;
  s.max_res_increase_total = p->max_res_increase_total;
  s.verbosity = (int)p->verbosity == INVALID_INT ? 1 : (int)p->verbosity;
  if (s.inv_type_precondition == INV_MR) {
    if (p->maxiter_precondition != INVALID_INT) {
      // the inner MR takes its own iteration count; GCR copies `param` for it and overrides below
    }
  }
}

}  // namespace

extern "C" {

// interface_quda.cpp:2276-2543
void invertQuda(void *hp_x, void *hp_b, QudaInvertParam *param) {
  require_gauge();
  if (!param) QB_ERROR("invertQuda: null parameter struct");
  check_invert_param_operator(param);
  Runtime &r = rt();
  const int saved_verbosity = r.verbosity;
  if ((int)param->verbosity != INVALID_INT) r.verbosity = (int)param->verbosity;
  if (param->solution_type == QUDA_INVALID_SOLUTION) QB_ERROR("Parameter solution_type undefined");
  if (param->solve_type == QUDA_INVALID_SOLVE) QB_ERROR("Parameter solve_type undefined");
  const bool pc_solution = param->solution_type == QUDA_MATPC_SOLUTION || param->solution_type == QUDA_MATPCDAG_MATPC_SOLUTION;
  const bool pc_solve = param->solve_type == QUDA_DIRECT_PC_SOLVE || param->solve_type == QUDA_NORMOP_PC_SOLVE;
  const bool mat_solution = param->solution_type == QUDA_MAT_SOLUTION || param->solution_type == QUDA_MATPC_SOLUTION;
  const bool direct_solve = param->solve_type == QUDA_DIRECT_SOLVE || param->solve_type == QUDA_DIRECT_PC_SOLVE;
  if (param->solve_type != QUDA_DIRECT_SOLVE && param->solve_type != QUDA_DIRECT_PC_SOLVE && param->solve_type != QUDA_NORMOP_SOLVE &&
      param->solve_type != QUDA_NORMOP_PC_SOLVE)
    QB_ERROR("solve_type %d not supported", (int)param->solve_type);
  if (pc_solution && !pc_solve) QB_ERROR("Preconditioned (PC) solution_type requires a PC solve_type");
  if (!mat_solution && !pc_solution && pc_solve) QB_ERROR("Unpreconditioned MATDAG_MAT solution_type requires an unpreconditioned solve_type");
  if (param->inv_type_precondition == QUDA_MG_INVERTER && (!direct_solve || !mat_solution)) QB_ERROR("Multigrid preconditioning only supported for direct solves");
  if (param->inv_type_precondition == QUDA_MG_INVERTER && pc_solve) {
    // the cycle then runs on single-parity fields of the even-odd system: the hierarchy must have been built for it
    // (coarse_grid_solution_type = QUDA_MATPC_SOLUTION, same symmetric matpc_type; multigrid.cpp:494-505)
    if (!param->preconditioner) QB_ERROR("inv_type_precondition is QUDA_MG_INVERTER but `preconditioner` is not set (call newMultigridQuda first)");
    const MultigridSolver *ms = (const MultigridSolver *)param->preconditioner;
    if (!ms->mp.level[0].coarse_pc) QB_ERROR("Unsupported solution type combination: an even-odd preconditioned outer solve needs a multigrid built with coarse_grid_solution_type = QUDA_MATPC_SOLUTION");
    if ((int)param->matpc_type != (int)QUDA_MATPC_EVEN_EVEN && (int)param->matpc_type != (int)QUDA_MATPC_ODD_ODD) QB_ERROR("Multigrid on the even-odd system needs a symmetric matpc_type");
    if (ms->diracSmooth->matpc() != (int)param->matpc_type) QB_ERROR("matpc_type of the solve (%d) differs from the one the multigrid was built with", (int)param->matpc_type);
  }
  if (!mat_solution && direct_solve) QB_ERROR("Two-pass MATDAG_MAT solves with a direct solver are not implemented; use a NORMOP solve_type");

  // stopping criterion: the L2 relative residual only (invert_quda.h:201-240 maps residual_type onto the solvers' convergence test);
  // anything else would silently change the stopping semantics, so it is refused
  if ((int)param->residual_type != (int)QUDA_L2_RELATIVE_RESIDUAL && (int)param->residual_type != INVALID_INT)
    QB_ERROR("residual_type %d is not supported: this build stops on QUDA_L2_RELATIVE_RESIDUAL only (no heavy-quark or absolute residual)", (int)param->residual_type);
  SolverParam sp;
  fill_solver_param(sp, param);
  param->secs = 0; param->gflops = 0; param->iter = 0;
  const Prec prec = sp.precision;
  // int16 solver vectors exist for single fields; a doublet solve keeps fp32 vectors in front of a half-precision operator
  struct HalfVecGuard { bool saved; HalfVecGuard(bool off) : saved(half_vectors_enabled()) { if (off) set_half_vectors(false); } ~HalfVecGuard() { set_half_vectors(saved); } } hv_guard(nflavor_of(param) == 2);
  const Prec prec_vec_sloppy = blas_prec(sp.precision_sloppy);
  if (prec == PREC_HALF) QB_ERROR("cuda_prec must be single or double for a solve");
  param->spinorGiB = (double)G.lat.geom.Vh * nflavor_of(param) * 24 * (pc_solve ? 1 : 2) * (int)prec * (param->preserve_source == QUDA_PRESERVE_SOURCE_NO ? 7 : 9) / (double)(1 << 30);

  // createDirac (interface_quda.cpp:1386-1410): precise, sloppy and preconditioner operators
  std::unique_ptr<DiracTM> d(make_dirac(param, pc_solve, pick_gauge(prec)));
  std::unique_ptr<DiracTM> dS(make_dirac(param, pc_solve, pick_gauge(sp.precision_sloppy)));
  std::unique_ptr<DiracTM> dP(make_dirac(param, pc_solve, pick_gauge(sp.precision_precondition)));
  if (dS->gauge->prec != prec_vec_sloppy) dS->gauge_vec = pick_gauge(prec_vec_sloppy);
  if (dP->gauge->prec != prec_vec_sloppy) dP->gauge_vec = pick_gauge(prec_vec_sloppy);

  const int nfl = nflavor_of(param);
  if (nfl == 2 && param->inv_type_precondition == QUDA_MG_INVERTER) QB_ERROR("Multigrid for the non-degenerate doublet is not implemented");

  std::unique_ptr<SpinorField> b(new SpinorField(G.lat.geom.Vh, pc_solution ? 1 : 2, prec, 4, 3, 1, nfl));
  std::unique_ptr<SpinorField> x(new SpinorField(G.lat.geom.Vh, pc_solution ? 1 : 2, prec, 4, 3, 1, nfl));
  load_host_spinor(*b, hp_b, param);
  if (param->use_init_guess == QUDA_USE_INIT_GUESS_YES) load_host_spinor(*x, hp_x, param);
  else blas::zero(*x);
  double nb = blas::norm2(*b);
  if (nb == 0.0) QB_ERROR("Source has zero norm");
  if (param->solver_normalization == QUDA_SOURCE_NORMALIZATION) { blas::ax(1.0 / sqrt(nb), *b); blas::ax(1.0 / sqrt(nb), *x); }
  // massRescale (interface_quda.cpp:1412-1494)
  {
    const double kappa = param->kappa;
    double s = 1.0;
    switch (param->solution_type) {
      case QUDA_MAT_SOLUTION:
        if (param->mass_normalization == QUDA_MASS_NORMALIZATION || param->mass_normalization == QUDA_ASYMMETRIC_MASS_NORMALIZATION) s = 2.0 * kappa;
        break;
      case QUDA_MATDAG_MAT_SOLUTION:
        if (param->mass_normalization == QUDA_MASS_NORMALIZATION || param->mass_normalization == QUDA_ASYMMETRIC_MASS_NORMALIZATION) s = 4.0 * kappa * kappa;
        break;
      case QUDA_MATPC_SOLUTION:
        if (param->mass_normalization == QUDA_MASS_NORMALIZATION) s = 4.0 * kappa * kappa;
        else if (param->mass_normalization == QUDA_ASYMMETRIC_MASS_NORMALIZATION) s = 2.0 * kappa;
        break;
      case QUDA_MATPCDAG_MATPC_SOLUTION:
        if (param->mass_normalization == QUDA_MASS_NORMALIZATION) s = 16.0 * pow(kappa, 4);
        else if (param->mass_normalization == QUDA_ASYMMETRIC_MASS_NORMALIZATION) s = 4.0 * kappa * kappa;
        break;
      default: QB_ERROR("Solution type %d not supported", (int)param->solution_type);
    }
    if (s != 1.0) blas::ax(s, *b);
  }

  SpinorField in, out;
  const SolutionType st = (SolutionType)(int)param->solution_type;
  d->prepare(in, out, *x, *b, st);

  Solver *K = nullptr;
  if (param->inv_type_precondition == QUDA_MG_INVERTER) {
    if (!param->preconditioner) QB_ERROR("inv_type_precondition is QUDA_MG_INVERTER but `preconditioner` is not set (call newMultigridQuda first)");
    K = ((MultigridSolver *)param->preconditioner)->mg.get();
  }
  std::unique_ptr<SpinorField> tmp_in;
  if (mat_solution && !direct_solve) {  // normal equations: b' = A^dag b
    tmp_in.reset(new SpinorField(in.Vh, in.nparity, in.prec, 4, 3, 1, in.nflavor));
    blas::copy(*tmp_in, in);
    d->Mdag(in, *tmp_in);
  }
  {
    const bool normal = !direct_solve;
    DiracMatrix m(d.get(), normal), mS(dS.get(), normal), mP(dP.get(), normal);
    if (sp.inv_type_precondition == INV_MR && param->maxiter_precondition != INVALID_INT) {
      // GCR builds its inner MR from a copy of sp; hand the inner iteration count through `maxiter` of that copy
    }
    std::unique_ptr<Solver> solve(Solver::create(sp, m, mS, mP, K));
    if (GCR *g = dynamic_cast<GCR *>(solve.get())) g->set_inner(param->maxiter_precondition == INVALID_INT ? 10 : param->maxiter_precondition,
                                                               param->tol_precondition == INVALID_DOUBLE ? 0.1 : param->tol_precondition);
    (*solve)(out, in);
  }
  if (K && getenv("QUDA_B200_MG_PROFILE")) ((MultigridSolver *)param->preconditioner)->mg->print_profile();
  d->reconstruct(*x, *b, st);
  if (param->solver_normalization == QUDA_SOURCE_NORMALIZATION) blas::ax(sqrt(nb), *x);
  // interface_quda.cpp:2493-2508: make_resident_solution keeps the solution on the device instead of copying it to h_x; it replaces the
  // previous resident solution and is handed out by residentSolutionQudaB200 (the reference's consumers of use_resident_solution are the
  // force routines, interface_quda.cpp:4107-4146, 4930-4985, outside this build; invertQuda itself never reads use_resident_solution)
  if (param->make_resident_solution == 1) {
    QB_CUDA(cudaStreamSynchronize(r.compute));
    G.solution_resident = std::move(x);
  } else {
    save_host_spinor(hp_x, *x, param);
  }

  param->true_res = sp.true_res;
  // the heavy-quark residual is not computed (residual_type other than L2 is refused above): NaN, not a value that reads as converged
  param->true_res_hq = std::numeric_limits<double>::quiet_NaN();
  param->iter += sp.iter;
  param->secs += sp.secs;
  const double gflops = (double)(d->flops + dS->flops + dP->flops + (double)blas::flops) * 1e-9;
  param->gflops += gflops;
  blas::flops = 0;
  r.verbosity = saved_verbosity;
}

// Block path of invertMultiSrcQuda: GCR + multigrid on a direct, unpreconditioned solve.  All sources advance in
// lock-step; the coarse levels of the K-cycle run on block fields through the multi-RHS tensor-core operator (block_solver.cu), so
// the coarse links are read once per operator application for all sources.  Returns false when the request does not qualify.
static bool invert_multi_src_block(void **hp_x, void **hp_b, QudaInvertParam *param) {
  if (param->num_src < 2 || param->inv_type != QUDA_GCR_INVERTER || param->inv_type_precondition != QUDA_MG_INVERTER || !param->preconditioner) return false;
  if ((param->solve_type != QUDA_DIRECT_SOLVE && param->solve_type != QUDA_DIRECT_PC_SOLVE) || param->solution_type != QUDA_MAT_SOLUTION ||
      param->use_init_guess == QUDA_USE_INIT_GUESS_YES)
    return false;
  const int mode = getenv("QB_BLOCK_MG_MODE") ? atoi(getenv("QB_BLOCK_MG_MODE")) : 3;
  MultigridSolver *ms = (MultigridSolver *)param->preconditioner;
  MG &mg = *ms->mg;
  // even-odd outer solve (the reference's default): the hierarchy must be coarsened on the even-odd system of the same symmetric matpc_type
  const bool pc = param->solve_type == QUDA_DIRECT_PC_SOLVE;
  if (pc && (!mg.pc_coarsen || ((int)param->matpc_type != (int)QUDA_MATPC_EVEN_EVEN && (int)param->matpc_type != (int)QUDA_MATPC_ODD_ODD) ||
             ms->diracSmooth->matpc() != (int)param->matpc_type))
    return false;   // invertQuda reports what is wrong
  SolverParam sp;
  fill_solver_param(sp, param);
  const Prec prec = sp.precision;
  if (prec == PREC_HALF) QB_ERROR("cuda_prec must be single or double for a solve");
  // sources per block: as many as the multi-RHS kernel takes, and as the Krylov space (2 x min(Nkrylov, 16) + 6 single-precision fields
  // per source, allocated as the iteration proceeds) leaves room for
  size_t free_b = 0, total_b = 0;
  QB_CUDA(cudaMemGetInfo(&free_b, &total_b));
  const int npar = pc ? 1 : 2;
  const size_t per_src = (size_t)G.lat.geom.Vh * npar * 96 * (2 * std::min(sp.Nkrylov, 16) + 6) + (size_t)G.lat.geom.Vh * 2 * 24 * (int)prec * 4;
  int Rblk = std::min<int>(param->num_src, (int)std::max<size_t>(1, (size_t)(0.7 * (double)(free_b + pool_cached_bytes())) / per_src));
  if (getenv("QB_BLOCK_MG_R")) Rblk = std::min(Rblk, atoi(getenv("QB_BLOCK_MG_R")));
  while (Rblk >= 2 && !block_mg_supported(mg, Rblk, mode)) Rblk--;
  if (Rblk < 2) return false;

  check_invert_param_operator(param);
  Runtime &r = rt();
  const int saved_verbosity = r.verbosity;
  if ((int)param->verbosity != INVALID_INT) r.verbosity = (int)param->verbosity;
  const Prec prec_vec_sloppy = sp.precision_sloppy == PREC_HALF ? PREC_SINGLE : sp.precision_sloppy;   // the lock-step Krylov space is fp32
  std::unique_ptr<DiracTM> d(make_dirac(param, pc, pick_gauge(prec)));
  std::unique_ptr<DiracTM> dS(make_dirac(param, pc, pick_gauge(sp.precision_sloppy)));
  if (dS->gauge->prec != prec_vec_sloppy) dS->gauge_vec = pick_gauge(prec_vec_sloppy);
  DiracMatrix m(d.get(), false), mS(dS.get(), false);
  param->secs = 0; param->gflops = 0; param->iter = 0; param->true_res = 0; param->true_res_hq = std::numeric_limits<double>::quiet_NaN();
  double gflops_single = 0.0;   // left-over single sources go through invertQuda, which resets the flop counters
  auto now = [&]() { QB_CUDA(cudaStreamSynchronize(r.compute)); return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  for (int first = 0; first < param->num_src; first += Rblk) {
    const int R = std::min(Rblk, param->num_src - first);
    if (R == 1) {  // a single left-over source: the ordinary path, its counters added to the block's
      const double secs = param->secs, tr0 = param->true_res; const int it0 = param->iter;
      const unsigned long long bf = blas::flops;
      invertQuda(hp_x[first], hp_b[first], param);
      gflops_single += param->gflops;
      blas::flops = bf;
      param->secs += secs; param->iter += it0; param->true_res = std::max(param->true_res, tr0);
      continue;
    }
    std::vector<std::unique_ptr<SpinorField>> bs(R), xs(R);
    std::vector<SpinorField> vin(R), vout(R);   // what the solver works on: the fields themselves, or the even-odd source / solution views
    std::vector<SpinorField *> pb(R), px(R);
    std::vector<double> nb(R);
    for (int c = 0; c < R; c++) {
      bs[c].reset(new SpinorField(G.lat.geom.Vh, 2, prec)); xs[c].reset(new SpinorField(G.lat.geom.Vh, 2, prec));
      load_host_spinor(*bs[c], hp_b[first + c], param);
      blas::zero(*xs[c]);
      nb[c] = blas::norm2(*bs[c]);
      if (nb[c] == 0.0) QB_ERROR("Source %d has zero norm", first + c);
      if (param->solver_normalization == QUDA_SOURCE_NORMALIZATION) blas::ax(1.0 / sqrt(nb[c]), *bs[c]);
      if (param->mass_normalization == QUDA_MASS_NORMALIZATION || param->mass_normalization == QUDA_ASYMMETRIC_MASS_NORMALIZATION)
        blas::ax(2.0 * param->kappa, *bs[c]);   // massRescale, MAT solution (interface_quda.cpp:1412-1494)
      d->prepare(vin[c], vout[c], *xs[c], *bs[c], SOL_MAT);   // full solve: the fields themselves
      pb[c] = &vin[c]; px[c] = &vout[c];
    }
    std::vector<double> tr;
    const double t0 = now();   // secs counts the solve only, like invertQuda (sources already on the device)
    const int it = block_mg_gcr_solve(mg, m, mS, px, pb, sp, mode, tr);
    param->secs += now() - t0;
    for (int c = 0; c < R; c++) {
      d->reconstruct(*xs[c], *bs[c], SOL_MAT);
      if (param->solver_normalization == QUDA_SOURCE_NORMALIZATION) blas::ax(sqrt(nb[c]), *xs[c]);
      save_host_spinor(hp_x[first + c], *xs[c], param);
      param->true_res = std::max(param->true_res, tr[c]);
    }
    param->iter += it;
    log_msg(1, "invertMultiSrcQuda: block of %d sources, %d lock-step GCR iterations, worst true residual %e\n", R, it, param->true_res);
  }
  QB_CUDA(cudaStreamSynchronize(r.compute));
  param->gflops = gflops_single + (double)(d->flops + dS->flops + (double)blas::flops) * 1e-9;
  blas::flops = 0;
  r.verbosity = saved_verbosity;
  return true;
}

// interface_quda.cpp: invertMultiSrcQuda (quda.h:647).  GCR + multigrid direct solves take the block path above; everything else
// runs the sources one after the other through invertQuda (gauge field, clover term and multigrid hierarchy stay resident);
// iter / secs / gflops accumulate, true_res reports the worst source.
void invertMultiSrcQuda(void **hp_x, void **hp_b, QudaInvertParam *param) {
  if (param->num_src < 1 || param->num_src == INVALID_INT) QB_ERROR("invertMultiSrcQuda: num_src undefined");
  require_gauge();
  if (invert_multi_src_block(hp_x, hp_b, param)) return;
  int iter = 0;
  double secs = 0, gflops = 0, worst = 0, worst_hq = 0;
  for (int i = 0; i < param->num_src; i++) {
    invertQuda(hp_x[i], hp_b[i], param);
    iter += param->iter; secs += param->secs; gflops += param->gflops;
    worst = std::max(worst, param->true_res); worst_hq = std::max(worst_hq, param->true_res_hq);
  }
  param->iter = iter; param->secs = secs; param->gflops = gflops; param->true_res = worst; param->true_res_hq = worst_hq;
}

// interface_quda.cpp:2161-2269
void *newMultigridQuda(QudaMultigridParam *mgp) {
  require_gauge();
  if (!mgp || !mgp->invert_param) QB_ERROR("newMultigridQuda: null parameter struct");
  QudaInvertParam *ip = mgp->invert_param;
  check_invert_param_operator(ip);
  if (nflavor_of(ip) == 2) QB_ERROR("Multigrid for the non-degenerate doublet is not implemented");
  if (mgp->n_level == INVALID_INT) QB_ERROR("Parameter n_level undefined");
  if (mgp->n_level < 2 || mgp->n_level > QUDA_MAX_MG_LEVEL) QB_ERROR("Maximum number of multigrid levels is %d (and at least 2), requested %d", QUDA_MAX_MG_LEVEL, mgp->n_level);
  if (ip->solve_type != QUDA_DIRECT_SOLVE) QB_ERROR("Outer MG solver can only use QUDA_DIRECT_SOLVE at present");
  Runtime &r = rt();
  const int saved_verbosity = r.verbosity;
  if ((int)ip->verbosity != INVALID_INT) r.verbosity = (int)ip->verbosity;
  const double t0 = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();

  long alloc_calls0 = 0;
  const double alloc_t0 = pool_driver_time(&alloc_calls0);
  MultigridSolver *ms = new MultigridSolver();
  MGParam &mp = ms->mp;
  mp.n_level = mgp->n_level;
  for (int l = 0; l < mp.n_level; l++) {
    MGLevelParam &lp = mp.level[l];
    if (mgp->smoother_solve_type[l] != QUDA_DIRECT_SOLVE && mgp->smoother_solve_type[l] != QUDA_DIRECT_PC_SOLVE)
      QB_ERROR("Unsupported smoother solve type %d on level %d", (int)mgp->smoother_solve_type[l], l);
    if (mgp->coarse_grid_solution_type[l] != QUDA_MATPC_SOLUTION && mgp->coarse_grid_solution_type[l] != QUDA_MAT_SOLUTION && l < mgp->n_level - 1)
      QB_ERROR("coarse_grid_solution_type[%d] must be QUDA_MAT_SOLUTION or QUDA_MATPC_SOLUTION", l);
    // single-parity injection into the coarse grid = coarsening of the even-odd preconditioned operator (multigrid.cpp:145-155)
    lp.coarse_pc = mgp->coarse_grid_solution_type[l] == QUDA_MATPC_SOLUTION;
    if (lp.coarse_pc && mgp->smoother_solve_type[l] != QUDA_DIRECT_PC_SOLVE)
      QB_ERROR("For this coarse grid solution type, a preconditioned smoother is required (level %d)", l);
    for (int d = 0; d < 4; d++) {
      lp.geo_bs[d] = mgp->geo_block_size[l][d];
      if (l < mp.n_level - 1 && (lp.geo_bs[d] == INVALID_INT || lp.geo_bs[d] < 1)) QB_ERROR("Parameter geo_block_size[%d][%d] undefined", l, d);
    }
    lp.spin_bs = l == 0 ? 2 : 1;
    if (l < mp.n_level - 1 && mgp->spin_block_size[l] != lp.spin_bs) QB_ERROR("spin_block_size[%d] must be %d", l, lp.spin_bs);
    lp.nvec = mgp->n_vec[l];
    if (l < mp.n_level - 1 && (lp.nvec == INVALID_INT || lp.nvec < 2)) QB_ERROR("Parameter n_vec[%d] undefined", l);
    lp.smoother = to_inverter(mgp->smoother[l]);
    if (lp.smoother != INV_MR && lp.smoother != INV_GCR) QB_ERROR("smoother[%d] must be MR or GCR", l);
    lp.smoother_pc = mgp->smoother_solve_type[l] == QUDA_DIRECT_PC_SOLVE;
    lp.nu_pre = mgp->nu_pre[l]; lp.nu_post = mgp->nu_post[l];
    if (lp.nu_pre == INVALID_INT || lp.nu_post == INVALID_INT) QB_ERROR("Parameter nu_pre/nu_post[%d] undefined", l);
    lp.smoother_tol = mgp->smoother_tol[l];
    lp.omega = mgp->omega[l] == INVALID_DOUBLE ? 1.0 : mgp->omega[l];
    lp.recursive = mgp->cycle_type[l] == QUDA_MG_CYCLE_RECURSIVE;
    if (mgp->cycle_type[l] != QUDA_MG_CYCLE_RECURSIVE && mgp->cycle_type[l] != QUDA_MG_CYCLE_VCYCLE) QB_ERROR("Multigrid cycle type %d not supported", (int)mgp->cycle_type[l]);
    lp.global_reduction = mgp->global_reduction[l] != QUDA_BOOLEAN_NO;
  }
  mp.setup_maxiter = mgp->setup_maxiter == INVALID_INT ? 500 : mgp->setup_maxiter;
  mp.setup_tol = mgp->setup_tol == INVALID_DOUBLE ? 5e-6 : mgp->setup_tol;
  mp.compute_null_vector = mgp->compute_null_vector == QUDA_COMPUTE_NULL_VECTOR_YES;
  mp.generate_all_levels = mgp->generate_all_levels == QUDA_BOOLEAN_YES;
  mp.verbosity = r.verbosity;
  mp.keep_null_vectors = mgp->run_verify == QUDA_BOOLEAN_YES;  // needed by mgVerifyQudaB200 / mgNullVectorQudaB200 only
  mp.half_storage = ip->cuda_prec_precondition == QUDA_HALF_PRECISION;
  if (getenv("QB_MG_HALF_STORAGE")) mp.half_storage = atoi(getenv("QB_MG_HALF_STORAGE")) != 0;
  mp.vec_infile = std::string(mgp->vec_infile, strnlen(mgp->vec_infile, sizeof(mgp->vec_infile)));
  mp.vec_outfile = std::string(mgp->vec_outfile, strnlen(mgp->vec_outfile, sizeof(mgp->vec_outfile)));
  if (!mp.compute_null_vector && mp.vec_infile.empty()) QB_ERROR("compute_null_vector = NO needs vec_infile (written by an earlier newMultigridQuda with vec_outfile set)");

  // fine operators: residual = full operator in the MG working precision (fp32 vectors); smoother = even-odd
  // preconditioned operator in cuda_prec_precondition, with the setup rescale of kappa / mu (interface_quda.cpp:2196-2233)
  const double dk = mgp->delta_kappaPR == 0.0 ? 1.0 : mgp->delta_kappaPR, dm = mgp->delta_muPR == 0.0 ? 1.0 : mgp->delta_muPR;
  const GaugeField *g32 = pick_gauge(PREC_SINGLE);
  ms->dirac.reset(make_dirac(ip, false, g32, dk, dm));
  const Prec pprec = ip->cuda_prec_precondition == QUDA_INVALID_PRECISION ? PREC_SINGLE : to_prec(ip->cuda_prec_precondition, "cuda_prec_precondition");
  ms->diracSmooth.reset(make_dirac(ip, mp.level[0].smoother_pc, pprec == PREC_DOUBLE ? g32 : pick_gauge(pprec), dk, dm));
  ms->diracSmooth->gauge_vec = g32;
  ms->dirac->dagger = false; ms->diracSmooth->dagger = false;
  ms->mg.reset(new MG(mp, 0, ms->dirac.get(), ms->diracSmooth.get(), nullptr));
  for (int l = 0; l < mp.n_level - 1; l++)
    for (int d = 0; d < 4; d++) mgp->geo_block_size[l][d] = mp.level[l].geo_bs[d];
  QB_CUDA(cudaDeviceSynchronize());
  // the setup's multi-GB scratch (site-major V, decompressed links) sits in the allocator's cache: hand it back to the driver only when it
  // is a sizeable part of the device (cudaFree / cudaMalloc of such blocks cost up to 0.5 s per setup, and the QKXTM drivers set up twice)
  {
    size_t free_b = 0, total_b = 0;
    QB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    if (free_b < total_b / 4) pool_release_all();
  }
  {
    long calls = 0;
    const double t = pool_driver_time(&calls);
    log_msg(1, "newMultigridQuda: %.3f s of the setup inside cudaMalloc / cudaFree (%ld calls of the caching allocator)\n", t - alloc_t0, calls - alloc_calls0);
  }
  mgp->secs = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count() - t0;
  mgp->gflops = 0;
  r.verbosity = saved_verbosity;
  return ms;
}

void destroyMultigridQuda(void *mg) {
  if (!mg) return;
  QB_CUDA(cudaDeviceSynchronize());
  delete (MultigridSolver *)mg;
}

// ---- multigrid introspection for tests (include/quda_b200_ext.h) ----------------------------------
static MG *mg_level(void *mg, int level) {
  if (!mg) QB_ERROR("null multigrid handle");
  MG *m = ((MultigridSolver *)mg)->mg.get();
  for (int l = 0; l < level; l++) {
    if (!m->coarse) QB_ERROR("multigrid has no level %d", level);
    m = m->coarse.get();
  }
  return m;
}

void mgVerifyQudaB200(void *mg, int level, double *dev3) { mg_level(mg, level)->verify(dev3); }

// Wall-clock profile of the cycle per level: on != 0 switches the (stream-synchronising) section timers on.  mgProfileGetQudaB200 returns
// and resets the accumulated seconds of level `level`: t6 = {pre-smooth (or the coarsest-grid solve), residual, restrict, coarse solve
// (everything below), prolong, post-smooth}, *ncycle = cycles counted.
void mgProfileEnableQudaB200(int on) { mg_profile_enable(on != 0); }
void mgProfileGetQudaB200(void *mg, int level, double *t6, long *ncycle) {
  MG *m = mg_level(mg, level);
  for (int i = 0; i < 6; i++) { t6[i] = m->t_prof[i]; m->t_prof[i] = 0; }
  if (ncycle) *ncycle = m->ncycle;
  m->ncycle = 0;
}

void mgLevelInfoQudaB200(void *mg, int level, int *info8) {
  MG *m = mg_level(mg, level);
  if (!m->transfer) QB_ERROR("level %d is the coarsest level: it has no transfer operator", level);
  const Transfer &T = *m->transfer;
  for (int d = 0; d < 4; d++) info8[d] = T.coarse.X[d];
  info8[4] = T.nvec; info8[5] = T.Nf; info8[6] = T.block_sites; info8[7] = 2 * T.nvec;
}

}  // extern "C"

namespace qb {
void import_generic(SpinorField &f, const float *h, cudaStream_t s);
void export_generic(float *h, const SpinorField &f, cudaStream_t s);
}

extern "C" {

// host order of generic fields: [parity][cb][component k][re,im], float32
void mgProlongQudaB200(void *mg, int level, float *h_fine_out, const float *h_coarse_in) {
  MG *m = mg_level(mg, level);
  std::unique_ptr<SpinorField> c(m->transfer->new_coarse_field());
  std::unique_ptr<SpinorField> f(new SpinorField(m->transfer->fine.Vh, 2, PREC_SINGLE, m->transfer->fine_nspin, m->transfer->fine_ncolor));
  import_generic(*c, h_coarse_in, rt().compute);
  m->transfer->P(*f, *c);
  export_generic(h_fine_out, *f, rt().compute);
}
void mgRestrictQudaB200(void *mg, int level, float *h_coarse_out, const float *h_fine_in) {
  MG *m = mg_level(mg, level);
  std::unique_ptr<SpinorField> c(m->transfer->new_coarse_field());
  std::unique_ptr<SpinorField> f(new SpinorField(m->transfer->fine.Vh, 2, PREC_SINGLE, m->transfer->fine_nspin, m->transfer->fine_ncolor));
  import_generic(*f, h_fine_in, rt().compute);
  m->transfer->R(*c, *f);
  export_generic(h_coarse_out, *c, rt().compute);
}
// applies the operator of level `level` (0: fine full operator in fp32, >= 1: coarse operator); pc != 0: the smoother's operator
void mgMatQudaB200(void *mg, int level, int pc, float *h_out, const float *h_in) {
  MG *m = mg_level(mg, level);
  const Dirac *d = pc ? m->matSmooth : m->matResidual;
  std::unique_ptr<SpinorField> in(d->new_field(PREC_SINGLE)), out(d->new_field(PREC_SINGLE));
  import_generic(*in, h_in, rt().compute);
  d->M(*out, *in);
  export_generic(h_out, *out, rt().compute);
}
// Link matrices of the coarse operator ON level `level` >= 1 (built by level - 1), row-major: out[site][d][row][col][re, im], site = full
// index (parity * Vh + x_cb), d = 0..7 the hop to x + e_d (e_d = +mu for d = 2 mu, -mu for d = 2 mu + 1), d = 8 the site-diagonal block;
// which = 0: L (the -kappa of the reference's  X - kappa sum Y  folded in), 1: Xinv ([site][row][col]), 2: Yhat = Xinv L (slot 8 = 1).
// In the reference's terms (dslash_coarse.cu:49-203): Y_{mu+4}(x) = -L_{2mu}(x) / kappa, Y_mu(x) = -L_{2mu+1}(x + mu)^dag / kappa, X = L_8.
void mgCoarseLinksQudaB200(void *mg, int level, int which, float *h_out) {
  if (level < 1) QB_ERROR("mgCoarseLinksQudaB200: level must be >= 1");
  MG *m = mg_level(mg, level - 1);
  if (!m->coarse_op) QB_ERROR("multigrid has no level %d", level);
  const CoarseOperator &op = *m->coarse_op;
  const float *src = which == 0 ? op.Y : (which == 1 ? op.Xinv : op.Yhat);
  if (!src) QB_ERROR("mgCoarseLinksQudaB200: link field %d has not been computed on level %d", which, level);
  const int N = op.N, nd = which == 1 ? 1 : 9;
  const size_t nmat = (size_t)op.geom.V() * nd, per = (size_t)N * N * 2;
  std::vector<float> raw(nmat * per);
  QB_CUDA(cudaStreamSynchronize(rt().compute));
  QB_CUDA(cudaMemcpy(raw.data(), src, raw.size() * sizeof(float), cudaMemcpyDeviceToHost));
  // device layout of one matrix: [col][row pair] float4 = (M[2rp][col], M[2rp+1][col])
  for (size_t k = 0; k < nmat; k++)
    for (int c = 0; c < N; c++)
      for (int r = 0; r < N; r++) {
        const float *e = raw.data() + k * per + ((size_t)c * (N / 2) + (r >> 1)) * 4 + (r & 1) * 2;
        float *o = h_out + k * per + ((size_t)r * N + c) * 2;
        o[0] = e[0]; o[1] = e[1];
      }
}
// null vector k of level `level` in host order
void mgNullVectorQudaB200(void *mg, int level, int k, float *h_out) {
  MG *m = mg_level(mg, level);
  if (k < 0 || k >= (int)m->B.size()) QB_ERROR("null vector index %d out of range", k);
  export_generic(h_out, *m->B[k], rt().compute);
}
// mean device time (ms, CUDA events on the compute stream) of `niter` applications of a level's operator / transfer:
// what = 0: full operator M, 1: smoother operator, 2: prolongator, 3: restrictor (of the transfer *from* this level)
double mgTimeQudaB200(void *mg, int level, int what, int niter) {
  MG *m = mg_level(mg, level);
  Runtime &r = rt();
  const Dirac *d = what == 1 ? m->matSmooth : m->matResidual;
  std::unique_ptr<SpinorField> in(d->new_field(PREC_SINGLE)), out(d->new_field(PREC_SINGLE)), c;
  random_fill(*in, 99);
  if (what >= 2) {
    if (!m->transfer) QB_ERROR("level %d has no transfer operator", level);
    c.reset(m->transfer->new_coarse_field());
    random_fill(*c, 98);
  }
  auto apply = [&]() {
    if (what <= 1) d->M(*out, *in);
    else if (what == 2) m->transfer->P(*out, *c);
    else m->transfer->R(*c, *in);
  };
  for (int i = 0; i < 3; i++) apply();
  cudaEvent_t e0, e1;
  QB_CUDA(cudaEventCreate(&e0)); QB_CUDA(cudaEventCreate(&e1));
  QB_CUDA(cudaEventRecord(e0, r.compute));
  for (int i = 0; i < niter; i++) apply();
  QB_CUDA(cudaEventRecord(e1, r.compute));
  QB_CUDA(cudaEventSynchronize(e1));
  float ms = 0;
  QB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return (double)ms / niter;
}
// Multi-RHS tensor-core coarse operator on `nrhs` host vectors of a coarse level (level >= 1), h_in / h_out = [rhs][generic field].
//   what = 0: full operator M_c;  1: hopping term into the even sites (reads the odd sites);  2: Xinv on the odd sites
//   mode = 1: tf32, 3: split tf32 (fp32-accurate)
static DiracCoarse *coarse_dirac_of(MG *m, int level) {
  if (level < 1) QB_ERROR("the multi-RHS tensor-core operator exists on coarse levels only (level >= 1)");
  DiracCoarse *d = dynamic_cast<DiracCoarse *>(const_cast<Dirac *>(m->matResidual));
  if (!d) QB_ERROR("level %d has no coarse operator", level);
  return d;
}
static void mrhs_args(CoarseMrhsArgs &a, const CoarseOperator &op, CoarseBlockField &out, const CoarseBlockField &in, int what, int nrhs, int mode) {
  a = CoarseMrhsArgs{};
  a.op = &op; a.out = out.v; a.in_hop = in.v; a.in_diag = in.v; a.xpay = nullptr;
  const long pb = (long)in.parity_float4();
  for (int p = 0; p < 2; p++) a.out_poff[p] = a.hop_poff[p] = a.diag_poff[p] = a.xpay_poff[p] = p * pb;
  a.a = 1.f; a.b = 0.f; a.R = nrhs; a.mode = mode;
  if (what == 0) { a.parity = -1; a.use_y = true; a.use_x = true; }
  else if (what == 1) { a.parity = 0; a.use_y = true; }
  else if (what == 2) { a.parity = 1; a.use_xinv = true; }
  else QB_ERROR("mgMatMrhs: unknown operator %d", what);
}
void mgMatMrhsQudaB200(void *mg, int level, int what, int nrhs, int mode, float *h_out, const float *h_in) {
  MG *m = mg_level(mg, level);
  DiracCoarse *d = coarse_dirac_of(m, level);
  CoarseOperator &op = *d->op;
  if (what == 2 && !op.Xinv) op.compute_xinv();
  op.prepare_mrhs();
  std::vector<std::unique_ptr<SpinorField>> in(nrhs), out(nrhs);
  std::vector<SpinorField *> pin(nrhs), pout(nrhs);
  const size_t reals = (size_t)op.geom.V() * op.N * 2;
  for (int r = 0; r < nrhs; r++) {
    in[r].reset(new SpinorField(op.geom.Vh, 2, PREC_SINGLE, 2, op.nvec));
    out[r].reset(new SpinorField(op.geom.Vh, 2, PREC_SINGLE, 2, op.nvec));
    import_generic(*in[r], h_in + r * reals, rt().compute);
    out[r]->zero(rt().compute);
    pin[r] = in[r].get(); pout[r] = out[r].get();
  }
  CoarseBlockField bin(op.geom.Vh, 2, op.N, nrhs), bout(op.geom.Vh, 2, op.N, nrhs);
  bin.pack(pin.data());
  bout.pack(pout.data());
  CoarseMrhsArgs a;
  mrhs_args(a, op, bout, bin, what, nrhs, mode);
  coarse_apply_mrhs(a);
  bout.unpack(pout.data());
  for (int r = 0; r < nrhs; r++) export_generic(h_out + r * reals, *out[r], rt().compute);
}
int mgMrhsMaxRhsQudaB200(void *mg, int level, int mode) { return coarse_mrhs_max_rhs(coarse_dirac_of(mg_level(mg, level), level)->op->N, mode); }
// mean device time (ms) of `niter` applications of the multi-RHS operator (what as above) on resident random block fields
double mgTimeMrhsQudaB200(void *mg, int level, int what, int nrhs, int mode, int niter) {
  MG *m = mg_level(mg, level);
  DiracCoarse *d = coarse_dirac_of(m, level);
  CoarseOperator &op = *d->op;
  if (what == 2 && !op.Xinv) op.compute_xinv();
  op.prepare_mrhs();
  Runtime &r = rt();
  CoarseBlockField bin(op.geom.Vh, 2, op.N, nrhs), bout(op.geom.Vh, 2, op.N, nrhs);
  {
    SpinorField view;  // random numbers through the generic generator: treat the block as one long single-parity field
    view.prec = PREC_SINGLE; view.nparity = 1; view.ncomplex = 2; view.nspin = 1; view.ncolor = 2;
    view.Vh = (long)(bin.bytes() / 16); view.v = bin.v; view.parity_bytes = bin.bytes(); view.owner = false;
    random_fill(view, 4242);
  }
  CoarseMrhsArgs a;
  mrhs_args(a, op, bout, bin, what, nrhs, mode);
  for (int i = 0; i < 3; i++) coarse_apply_mrhs(a);
  cudaEvent_t e0, e1;
  QB_CUDA(cudaEventCreate(&e0)); QB_CUDA(cudaEventCreate(&e1));
  QB_CUDA(cudaEventRecord(e0, r.compute));
  for (int i = 0; i < niter; i++) coarse_apply_mrhs(a);
  QB_CUDA(cudaEventRecord(e1, r.compute));
  QB_CUDA(cudaEventSynchronize(e1));
  float ms = 0;
  QB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return (double)ms / niter;
}
// one multigrid cycle on level `level`: x = MG(b)
void mgCycleQudaB200(void *mg, int level, float *h_x, const float *h_b) {
  MG *m = mg_level(mg, level);
  std::unique_ptr<SpinorField> b(m->matResidual->new_field(PREC_SINGLE)), x(m->matResidual->new_field(PREC_SINGLE));
  import_generic(*b, h_b, rt().compute);
  (*m)(*x, *b);
  export_generic(h_x, *x, rt().compute);
}

}  // extern "C"

// ---- BLAS / reduction test hook (what tests/blas_test.cu drives inside the reference) ------------------------
// Operands are flat arrays of n complex numbers in precision `prec` (4 or 8 bytes per real), n % 12 == 0.
// coef = {a_re, a_im, b_re, b_im}; result receives up to 3 doubles.  Returns the number of result doubles.
extern "C" int blasQudaB200(const char *name, long n, int prec, const double *coef, void *x, void *y, void *z, void *w, double *result) {
  require_init();
  using namespace qb::blas;
  if (n <= 0 || n % 12) QB_ERROR("blasQudaB200: n must be a positive multiple of 12");
  const Prec pr = to_prec((QudaPrecision)prec, "prec");
  // half precision (int16 + norm fields): the host arrays are fp32; they go through an fp32 device field and the library's own
  // fp32 <-> int16 conversion (blas::copy), in the internal plane order [plane][site] of a parity field of n / 12 sites
  const bool half = pr == PREC_HALF;
  const long Vh = n / 12;
  const size_t bytes = (size_t)n * 2 * (half ? 4 : (int)pr);
  SpinorField fx(Vh, 1, pr), fy(Vh, 1, pr), fz(Vh, 1, pr), fw(Vh, 1, pr);
  SpinorField stage(Vh, 1, half ? PREC_SINGLE : pr);
  cudaStream_t s = rt().compute;
  void *hp[4] = {x, y, z, w};
  SpinorField *fp[4] = {&fx, &fy, &fz, &fw};
  for (int i = 0; i < 4; i++)
    if (hp[i]) {
      if (half) { QB_CUDA(cudaMemcpyAsync(stage.v, hp[i], bytes, cudaMemcpyHostToDevice, s)); copy(*fp[i], stage); }
      else QB_CUDA(cudaMemcpyAsync(fp[i]->v, hp[i], bytes, cudaMemcpyHostToDevice, s));
    }
  const double a = coef[0];
  const Complex ca(coef[0], coef[1]), cb(coef[2], coef[3]);
  const std::string nm(name);
  int nres = 0;
  if (nm == "ax") ax(a, fx);
  else if (nm == "axpy") axpy(a, fx, fy);
  else if (nm == "xpy") xpy(fx, fy);
  else if (nm == "xpay") xpay(fx, a, fy);
  else if (nm == "mxpy") mxpy(fx, fy);
  else if (nm == "axpby") axpby(a, fx, coef[2], fy);
  else if (nm == "caxpy") caxpy(ca, fx, fy);
  else if (nm == "caxpby") caxpby(ca, fx, cb, fy);
  else if (nm == "cxpaypbz") cxpaypbz(fx, ca, fy, cb, fz);
  else if (nm == "caxpbypz") caxpbypz(ca, fx, cb, fy, fz);
  else if (nm == "caxpbypzYmbw") caxpbypzYmbw(ca, fx, cb, fy, fz, fw);
  else if (nm == "cabxpyAx") cabxpyAx(a, cb, fx, fy);
  else if (nm == "caxpyXmaz") caxpyXmaz(ca, fx, fy, fz);
  else if (nm == "norm2") { result[0] = norm2(fx); nres = 1; }
  else if (nm == "reDotProduct") { result[0] = reDotProduct(fx, fy); nres = 1; }
  else if (nm == "cDotProduct") { Complex c = cDotProduct(fx, fy); result[0] = c.real(); result[1] = c.imag(); nres = 2; }
  else if (nm == "cDotProductNormA") { double3_ d = cDotProductNormA(fx, fy); result[0] = d.x; result[1] = d.y; result[2] = d.z; nres = 3; }
  else if (nm == "cDotProductNormB") { double3_ d = cDotProductNormB(fx, fy); result[0] = d.x; result[1] = d.y; result[2] = d.z; nres = 3; }
  else if (nm == "axpyNorm") { result[0] = axpyNorm(a, fx, fy); nres = 1; }
  else if (nm == "xmyNorm") { result[0] = xmyNorm(fx, fy); nres = 1; }
  else if (nm == "caxpyNorm") { result[0] = caxpyNorm(ca, fx, fy); nres = 1; }
  else if (nm == "cabxpyAxNorm") { result[0] = cabxpyAxNorm(a, cb, fx, fy); nres = 1; }
  else if (nm == "caxpyDotzy") { Complex c = caxpyDotzy(ca, fx, fy, fz); result[0] = c.real(); result[1] = c.imag(); nres = 2; }
  else if (nm == "caxpyXmazNormX") { result[0] = caxpyXmazNormX(ca, fx, fy, fz); nres = 1; }
  else if (nm == "xpaycDotzy") { Complex c = xpaycDotzy(fx, a, fy, fz); result[0] = c.real(); result[1] = c.imag(); nres = 2; }
  else if (nm == "bicgstabUpdate") {  // w += a x + b y; y -= b z; (<x, y>, |y|^2)
    double3_ d = bicgstabUpdate(ca, fx, cb, fy, fz, fw, fx); result[0] = d.x; result[1] = d.y; result[2] = d.z; nres = 3;
  } else if (nm == "block_cDotProduct") {  // (x,w), (y,w), (z,w)
    std::vector<SpinorField *> v{&fx, &fy, &fz};
    Complex r3[3];
    cDotProduct(r3, v, fw);
    for (int i = 0; i < 3; i++) { result[2 * i] = r3[i].real(); result[2 * i + 1] = r3[i].imag(); }
    nres = 6;
  } else if (nm == "block_caxpy") {  // w += a x + b y + conj(a) z
    std::vector<SpinorField *> v{&fx, &fy, &fz};
    Complex c3[3] = {ca, cb, std::conj(ca)};
    caxpy(c3, v, fw);
  } else QB_ERROR("blasQudaB200: unknown operation %s", name);
  for (int i = 0; i < 4; i++)
    if (hp[i]) {
      if (half) { copy(stage, *fp[i]); QB_CUDA(cudaMemcpyAsync(hp[i], stage.v, bytes, cudaMemcpyDeviceToHost, s)); QB_CUDA(cudaStreamSynchronize(s)); }
      else QB_CUDA(cudaMemcpyAsync(hp[i], fp[i]->v, bytes, cudaMemcpyDeviceToHost, s));
    }
  QB_CUDA(cudaStreamSynchronize(s));
  return nres;
}

// ---- bridges for the C++ facade (quda_cpp.cu) ------------------------------------------------------------------------------
namespace qb {
DiracTM *facade_make_dirac(const QudaInvertParam *p, bool pc, Prec gauge_prec) {
  require_gauge();
  DiracTM *d = make_dirac(p, pc, pick_gauge(gauge_prec));
  if (gauge_prec == PREC_HALF) d->gauge_vec = pick_gauge(PREC_SINGLE);   // int16 links, fp32 solver vectors
  return d;
}
void facade_load_spinor(SpinorField &f, const void *h, const QudaInvertParam *p) {
  import_spinor(f, h, to_prec(p->cpu_prec, "cpu_prec"), to_basis(p->gamma_basis), to_order(p->dirac_order), rt().compute);
  QB_CUDA(cudaStreamSynchronize(rt().compute));
}
void facade_save_spinor(void *h, const SpinorField &f, const QudaInvertParam *p) {
  export_spinor(h, f, to_prec(p->cpu_prec, "cpu_prec"), to_basis(p->gamma_basis), to_order(p->dirac_order), rt().compute);
  QB_CUDA(cudaStreamSynchronize(rt().compute));
}
void facade_fill_solver_param(SolverParam &s, const QudaInvertParam *p) { fill_solver_param(s, p); }
void facade_lattice(int *X4) {
  require_gauge();
  for (int d = 0; d < 4; d++) X4[d] = G.lat.geom.X[d];
}
}  // namespace qb
