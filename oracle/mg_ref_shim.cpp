// TEST INFRASTRUCTURE ONLY -- never linked into or called by the product.
//
// Driver around the REFERENCE's own multigrid host (CPU) code paths, compiled unmodified from where they lie under /root/reference
// by oracle/Makefile (target _ref/libmgref.so):
//   lib/transfer.cpp            Transfer::Transfer (geo map :220-258, spin map, fillV), Transfer::P / R (:270-348)
//   lib/transfer_util.cu        FillV :152, BlockOrthogonalize :441-469 (blockOrderV :168, blockGramSchmidt :327)
//   lib/prolongator.cu          Prolongate, CPU path :102-116
//   lib/restrictor.cu           Restrict, CPU path :90-125
//   lib/coarse_op.cu / .cuh     calculateY :1309-1497 on the host (ComputeUVCPU, ComputeTMAVCPU, ComputeVUVCPU, ...)
//   lib/coarsecoarse_op.cu      CoarseCoarseOp :148-184
//   lib/dslash_coarse.cu        ApplyCoarse, CPU coarseDslash :263-290
// plus the field classes they need (color_spinor_field.cpp, cpu_color_spinor_field.cpp, gauge_field.cpp, cpu_gauge_field.cpp,
// lattice_field.cpp, clover_field.cpp, malloc.cpp, util_quda.cpp, comm_single.cpp, comm_common.cpp, tune.cpp, timer.cpp).
// This file only (a) wraps caller arrays into the reference's cpuColorSpinorField / cpuGaugeField objects, (b) calls the functions
// above, (c) supplies the one routine whose implementation is NOT in the reference tree: BlasMagmaArgs::BatchInvertMatrix (MAGMA
// 1.7.0, un-vendored; call site lib/coarse_op.cuh:1466-1474) as a plain Gauss-Jordan inversion -- Xinv therefore stays "parity
// unpinned" and is only checked through X * Xinv = 1 -- and (d) defines a few symbols of the GPU side that the host paths never call
// but the linker wants.  Nothing here runs on a GPU.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <complex>
#include <vector>

#include <quda.h>
#include <quda_internal.h>
#include <color_spinor_field.h>
#include <gauge_field.h>
#include <clover_field.h>
#include <transfer.h>
#include <multigrid.h>
#include <blas_magma.h>
#include <comm_quda.h>

namespace quda {
// defined in lib/coarse_op.cu (no header declares it): the host-side worker CoarseOp() calls after downloading the links
void calculateY(GaugeField &Y, GaugeField &X, GaugeField &Xinv, GaugeField &Yhat, ColorSpinorField &uv, ColorSpinorField &av, const Transfer &T,
                const GaugeField &g, const CloverField &c, double kappa, double mu, QudaDiracType dirac, QudaMatPCType matpc);
}

using namespace quda;

// ---- (c) MAGMA stand-in ---------------------------------------------------------------------------------------------------------
void BlasMagmaArgs::OpenMagma() {}
void BlasMagmaArgs::CloseMagma() {}
BlasMagmaArgs::BlasMagmaArgs(const int prec) : m(0), max_nev(0), prec(prec), ldm(0), info(-1), init(false), alloc(false) {}
BlasMagmaArgs::~BlasMagmaArgs() {}
void BlasMagmaArgs::BatchInvertMatrix(void *Ainv_h, void *A_h, const int n, const int batch) {
  if (prec != 4) { fprintf(stderr, "mg_ref_shim: BatchInvertMatrix stand-in handles single precision only\n"); exit(1); }
  typedef std::complex<double> Z;
  std::vector<Z> a((size_t)n * 2 * n);
  const std::complex<float> *A = (const std::complex<float> *)A_h;
  std::complex<float> *Ai = (std::complex<float> *)Ainv_h;
  for (int b = 0; b < batch; b++) {
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        a[(size_t)i * 2 * n + j] = Z(A[(size_t)b * n * n + i * n + j]);
        a[(size_t)i * 2 * n + n + j] = i == j ? 1.0 : 0.0;
      }
    for (int c = 0; c < n; c++) {
      int piv = c;
      for (int r = c + 1; r < n; r++)
        if (std::abs(a[(size_t)r * 2 * n + c]) > std::abs(a[(size_t)piv * 2 * n + c])) piv = r;
      if (piv != c)
        for (int j = 0; j < 2 * n; j++) std::swap(a[(size_t)c * 2 * n + j], a[(size_t)piv * 2 * n + j]);
      const Z inv = 1.0 / a[(size_t)c * 2 * n + c];
      for (int j = 0; j < 2 * n; j++) a[(size_t)c * 2 * n + j] *= inv;
      for (int r = 0; r < n; r++) {
        if (r == c) continue;
        const Z f = a[(size_t)r * 2 * n + c];
        if (f == 0.0) continue;
        for (int j = 0; j < 2 * n; j++) a[(size_t)r * 2 * n + j] -= f * a[(size_t)c * 2 * n + j];
      }
    }
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) Ai[(size_t)b * n * n + i * n + j] = std::complex<float>(a[(size_t)i * 2 * n + n + j]);
  }
}

// ---- (d) symbols of the GPU / communication side (defined in lib/interface_quda.cpp, lib/blas_quda.cu, ... which are not part of
// this build); single process, no device ---------------------------------------------------------------------------------------
cudaDeviceProp deviceProp;   // read by Tunable::checkLaunchParam (include/tune_quda.h:230-260) even for host "kernels": nominal limits, no device
namespace {
struct DevicePropInit {
  DevicePropInit() {
    memset(&deviceProp, 0, sizeof(deviceProp));
    deviceProp.major = 2; deviceProp.minor = 0;
    deviceProp.maxThreadsPerBlock = 1024;
    deviceProp.maxThreadsDim[0] = 1024; deviceProp.maxThreadsDim[1] = 1024; deviceProp.maxThreadsDim[2] = 64;
    deviceProp.maxGridSize[0] = 2147483647; deviceProp.maxGridSize[1] = 65535; deviceProp.maxGridSize[2] = 65535;
    deviceProp.sharedMemPerBlock = 48 * 1024;
    deviceProp.warpSize = 32;
    deviceProp.multiProcessorCount = 1;
    // single-process communicator (what initCommsGridQuda / initQuda set up, lib/interface_quda.cpp:340-420), lib/comm_single.cpp
    const int dims[4] = {1, 1, 1, 1};
    comm_init(4, dims, rank_of, 0);
    // the reference logs every setup stage with printfQuda on stdout; bench.py owns stdout (one JSON line): keep the log only on request
    setOutputFile(getenv("MGREF_VERBOSE") ? stderr : fopen("/dev/null", "w"));
  }
  static int rank_of(const int *, void *) { return 0; }
} device_prop_init;
}
int commDim(int) { return 1; }
int commDimPartitioned(int) { return 0; }
bool commGlobalReduction() { return true; }
void commGlobalReductionSet(bool) {}
namespace quda {
namespace blas {
// only used for log lines on host fields here
double norm2(const ColorSpinorField &a) {
  if (a.Location() != QUDA_CPU_FIELD_LOCATION || a.Precision() != QUDA_SINGLE_PRECISION) return 0.0;
  const float *v = (const float *)a.V();
  double s = 0.0;
  for (size_t i = 0; i < a.Bytes() / sizeof(float); i++) s += (double)v[i] * v[i];
  return s;
}
double norm1(const ColorSpinorField &) { return 0.0; }
}  // namespace blas
}  // namespace quda

// (the run-time type information of the device field classes lives in mg_ref_rtti.c)

// ---- (a) wrappers -----------------------------------------------------------------------------------------------------------------
namespace {

TimeProfile g_profile("mg_ref_shim");

// host field [parity][x_cb][spin][colour][re,im] (QUDA_SPACE_SPIN_COLOR_FIELD_ORDER, even-odd site order, DeGrand-Rossi basis)
ColorSpinorParam cs_param(const int *X, int nspin, int ncolor, void *v, bool parity_subset = false) {
  ColorSpinorParam p;
  p.nDim = 4;
  for (int d = 0; d < 4; d++) p.x[d] = X[d];
  if (parity_subset) p.x[0] /= 2;
  p.pad = 0;
  p.precision = QUDA_SINGLE_PRECISION;
  p.siteSubset = parity_subset ? QUDA_PARITY_SITE_SUBSET : QUDA_FULL_SITE_SUBSET;
  p.location = QUDA_CPU_FIELD_LOCATION;
  p.nColor = ncolor;
  p.nSpin = nspin;
  p.twistFlavor = QUDA_TWIST_PLUS;
  p.siteOrder = QUDA_EVEN_ODD_SITE_ORDER;
  p.fieldOrder = QUDA_SPACE_SPIN_COLOR_FIELD_ORDER;
  p.gammaBasis = QUDA_DEGRAND_ROSSI_GAMMA_BASIS;
  p.create = v ? QUDA_REFERENCE_FIELD_CREATE : QUDA_ZERO_FIELD_CREATE;
  p.v = v;
  p.norm = 0;
  return p;
}

struct TransferH {
  int X[4], nspin, ncolor, nvec, geo_bs[4], spin_bs;
  std::vector<ColorSpinorField *> B;
  Transfer *T;
};

struct CoarseH {
  int Xc[4], N;
  cpuGaugeField *Y, *X, *Xinv, *Yhat;
};

CoarseH *new_coarse(const TransferH *t) {
  // DiracCoarse::initializeCoarse, lib/dirac_coarse.cpp:51-98 (the CPU fields)
  CoarseH *c = new CoarseH();
  int x[QUDA_MAX_DIM] = {0};
  for (int d = 0; d < 4; d++) { c->Xc[d] = t->X[d] / t->geo_bs[d]; x[d] = c->Xc[d]; }
  c->N = t->nvec * (t->nspin / t->spin_bs);
  GaugeFieldParam g;
  memcpy(g.x, x, sizeof(x));
  g.nColor = c->N;
  g.reconstruct = QUDA_RECONSTRUCT_NO;
  g.order = QUDA_QDP_GAUGE_ORDER;
  g.link_type = QUDA_COARSE_LINKS;
  g.t_boundary = QUDA_PERIODIC_T;
  g.create = QUDA_ZERO_FIELD_CREATE;
  g.precision = QUDA_SINGLE_PRECISION;
  g.nDim = 4;
  g.siteSubset = QUDA_FULL_SITE_SUBSET;
  g.ghostExchange = QUDA_GHOST_EXCHANGE_PAD;
  g.nFace = 1;
  g.geometry = QUDA_COARSE_GEOMETRY;
  c->Y = new cpuGaugeField(g);
  c->Yhat = new cpuGaugeField(g);
  g.ghostExchange = QUDA_GHOST_EXCHANGE_NO;
  g.nFace = 0;
  g.geometry = QUDA_SCALAR_GEOMETRY;
  c->X = new cpuGaugeField(g);
  c->Xinv = new cpuGaugeField(g);
  return c;
}

}  // namespace

extern "C" {

// enumerator values of the reference's include/enum_quda.h, by name (so that the python side never hard-codes them)
int mgref_enum(const char *name) {
#define E(x) if (!strcmp(name, #x)) return (int)x;
  E(QUDA_WILSON_DIRAC) E(QUDA_TWISTED_MASS_DIRAC) E(QUDA_TWISTED_MASSPC_DIRAC) E(QUDA_COARSE_DIRAC) E(QUDA_COARSEPC_DIRAC)
  E(QUDA_MATPC_EVEN_EVEN) E(QUDA_MATPC_ODD_ODD) E(QUDA_MATPC_INVALID)
#undef E
  fprintf(stderr, "mgref_enum: unknown enumerator %s\n", name);
  exit(1);
}

// B: nvec null vectors, each a full host field; geo_bs is adjusted in place as the reference does (transfer.cpp:31-44)
void *mgref_transfer_new(const float *B, int nvec, const int *X, int nspin, int ncolor, int *geo_bs, int spin_bs) {
  TransferH *t = new TransferH();
  for (int d = 0; d < 4; d++) t->X[d] = X[d];
  t->nspin = nspin; t->ncolor = ncolor; t->nvec = nvec; t->spin_bs = spin_bs;
  const size_t len = (size_t)X[0] * X[1] * X[2] * X[3] * nspin * ncolor * 2;
  for (int i = 0; i < nvec; i++) {
    ColorSpinorParam p = cs_param(X, nspin, ncolor, 0);
    ColorSpinorField *b = ColorSpinorField::Create(p);
    memcpy(b->V(), B + (size_t)i * len, len * sizeof(float));
    t->B.push_back(b);
  }
  int bs[QUDA_MAX_DIM] = {1, 1, 1, 1, 1, 1};
  for (int d = 0; d < 4; d++) bs[d] = geo_bs[d];
  t->T = new Transfer(t->B, nvec, bs, spin_bs, false, g_profile);
  for (int d = 0; d < 4; d++) { geo_bs[d] = bs[d]; t->geo_bs[d] = bs[d]; }
  return t;
}

void mgref_transfer_free(void *h) {
  TransferH *t = (TransferH *)h;
  delete t->T;
  for (auto b : t->B) delete b;
  delete t;
}

// the block-orthonormalised V in the reference's host order: [parity][x_cb][spin][fine colour][vector][re,im]
// (colour index of the packed V field = c * nvec + j, color_spinor_field_order.h:44-49 with nVec, transfer_util.cu:152-166)
void mgref_transfer_V(void *h, float *out) {
  TransferH *t = (TransferH *)h;
  const ColorSpinorField &V = t->T->Vectors();
  memcpy(out, V.V(), V.Bytes());
}

// parity < 0: full fine field; parity = 0 / 1: single-parity fine field (lib/multigrid.cpp:300-309 -> Transfer::setSiteSubset)
void mgref_P(void *h, float *fine_out, const float *coarse_in, int parity) {
  TransferH *t = (TransferH *)h;
  int Xc[4];
  for (int d = 0; d < 4; d++) Xc[d] = t->X[d] / t->geo_bs[d];
  ColorSpinorParam pc = cs_param(Xc, t->nspin / t->spin_bs, t->nvec, (void *)coarse_in);
  ColorSpinorParam pf = cs_param(t->X, t->nspin, t->ncolor, fine_out, parity >= 0);
  cpuColorSpinorField c(pc), f(pf);
  // MG::reset (lib/multigrid.cpp:300-309) always hands a valid parity, also for full fields
  t->T->setSiteSubset(parity >= 0 ? QUDA_PARITY_SITE_SUBSET : QUDA_FULL_SITE_SUBSET, parity == 1 ? QUDA_ODD_PARITY : QUDA_EVEN_PARITY);
  t->T->P(f, c);
}

void mgref_R(void *h, float *coarse_out, const float *fine_in, int parity) {
  TransferH *t = (TransferH *)h;
  int Xc[4];
  for (int d = 0; d < 4; d++) Xc[d] = t->X[d] / t->geo_bs[d];
  ColorSpinorParam pc = cs_param(Xc, t->nspin / t->spin_bs, t->nvec, coarse_out);
  ColorSpinorParam pf = cs_param(t->X, t->nspin, t->ncolor, (void *)fine_in, parity >= 0);
  cpuColorSpinorField c(pc), f(pf);
  // MG::reset (lib/multigrid.cpp:300-309) always hands a valid parity, also for full fields
  t->T->setSiteSubset(parity >= 0 ? QUDA_PARITY_SITE_SUBSET : QUDA_FULL_SITE_SUBSET, parity == 1 ? QUDA_ODD_PARITY : QUDA_EVEN_PARITY);
  t->T->R(c, f);
}

// Coarse links of the fine Wilson / twisted-mass operator: what CoarseOp (lib/coarse_op.cu:152-213) does after it has downloaded the
// links to a cpuGaugeField in QDP order.  gauge: 4 host arrays [parity][x_cb][row][col][re,im] fp32 (links already carry boundary
// conditions / anisotropy, as in the reference).  dirac: QudaDiracType (QUDA_WILSON_DIRAC, QUDA_TWISTED_MASS_DIRAC,
// QUDA_TWISTED_MASSPC_DIRAC, ...), mu = the `a` argument createCoarseOp passes (dirac_twisted_mass.cpp:224-228, :572-576).
void *mgref_coarse_op(void *h, float *const *gauge, double kappa, double mu, int dirac, int matpc) {
  TransferH *t = (TransferH *)h;
  CoarseH *c = new_coarse(t);
  GaugeFieldParam gp((void *)gauge);
  gp.nDim = 4;
  for (int d = 0; d < 4; d++) gp.x[d] = t->X[d];
  gp.precision = QUDA_SINGLE_PRECISION;
  gp.pad = 0;
  gp.siteSubset = QUDA_FULL_SITE_SUBSET;
  gp.nColor = 3;
  gp.nFace = 1;
  gp.reconstruct = QUDA_RECONSTRUCT_NO;
  gp.order = QUDA_QDP_GAUGE_ORDER;
  gp.link_type = QUDA_WILSON_LINKS;
  gp.t_boundary = QUDA_PERIODIC_T;   // the sign, if any, is already in the links
  gp.create = QUDA_REFERENCE_FIELD_CREATE;
  gp.geometry = QUDA_VECTOR_GEOMETRY;
  gp.ghostExchange = QUDA_GHOST_EXCHANGE_PAD;
  cpuGaugeField g(gp);

  ColorSpinorParam uvp(t->T->Vectors());
  uvp.create = QUDA_ZERO_FIELD_CREATE;
  uvp.location = QUDA_CPU_FIELD_LOCATION;
  ColorSpinorField *uv = ColorSpinorField::Create(uvp);
  ColorSpinorField *av = dirac == QUDA_TWISTED_MASSPC_DIRAC ? ColorSpinorField::Create(uvp) : &const_cast<ColorSpinorField &>(t->T->Vectors());

  CloverFieldParam cf;   // empty clover field, as CoarseOp builds when clover == NULL
  cf.nDim = 4;
  cf.pad = 0;
  cf.precision = QUDA_INVALID_PRECISION;
  for (int i = 0; i < cf.nDim; i++) cf.x[i] = 0;
  cf.order = QUDA_PACKED_CLOVER_ORDER;
  cf.direct = true;
  cf.inverse = true;
  cf.clover = NULL;
  cf.norm = 0;
  cf.cloverInv = NULL;
  cf.invNorm = 0;
  cf.create = QUDA_NULL_FIELD_CREATE;
  cf.siteSubset = QUDA_FULL_SITE_SUBSET;
  cpuCloverField cl(cf);

  calculateY(*c->Y, *c->X, *c->Xinv, *c->Yhat, *uv, *av, *t->T, g, cl, kappa, mu, (QudaDiracType)dirac, (QudaMatPCType)matpc);
  if (av != &t->T->Vectors()) delete av;
  delete uv;
  return c;
}

// Coarse links of a coarse operator (DiracCoarse::createCoarseOp / DiracCoarsePC::createCoarseOp, lib/dirac_coarse.cpp:209-212, :377-380):
// pc = 0 coarsens Y with QUDA_COARSE_DIRAC, pc = 1 coarsens Yhat with QUDA_COARSEPC_DIRAC
void *mgref_coarse_coarse_op(void *h_transfer, void *h_fine, double kappa, int pc, int matpc) {
  TransferH *t = (TransferH *)h_transfer;
  CoarseH *f = (CoarseH *)h_fine;
  CoarseH *c = new_coarse(t);
  if (pc) CoarseCoarseOp(*c->Y, *c->X, *c->Xinv, *c->Yhat, *t->T, *f->Yhat, *f->X, *f->Xinv, kappa, 0.0, QUDA_COARSEPC_DIRAC, (QudaMatPCType)matpc);
  else CoarseCoarseOp(*c->Y, *c->X, *c->Xinv, *c->Yhat, *t->T, *f->Y, *f->X, *f->Xinv, kappa, 0.0, QUDA_COARSE_DIRAC, QUDA_MATPC_INVALID);
  return c;
}

void mgref_coarse_free(void *h) {
  CoarseH *c = (CoarseH *)h;
  delete c->Y; delete c->X; delete c->Xinv; delete c->Yhat;
  delete c;
}

void mgref_coarse_dims(void *h, int *info5) {
  CoarseH *c = (CoarseH *)h;
  for (int d = 0; d < 4; d++) info5[d] = c->Xc[d];
  info5[4] = c->N;
}

// raw copy of a link field in the reference's QDP host order: which = 0 Y, 1 X, 2 Xinv, 3 Yhat; Y / Yhat: 8 directions
// (0..3 backward-type, 4..7 forward, dslash_coarse.cu:80,148), each [parity][x_cb][row][col][re,im], row = s * n_vec + c
void mgref_coarse_links(void *h, int which, float *out) {
  CoarseH *c = (CoarseH *)h;
  cpuGaugeField *f = which == 0 ? c->Y : (which == 1 ? c->X : (which == 2 ? c->Xinv : c->Yhat));
  const int geo = (which == 0 || which == 3) ? 8 : 1;
  const size_t per = (size_t)c->Xc[0] * c->Xc[1] * c->Xc[2] * c->Xc[3] * c->N * c->N * 2;
  for (int d = 0; d < geo; d++) memcpy(out + (size_t)d * per, ((float **)f->Gauge_p())[d], per * sizeof(float));
}

// ApplyCoarse on host fields (lib/dslash_coarse.cu:806-814 -> CPU coarseDslash :263-290):
//   out = [clover] X inB - [dslash] kappa * sum_mu (Y_{mu+4}(x) inA(x+mu) + Y_mu(x-mu)^dag inA(x-mu))
// parity < 0: full fields; use_yhat / use_xinv select the preconditioned links / the inverse of X
void mgref_apply_coarse(void *h, float *out, const float *inA, const float *inB, double kappa, int parity, int dslash, int clover, int use_yhat, int use_xinv) {
  CoarseH *c = (CoarseH *)h;
  const bool ps = parity >= 0;
  ColorSpinorParam po = cs_param(c->Xc, 2, c->N / 2, out, ps), pa = cs_param(c->Xc, 2, c->N / 2, (void *)inA, ps), pb = cs_param(c->Xc, 2, c->N / 2, (void *)inB, ps);
  cpuColorSpinorField o(po), a(pa), b(pb);
  ApplyCoarse(o, a, b, use_yhat ? *c->Yhat : *c->Y, use_xinv ? *c->Xinv : *c->X, kappa, ps ? parity : QUDA_INVALID_PARITY, dslash != 0, clover != 0);
}

}  // extern "C"
