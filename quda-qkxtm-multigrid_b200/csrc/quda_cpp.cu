// Implementation of the `namespace quda` C++ facade declared in include/quda_cpp.h (reference: include/dirac_quda.h,
// include/invert_quda.h, include/color_spinor_field.h, include/blas_quda.h, lib/interface_quda.cpp:1265-1494).  Every class is a
// thin handle around the library's own device objects (qb::SpinorField, qb::DiracTM, qb::Solver); nothing is computed on the host.
#include <cstring>
#include <memory>
#include <vector>
#include "../../include/quda_cpp.h"
#include "blas.h"
#include "dirac.h"
#include "multigrid.h"
#include "solver.h"

namespace qb {
// bridges into interface.cu (process-global gauge / clover state lives there)
DiracTM *facade_make_dirac(const QudaInvertParam *p, bool pc, Prec gauge_prec);
void facade_load_spinor(SpinorField &f, const void *h, const QudaInvertParam *p);
void facade_save_spinor(void *h, const SpinorField &f, const QudaInvertParam *p);
void facade_fill_solver_param(SolverParam &s, const QudaInvertParam *p);
void facade_lattice(int *X4);
}  // namespace qb

namespace quda {

// ---- fields --------------------------------------------------------------------------------------------------------
struct ColorSpinorField::Impl {
  std::unique_ptr<qb::SpinorField> dev;   // device field (owner) ...
  qb::SpinorField view;                   // ... or a view into another field (prepare)
  bool is_view = false;
  void *host = nullptr;                   // host field
  bool own_host = false;
  size_t host_bytes = 0;
  QudaInvertParam inv;                    // host layout
  qb::SpinorField &field() { return is_view ? view : *dev; }
};
static qb::SpinorField &dev_of(const ColorSpinorField &f) {
  if (f.Location() != QUDA_CUDA_FIELD_LOCATION || !f.impl()) QB_ERROR("this operation needs a cudaColorSpinorField");
  return f.impl()->field();
}

ColorSpinorParam::ColorSpinorParam()
    : location(QUDA_INVALID_FIELD_LOCATION), nColor(3), nSpin(4), nDim(4), precision(QUDA_INVALID_PRECISION), siteSubset(QUDA_INVALID_SITE_SUBSET),
      siteOrder(QUDA_EVEN_ODD_SITE_ORDER), fieldOrder(QUDA_SPACE_SPIN_COLOR_FIELD_ORDER), gammaBasis(QUDA_DEGRAND_ROSSI_GAMMA_BASIS),
      create(QUDA_ZERO_FIELD_CREATE), v(nullptr) {
  for (int i = 0; i < QUDA_MAX_DIM; i++) x[i] = 1;
  memset(&inv_param, 0, sizeof(inv_param));
}

ColorSpinorParam::ColorSpinorParam(void *V, QudaInvertParam &ip, const int *X, const bool pc_solution, QudaFieldLocation loc) : ColorSpinorParam() {
  location = loc;
  for (int d = 0; d < 4; d++) x[d] = X[d];
  if (pc_solution) x[0] /= 2;
  precision = loc == QUDA_CPU_FIELD_LOCATION ? ip.cpu_prec : ip.cuda_prec;
  siteSubset = pc_solution ? QUDA_PARITY_SITE_SUBSET : QUDA_FULL_SITE_SUBSET;
  gammaBasis = ip.gamma_basis;
  create = QUDA_REFERENCE_FIELD_CREATE;
  v = V;
  inv_param = ip;
}

ColorSpinorParam::ColorSpinorParam(const ColorSpinorParam &cpu, QudaInvertParam &ip) : ColorSpinorParam() {
  *this = cpu;
  location = QUDA_CUDA_FIELD_LOCATION;
  precision = ip.cuda_prec;
  create = QUDA_NULL_FIELD_CREATE;
  v = nullptr;
  inv_param = ip;
}

ColorSpinorField *ColorSpinorField::Create(const ColorSpinorParam &param) {
  if (param.location == QUDA_CPU_FIELD_LOCATION) return new cpuColorSpinorField(param);
  if (param.location == QUDA_CUDA_FIELD_LOCATION) return new cudaColorSpinorField(param);
  QB_ERROR("ColorSpinorField::Create: invalid field location %d", (int)param.location);
}

static long param_volume(const ColorSpinorParam &p) { return (long)p.x[0] * p.x[1] * p.x[2] * p.x[3]; }

cpuColorSpinorField::cpuColorSpinorField(const ColorSpinorParam &p) {
  location_ = QUDA_CPU_FIELD_LOCATION; precision_ = p.precision; subset_ = p.siteSubset;
  for (int d = 0; d < QUDA_MAX_DIM; d++) x_[d] = p.x[d];
  volume_ = param_volume(p);
  impl_ = new Impl();
  impl_->inv = p.inv_param;
  impl_->host_bytes = (size_t)volume_ * 24 * (size_t)p.precision;
  if (p.create == QUDA_REFERENCE_FIELD_CREATE) {
    if (!p.v) QB_ERROR("cpuColorSpinorField: QUDA_REFERENCE_FIELD_CREATE without a pointer");
    impl_->host = p.v;
  } else {
    impl_->host = calloc(impl_->host_bytes, 1);
    impl_->own_host = true;
  }
}
cpuColorSpinorField::~cpuColorSpinorField() {
  if (impl_->own_host) free(impl_->host);
  delete impl_;
}
void *cpuColorSpinorField::V() { return impl_->host; }
const void *cpuColorSpinorField::V() const { return impl_->host; }
ColorSpinorField &cpuColorSpinorField::operator=(const ColorSpinorField &src) {
  if (&src == this) return *this;
  if (src.Location() == QUDA_CUDA_FIELD_LOCATION) {
    qb::facade_save_spinor(impl_->host, dev_of(src), &impl_->inv);
  } else {
    if (src.impl()->host_bytes != impl_->host_bytes) QB_ERROR("host field copy: size mismatch");
    memcpy(impl_->host, src.impl()->host, impl_->host_bytes);
  }
  return *this;
}

static void init_device_field(ColorSpinorField::Impl *&impl, const ColorSpinorParam &p) {
  int X[4];
  qb::facade_lattice(X);
  const bool parity = p.siteSubset == QUDA_PARITY_SITE_SUBSET;
  if (p.x[1] != X[1] || p.x[2] != X[2] || p.x[3] != X[3] || p.x[0] != (parity ? X[0] / 2 : X[0]))
    QB_ERROR("cudaColorSpinorField: extents %d %d %d %d do not match the loaded gauge field", p.x[0], p.x[1], p.x[2], p.x[3]);
  const long Vh = (long)X[0] * X[1] * X[2] * X[3] / 2;
  const qb::Prec prec = p.precision == QUDA_DOUBLE_PRECISION ? qb::PREC_DOUBLE : (p.precision == QUDA_SINGLE_PRECISION ? qb::PREC_SINGLE : qb::PREC_HALF);
  if (p.precision != QUDA_DOUBLE_PRECISION && p.precision != QUDA_SINGLE_PRECISION && p.precision != QUDA_HALF_PRECISION) QB_ERROR("cudaColorSpinorField: precision undefined");
  impl = new ColorSpinorField::Impl();
  impl->dev.reset(new qb::SpinorField(Vh, parity ? 1 : 2, prec));
  impl->inv = p.inv_param;
}

cudaColorSpinorField::cudaColorSpinorField(const ColorSpinorParam &p) {
  location_ = QUDA_CUDA_FIELD_LOCATION; precision_ = p.precision; subset_ = p.siteSubset;
  for (int d = 0; d < QUDA_MAX_DIM; d++) x_[d] = p.x[d];
  volume_ = param_volume(p);
  init_device_field(impl_, p);
  impl_->dev->zero(qb::rt().compute);   // NULL create leaves memory undefined in the reference; zero is a valid instance of that
}
cudaColorSpinorField::cudaColorSpinorField(const ColorSpinorField &src, const ColorSpinorParam &p) {
  location_ = QUDA_CUDA_FIELD_LOCATION; precision_ = p.precision; subset_ = p.siteSubset;
  for (int d = 0; d < QUDA_MAX_DIM; d++) x_[d] = p.x[d];
  volume_ = param_volume(p);
  init_device_field(impl_, p);
  *this = src;
}
cudaColorSpinorField::~cudaColorSpinorField() { delete impl_; }
void *cudaColorSpinorField::V() { return impl_->field().v; }
const void *cudaColorSpinorField::V() const { return impl_->field().v; }
ColorSpinorField &cudaColorSpinorField::operator=(const ColorSpinorField &src) {
  if (&src == this) return *this;
  if (src.Location() == QUDA_CPU_FIELD_LOCATION) {
    if (src.Volume() != volume_) QB_ERROR("host -> device copy: volume mismatch");
    qb::facade_load_spinor(impl_->field(), src.V(), &src.impl()->inv);
  } else {
    qb::blas::copy(impl_->field(), dev_of(src));
  }
  return *this;
}

// a device field object that views part of another one (prepare's src / sol)
namespace {
class ViewField : public cudaColorSpinorField {
 public:
  ViewField(const ColorSpinorParam &p) : cudaColorSpinorField(p) {}
};
ColorSpinorField *make_view(const qb::SpinorField &v, const ColorSpinorField &like) {
  // build through the public constructor on a minimal parameter set, then swap the storage for the view
  ColorSpinorParam p;
  p.location = QUDA_CUDA_FIELD_LOCATION;
  p.precision = like.Precision();
  p.siteSubset = v.nparity == 1 ? QUDA_PARITY_SITE_SUBSET : QUDA_FULL_SITE_SUBSET;
  int X[4];
  qb::facade_lattice(X);
  for (int d = 0; d < 4; d++) p.x[d] = X[d];
  if (v.nparity == 1) p.x[0] /= 2;
  p.inv_param = like.impl()->inv;
  ColorSpinorField *f = new ViewField(p);
  f->impl()->dev.reset();
  qb::SpinorField &w = f->impl()->view;   // field-by-field: SpinorField is not copyable (it may own memory); the view never does
  w.prec = v.prec; w.nparity = v.nparity; w.ncomplex = v.ncomplex; w.nspin = v.nspin; w.ncolor = v.ncolor; w.Vh = v.Vh;
  w.v = v.v; w.norm = v.norm; w.parity_bytes = v.parity_bytes; w.owner = false;
  f->impl()->is_view = true;
  return f;
}
}  // namespace

// ---- operators -----------------------------------------------------------------------------------------------------
struct Dirac::Impl {
  std::unique_ptr<qb::DiracTM> d;
  mutable std::vector<std::unique_ptr<ColorSpinorField>> views;   // src / sol handed out by prepare
};

DiracParam::DiracParam()
    : type(QUDA_INVALID_DIRAC), kappa(0.0), mass(0.0), mu(0.0), epsilon(0.0), matpcType(QUDA_MATPC_INVALID), dagger(QUDA_DAG_INVALID), gauge(nullptr),
      clover(nullptr), gauge_precision(QUDA_INVALID_PRECISION), twist_flavor(QUDA_TWIST_INVALID) {
  for (int i = 0; i < QUDA_MAX_DIM; i++) commDim[i] = 1;
  memset(&inv_param, 0, sizeof(inv_param));
}

// interface_quda.cpp:1265-1340
void setDiracParam(DiracParam &dp, QudaInvertParam *ip, bool pc) {
  switch (ip->dslash_type) {
    case QUDA_WILSON_DSLASH: dp.type = pc ? QUDA_WILSONPC_DIRAC : QUDA_WILSON_DIRAC; break;
    case QUDA_CLOVER_WILSON_DSLASH: dp.type = pc ? QUDA_CLOVERPC_DIRAC : QUDA_CLOVER_DIRAC; break;
    case QUDA_TWISTED_MASS_DSLASH: dp.type = pc ? QUDA_TWISTED_MASSPC_DIRAC : QUDA_TWISTED_MASS_DIRAC; break;
    case QUDA_TWISTED_CLOVER_DSLASH: dp.type = pc ? QUDA_TWISTED_CLOVERPC_DIRAC : QUDA_TWISTED_CLOVER_DIRAC; break;
    default: QB_ERROR("Unsupported dslash_type %d", (int)ip->dslash_type);
  }
  dp.matpcType = ip->matpc_type;
  dp.dagger = ip->dagger;
  dp.kappa = ip->kappa;
  dp.mass = ip->mass;
  dp.mu = ip->mu;
  dp.epsilon = ip->epsilon;
  dp.twist_flavor = ip->twist_flavor;
  dp.gauge_precision = ip->cuda_prec;
  dp.inv_param = *ip;
}
void setDiracSloppyParam(DiracParam &dp, QudaInvertParam *ip, bool pc) {
  setDiracParam(dp, ip, pc);
  dp.gauge_precision = ip->cuda_prec_sloppy;
}
void setDiracPreParam(DiracParam &dp, QudaInvertParam *ip, bool pc, bool comms) {
  setDiracParam(dp, ip, pc);
  dp.gauge_precision = ip->cuda_prec_precondition;
  for (int i = 0; i < 4; i++) dp.commDim[i] = comms ? 1 : 0;
}

Dirac::~Dirac() { delete impl_; }

Dirac *Dirac::create(const DiracParam &dp) {
  QudaInvertParam ip = dp.inv_param;
  ip.kappa = dp.kappa; ip.mu = dp.mu; ip.matpc_type = dp.matpcType; ip.dagger = dp.dagger; ip.twist_flavor = dp.twist_flavor;
  bool pc;
  switch (dp.type) {
    case QUDA_WILSON_DIRAC: ip.dslash_type = QUDA_WILSON_DSLASH; pc = false; break;
    case QUDA_WILSONPC_DIRAC: ip.dslash_type = QUDA_WILSON_DSLASH; pc = true; break;
    case QUDA_CLOVER_DIRAC: ip.dslash_type = QUDA_CLOVER_WILSON_DSLASH; pc = false; break;
    case QUDA_CLOVERPC_DIRAC: ip.dslash_type = QUDA_CLOVER_WILSON_DSLASH; pc = true; break;
    case QUDA_TWISTED_MASS_DIRAC: ip.dslash_type = QUDA_TWISTED_MASS_DSLASH; pc = false; break;
    case QUDA_TWISTED_MASSPC_DIRAC: ip.dslash_type = QUDA_TWISTED_MASS_DSLASH; pc = true; break;
    case QUDA_TWISTED_CLOVER_DIRAC: ip.dslash_type = QUDA_TWISTED_CLOVER_DSLASH; pc = false; break;
    case QUDA_TWISTED_CLOVERPC_DIRAC: ip.dslash_type = QUDA_TWISTED_CLOVER_DSLASH; pc = true; break;
    default: QB_ERROR("Unsupported Dirac type %d (this build: Wilson, clover, twisted mass, twisted clover)", (int)dp.type);
  }
  const QudaPrecision gp = dp.gauge_precision == QUDA_INVALID_PRECISION ? ip.cuda_prec : dp.gauge_precision;
  const qb::Prec prec = gp == QUDA_DOUBLE_PRECISION ? qb::PREC_DOUBLE : (gp == QUDA_SINGLE_PRECISION ? qb::PREC_SINGLE : qb::PREC_HALF);
  Dirac *d = new Dirac();
  d->impl_ = new Impl();
  d->impl_->d.reset(qb::facade_make_dirac(&ip, pc, prec));
  return d;
}

void Dirac::Dslash(ColorSpinorField &out, const ColorSpinorField &in, const QudaParity parity) const { impl_->d->Dslash(dev_of(out), dev_of(in), (int)parity); }
void Dirac::DslashXpay(ColorSpinorField &out, const ColorSpinorField &in, const QudaParity parity, const ColorSpinorField &x, const double &k) const {
  impl_->d->DslashXpay(dev_of(out), dev_of(in), (int)parity, dev_of(x), k);
}
void Dirac::M(ColorSpinorField &out, const ColorSpinorField &in) const { impl_->d->M(dev_of(out), dev_of(in)); }
void Dirac::MdagM(ColorSpinorField &out, const ColorSpinorField &in) const { impl_->d->MdagM(dev_of(out), dev_of(in)); }
void Dirac::Mdag(ColorSpinorField &out, const ColorSpinorField &in) const { impl_->d->Mdag(dev_of(out), dev_of(in)); }
void Dirac::Dagger(QudaDagType dag) { impl_->d->dagger = dag == QUDA_DAG_YES; }
void Dirac::flipDagger() { impl_->d->dagger = !impl_->d->dagger; }
unsigned long long Dirac::Flops() const { const unsigned long long f = (unsigned long long)impl_->d->flops; impl_->d->flops = 0; return f; }

void Dirac::prepare(ColorSpinorField *&src, ColorSpinorField *&sol, ColorSpinorField &x, ColorSpinorField &b, const QudaSolutionType st) const {
  qb::SpinorField s, o;
  impl_->d->prepare(s, o, dev_of(x), dev_of(b), (qb::SolutionType)(int)st);
  impl_->views.clear();
  impl_->views.emplace_back(make_view(s, b));
  impl_->views.emplace_back(make_view(o, x));
  src = impl_->views[0].get();
  sol = impl_->views[1].get();
}
void Dirac::reconstruct(ColorSpinorField &x, const ColorSpinorField &b, const QudaSolutionType st) const {
  impl_->d->reconstruct(dev_of(x), dev_of(b), (qb::SolutionType)(int)st);
}

// interface_quda.cpp:1386-1410
void createDirac(Dirac *&d, Dirac *&dSloppy, Dirac *&dPre, QudaInvertParam &param, const bool pc_solve) {
  DiracParam dp, dps, dpp;
  setDiracParam(dp, &param, pc_solve);
  setDiracSloppyParam(dps, &param, pc_solve);
  setDiracPreParam(dpp, &param, pc_solve, true);
  d = Dirac::create(dp);
  dSloppy = Dirac::create(dps);
  dPre = Dirac::create(dpp);
}

// interface_quda.cpp:1412-1494
void massRescale(cudaColorSpinorField &b, QudaInvertParam &param) {
  const double kappa = param.kappa;
  double s = 1.0;
  switch (param.solution_type) {
    case QUDA_MAT_SOLUTION:
      if (param.mass_normalization == QUDA_MASS_NORMALIZATION || param.mass_normalization == QUDA_ASYMMETRIC_MASS_NORMALIZATION) s = 2.0 * kappa;
      break;
    case QUDA_MATDAG_MAT_SOLUTION:
      if (param.mass_normalization == QUDA_MASS_NORMALIZATION || param.mass_normalization == QUDA_ASYMMETRIC_MASS_NORMALIZATION) s = 4.0 * kappa * kappa;
      break;
    case QUDA_MATPC_SOLUTION:
      if (param.mass_normalization == QUDA_MASS_NORMALIZATION) s = 4.0 * kappa * kappa;
      else if (param.mass_normalization == QUDA_ASYMMETRIC_MASS_NORMALIZATION) s = 2.0 * kappa;
      break;
    case QUDA_MATPCDAG_MATPC_SOLUTION:
      if (param.mass_normalization == QUDA_MASS_NORMALIZATION) s = 16.0 * kappa * kappa * kappa * kappa;
      else if (param.mass_normalization == QUDA_ASYMMETRIC_MASS_NORMALIZATION) s = 4.0 * kappa * kappa;
      break;
    default: QB_ERROR("Solution type %d not supported", (int)param.solution_type);
  }
  if (s != 1.0) qb::blas::ax(s, dev_of(b));
}

// ---- solvers -------------------------------------------------------------------------------------------------------
SolverParam::SolverParam(QudaInvertParam &p)
    : inv_type(p.inv_type), inv_type_precondition(p.inv_type_precondition), preconditioner(p.preconditioner), use_init_guess(p.use_init_guess), tol(p.tol),
      delta(p.reliable_delta), omega(p.omega), maxiter(p.maxiter), Nkrylov(p.gcrNkrylov), precision(p.cuda_prec), precision_sloppy(p.cuda_prec_sloppy),
      precision_precondition(p.cuda_prec_precondition), tol_precondition(p.tol_precondition), maxiter_precondition(p.maxiter_precondition), true_res(0.0),
      true_res_hq(0.0), secs(0.0), gflops(0.0), iter(0), inv_param(p) {}

void SolverParam::updateInvertParam(QudaInvertParam &p) const {
  p.true_res = true_res; p.true_res_hq = true_res_hq; p.iter += iter; p.secs += secs; p.gflops += gflops;
}

struct Solver::Impl {
  SolverParam *user;
  qb::SolverParam sp;
  std::unique_ptr<qb::Solver> solver;
  const Dirac *d, *dS, *dP;
  bool normal;
};

Solver::~Solver() { delete impl_; }

Solver *Solver::create(SolverParam &param, DiracMatrix &mat, DiracMatrix &matSloppy, DiracMatrix &matPrecon, TimeProfile &) {
  Solver *s = new Solver();
  s->impl_ = new Impl();
  Impl &I = *s->impl_;
  I.user = &param;
  QudaInvertParam ip = param.inv_param;
  ip.inv_type = param.inv_type; ip.inv_type_precondition = param.inv_type_precondition; ip.tol = param.tol; ip.maxiter = param.maxiter;
  ip.gcrNkrylov = param.Nkrylov; ip.reliable_delta = param.delta; ip.omega = param.omega; ip.use_init_guess = param.use_init_guess;
  ip.cuda_prec = param.precision; ip.cuda_prec_sloppy = param.precision_sloppy; ip.cuda_prec_precondition = param.precision_precondition;
  qb::facade_fill_solver_param(I.sp, &ip);
  I.d = mat.Expose(); I.dS = matSloppy.Expose(); I.dP = matPrecon.Expose();
  I.normal = mat.isNormal();
  qb::Solver *K = nullptr;
  if (param.inv_type_precondition == QUDA_MG_INVERTER) {
    if (!param.preconditioner) QB_ERROR("inv_type_precondition is QUDA_MG_INVERTER but `preconditioner` is not set (call newMultigridQuda first)");
    K = ((qb::MultigridSolver *)param.preconditioner)->mg.get();
  }
  qb::DiracMatrix m(I.d->impl()->d.get(), I.normal), mS(I.dS->impl()->d.get(), I.normal), mP(I.dP->impl()->d.get(), I.normal);
  // the sloppy / preconditioner operators act on fp32 vectors even when their links are int16 (as invertQuda sets them up)
  I.solver.reset(qb::Solver::create(I.sp, m, mS, mP, K));
  if (qb::GCR *g = dynamic_cast<qb::GCR *>(I.solver.get()))
    g->set_inner(param.maxiter_precondition <= 0 ? 10 : param.maxiter_precondition, param.tol_precondition <= 0 ? 0.1 : param.tol_precondition);
  return s;
}

void Solver::operator()(ColorSpinorField &out, ColorSpinorField &in) {
  Impl &I = *impl_;
  I.sp.iter = 0; I.sp.secs = 0;
  (*I.solver)(dev_of(out), dev_of(in));
  I.user->true_res = I.sp.true_res;
  I.user->iter = I.sp.iter;
  I.user->secs = I.sp.secs;
  I.user->gflops = 0.0;
}

// ---- BLAS ----------------------------------------------------------------------------------------------------------
namespace blas {
void zero(ColorSpinorField &a) { qb::blas::zero(dev_of(a)); }
void copy(ColorSpinorField &dst, const ColorSpinorField &src) { dst = src; }
void ax(const double &a, ColorSpinorField &x) { qb::blas::ax(a, dev_of(x)); }
void axpy(const double &a, ColorSpinorField &x, ColorSpinorField &y) { qb::blas::axpy(a, dev_of(x), dev_of(y)); }
void xpy(ColorSpinorField &x, ColorSpinorField &y) { qb::blas::xpy(dev_of(x), dev_of(y)); }
void axpby(const double &a, ColorSpinorField &x, const double &b, ColorSpinorField &y) { qb::blas::axpby(a, dev_of(x), b, dev_of(y)); }
void caxpy(const Complex &a, ColorSpinorField &x, ColorSpinorField &y) { qb::blas::caxpy(a, dev_of(x), dev_of(y)); }
double norm2(const ColorSpinorField &a) { return qb::blas::norm2(dev_of(a)); }
double xmyNorm(ColorSpinorField &x, ColorSpinorField &y) { return qb::blas::xmyNorm(dev_of(x), dev_of(y)); }
Complex cDotProduct(ColorSpinorField &x, ColorSpinorField &y) { return qb::blas::cDotProduct(dev_of(x), dev_of(y)); }
}  // namespace blas

}  // namespace quda
