/* C++ facade over libquda_b200.so: the `namespace quda` classes that the QKXTM code inside the reference library is written
 * against ("second boundary").  Names, argument order and meaning follow the reference headers:
 *
 *   ColorSpinorParam / ColorSpinorField / cpuColorSpinorField / cudaColorSpinorField   include/color_spinor_field.h:24-420
 *   DiracParam, Dirac (+ create, prepare, reconstruct), DiracM / DiracMdagM / DiracMdag  include/dirac_quda.h:18-145, 869-1030
 *   setDiracParam / createDirac / massRescale                                           lib/interface_quda.cpp:1265-1494
 *   SolverParam(QudaInvertParam&), Solver::create, Solver::operator()                   include/invert_quda.h:15-388
 *   blas::{zero, copy, ax, axpy, xpy, axpby, caxpy, norm2, xmyNorm, cDotProduct}         include/blas_quda.h
 *   TimeProfile                                                                         include/quda_internal.h
 *
 * so that the pattern of lib/interface_quda.cpp:6285-6500 (calcMG_* of the QKXTM fork)
 *
 *     createDirac(d, dSloppy, dPre, *param, pc_solve);
 *     ColorSpinorParam cpuParam(h_src, *param, X, pc_solution, QUDA_CPU_FIELD_LOCATION);
 *     ColorSpinorField *h_b = ColorSpinorField::Create(cpuParam);
 *     ColorSpinorParam cudaParam(cpuParam, *param);
 *     ColorSpinorField *b = new cudaColorSpinorField(*h_b, cudaParam), *x = new cudaColorSpinorField(cudaParam);
 *     DiracM m(*d), mSloppy(*dSloppy), mPre(*dPre);
 *     d->prepare(in, out, *x, *b, param->solution_type);
 *     SolverParam solverParam(*param);
 *     Solver *solve = Solver::create(solverParam, m, mSloppy, mPre, profile);
 *     (*solve)(*out, *in);
 *     d->reconstruct(*x, *b, param->solution_type);
 *     *h_x = *x;
 *
 * compiles and runs unchanged.  Only the operators of this build exist (Wilson, Wilson-clover, degenerate twisted mass, twisted
 * clover; GCR / MR / BiCGStab / MG).  The classes are thin handles: all data lives in the library's own device fields, there is no
 * CPU arithmetic behind cpuColorSpinorField (it only describes caller-owned host memory for the copy kernels).
 * Plain C++11, no CUDA or torch types. */
#ifndef QUDA_B200_CPP_H
#define QUDA_B200_CPP_H

#include <complex>
#include "quda.h"

namespace quda {

typedef std::complex<double> Complex;

class TimeProfile {
 public:
  explicit TimeProfile(const char *name_) : name(name_) {}
  void Print() const {}
  const char *name;
};

// ---- fields --------------------------------------------------------------------------------------------------------
class ColorSpinorField;

class ColorSpinorParam {
 public:
  QudaFieldLocation location;
  int nColor, nSpin, nDim;
  int x[QUDA_MAX_DIM];            // full-lattice extents (x[0] is halved for a parity field, as in the reference)
  QudaPrecision precision;
  QudaSiteSubset siteSubset;
  QudaSiteOrder siteOrder;
  QudaFieldOrder fieldOrder;
  QudaGammaBasis gammaBasis;
  QudaFieldCreate create;
  void *v;                        // host pointer for QUDA_REFERENCE_FIELD_CREATE
  QudaInvertParam inv_param;      // copy: host precision / order / basis used by the copy kernels
  ColorSpinorParam();
  // host field over caller memory (color_spinor_field.h:121-170)
  ColorSpinorParam(void *V, QudaInvertParam &inv_param, const int *X, const bool pc_solution,
                   QudaFieldLocation location = QUDA_CPU_FIELD_LOCATION);
  // device twin of a host field (color_spinor_field.h:173-205)
  ColorSpinorParam(const ColorSpinorParam &cpuParam, QudaInvertParam &inv_param);
};

class ColorSpinorField {
 public:
  virtual ~ColorSpinorField() {}
  static ColorSpinorField *Create(const ColorSpinorParam &param);
  virtual ColorSpinorField &operator=(const ColorSpinorField &src) = 0;   // copies across host / device, converting precision
  QudaFieldLocation Location() const { return location_; }
  QudaPrecision Precision() const { return precision_; }
  QudaSiteSubset SiteSubset() const { return subset_; }
  int Ncolor() const { return 3; }
  int Nspin() const { return 4; }
  int Ndim() const { return 4; }
  const int *X() const { return x_; }
  long Volume() const { return volume_; }
  virtual void *V() = 0;
  virtual const void *V() const = 0;
  struct Impl;
  Impl *impl() const { return impl_; }   // library internal

 protected:
  ColorSpinorField() : location_(QUDA_INVALID_FIELD_LOCATION), precision_(QUDA_INVALID_PRECISION), subset_(QUDA_INVALID_SITE_SUBSET), volume_(0), impl_(0) {}
  QudaFieldLocation location_;
  QudaPrecision precision_;
  QudaSiteSubset subset_;
  int x_[QUDA_MAX_DIM];
  long volume_;
  Impl *impl_;
};

// describes caller-owned host memory (QUDA_REFERENCE_FIELD_CREATE) or owns a zeroed host buffer
class cpuColorSpinorField : public ColorSpinorField {
 public:
  explicit cpuColorSpinorField(const ColorSpinorParam &param);
  virtual ~cpuColorSpinorField();
  virtual ColorSpinorField &operator=(const ColorSpinorField &src);
  virtual void *V();
  virtual const void *V() const;
};

class cudaColorSpinorField : public ColorSpinorField {
 public:
  explicit cudaColorSpinorField(const ColorSpinorParam &param);                       // zero / null create
  cudaColorSpinorField(const ColorSpinorField &src, const ColorSpinorParam &param);   // create and copy (host or device source)
  virtual ~cudaColorSpinorField();
  virtual ColorSpinorField &operator=(const ColorSpinorField &src);
  virtual void *V();
  virtual const void *V() const;
};

// ---- operators -----------------------------------------------------------------------------------------------------
class cudaGaugeField;   // opaque: the resident gauge fields are selected by precision
class cudaCloverField;

class DiracParam {
 public:
  QudaDiracType type;
  double kappa, mass, mu, epsilon;
  QudaMatPCType matpcType;
  QudaDagType dagger;
  cudaGaugeField *gauge;
  cudaCloverField *clover;
  QudaPrecision gauge_precision;   // which resident copy of the links this operator uses (set by setDirac*Param)
  QudaTwistFlavorType twist_flavor;
  int commDim[QUDA_MAX_DIM];
  QudaInvertParam inv_param;
  DiracParam();
};
void setDiracParam(DiracParam &diracParam, QudaInvertParam *inv_param, bool pc);
void setDiracSloppyParam(DiracParam &diracParam, QudaInvertParam *inv_param, bool pc);
void setDiracPreParam(DiracParam &diracParam, QudaInvertParam *inv_param, bool pc, bool comms);

class Dirac {
 public:
  virtual ~Dirac();
  static Dirac *create(const DiracParam &param);
  void Dslash(ColorSpinorField &out, const ColorSpinorField &in, const QudaParity parity) const;
  void DslashXpay(ColorSpinorField &out, const ColorSpinorField &in, const QudaParity parity, const ColorSpinorField &x, const double &k) const;
  void M(ColorSpinorField &out, const ColorSpinorField &in) const;
  void MdagM(ColorSpinorField &out, const ColorSpinorField &in) const;
  void Mdag(ColorSpinorField &out, const ColorSpinorField &in) const;
  // src / sol point into x, b (parity views) or at operator-owned temporaries, as in dirac_twisted_mass.cpp:418-520
  void prepare(ColorSpinorField *&src, ColorSpinorField *&sol, ColorSpinorField &x, ColorSpinorField &b, const QudaSolutionType) const;
  void reconstruct(ColorSpinorField &x, const ColorSpinorField &b, const QudaSolutionType) const;
  void Dagger(QudaDagType dag);
  void flipDagger();
  unsigned long long Flops() const;
  struct Impl;
  Impl *impl() const { return impl_; }   // library internal

 protected:
  Dirac() : impl_(0) {}
  Impl *impl_;
};
void createDirac(Dirac *&d, Dirac *&dSloppy, Dirac *&dPre, QudaInvertParam &param, const bool pc_solve);

class DiracMatrix {
 public:
  explicit DiracMatrix(const Dirac &d) : dirac(&d) {}
  explicit DiracMatrix(const Dirac *d) : dirac(d) {}
  virtual ~DiracMatrix() {}
  virtual void operator()(ColorSpinorField &out, const ColorSpinorField &in) const = 0;
  virtual bool isNormal() const = 0;
  const Dirac *Expose() const { return dirac; }
  unsigned long long flops() const { return dirac->Flops(); }

 protected:
  const Dirac *dirac;
};
class DiracM : public DiracMatrix {
 public:
  explicit DiracM(const Dirac &d) : DiracMatrix(d) {}
  explicit DiracM(const Dirac *d) : DiracMatrix(d) {}
  void operator()(ColorSpinorField &out, const ColorSpinorField &in) const { dirac->M(out, in); }
  bool isNormal() const { return false; }
};
class DiracMdagM : public DiracMatrix {
 public:
  explicit DiracMdagM(const Dirac &d) : DiracMatrix(d) {}
  explicit DiracMdagM(const Dirac *d) : DiracMatrix(d) {}
  void operator()(ColorSpinorField &out, const ColorSpinorField &in) const { dirac->MdagM(out, in); }
  bool isNormal() const { return true; }
};
class DiracMdag : public DiracMatrix {
 public:
  explicit DiracMdag(const Dirac &d) : DiracMatrix(d) {}
  explicit DiracMdag(const Dirac *d) : DiracMatrix(d) {}
  void operator()(ColorSpinorField &out, const ColorSpinorField &in) const { dirac->Mdag(out, in); }
  bool isNormal() const { return false; }
};

// out *= the source normalisation of the mass convention (interface_quda.cpp:1412-1494)
void massRescale(cudaColorSpinorField &b, QudaInvertParam &param);

// ---- solvers -------------------------------------------------------------------------------------------------------
struct SolverParam {
  QudaInverterType inv_type, inv_type_precondition;
  void *preconditioner;          // multigrid handle of newMultigridQuda when inv_type_precondition == QUDA_MG_INVERTER
  QudaUseInitGuess use_init_guess;
  double tol, delta, omega;
  int maxiter, Nkrylov;
  QudaPrecision precision, precision_sloppy, precision_precondition;
  double tol_precondition;
  int maxiter_precondition;
  // results
  double true_res, true_res_hq, secs, gflops;
  int iter;
  QudaInvertParam inv_param;
  explicit SolverParam(QudaInvertParam &param);
  void updateInvertParam(QudaInvertParam &param) const;   // writes true_res, iter, secs, gflops back (invert_quda.h:242-260)
};

class Solver {
 public:
  virtual ~Solver();
  static Solver *create(SolverParam &param, DiracMatrix &mat, DiracMatrix &matSloppy, DiracMatrix &matPrecon, TimeProfile &profile);
  void operator()(ColorSpinorField &out, ColorSpinorField &in);
  struct Impl;

 protected:
  Solver() : impl_(0) {}
  Impl *impl_;
};

// ---- BLAS on device fields -----------------------------------------------------------------------------------------
namespace blas {
void zero(ColorSpinorField &a);
void copy(ColorSpinorField &dst, const ColorSpinorField &src);
void ax(const double &a, ColorSpinorField &x);
void axpy(const double &a, ColorSpinorField &x, ColorSpinorField &y);
void xpy(ColorSpinorField &x, ColorSpinorField &y);
void axpby(const double &a, ColorSpinorField &x, const double &b, ColorSpinorField &y);
void caxpy(const Complex &a, ColorSpinorField &x, ColorSpinorField &y);
double norm2(const ColorSpinorField &a);
double xmyNorm(ColorSpinorField &x, ColorSpinorField &y);   // y = x - y, returns |y|^2
Complex cDotProduct(ColorSpinorField &x, ColorSpinorField &y);
}  // namespace blas

}  // namespace quda

#endif
