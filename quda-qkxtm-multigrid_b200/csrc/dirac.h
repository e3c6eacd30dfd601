// Dirac operators on resident fields (host-side objects choosing kernel coefficients).
// Mirrors the reference's operator interface for this path (/root/reference/include/dirac_quda.h:88-165):
// Dslash, DslashXpay, M, Mdag, MdagM, prepare, reconstruct -- same names, argument meaning and
// parity conventions -- for Wilson, degenerate twisted mass (full and even-odd preconditioned)
// and, in coarse.h, the multigrid coarse operator.
#pragma once
#include <memory>
#include "clover.h"
#include "dslash_api.h"

namespace qb {

enum MatPC { MATPC_EVEN_EVEN = 0, MATPC_ODD_ODD = 1, MATPC_EVEN_EVEN_ASYM = 2, MATPC_ODD_ODD_ASYM = 3 };
enum DiracType { DIRAC_WILSON, DIRAC_WILSON_PC, DIRAC_TM, DIRAC_TM_PC, DIRAC_COARSE, DIRAC_COARSE_PC };
enum SolutionType { SOL_MAT = 0, SOL_MATDAG_MAT = 1, SOL_MATPC = 2, SOL_MATPC_DAG = 3, SOL_MATPCDAG_MATPC = 4 };

class Transfer;
struct CoarseOperator;

// Abstract operator: what the solvers and the multigrid see.
class Dirac {
 public:
  bool dagger = false;
  mutable long long flops = 0;
  virtual ~Dirac() {}
  virtual DiracType type() const = 0;
  virtual bool is_pc() const = 0;
  virtual int matpc() const { return MATPC_EVEN_EVEN; }
  virtual Prec precision() const = 0;
  // field factory matching what M() acts on (parity field for PC operators, full field otherwise)
  virtual SpinorField *new_field(Prec prec) const = 0;
  virtual SpinorField *new_parity_field(Prec prec) const = 0;

  virtual void Dslash(SpinorField &out, const SpinorField &in, int parity) const = 0;
  virtual void DslashXpay(SpinorField &out, const SpinorField &in, int parity, const SpinorField &x, double k) const = 0;
  virtual void M(SpinorField &out, const SpinorField &in) const = 0;
  virtual void MdagM(SpinorField &out, const SpinorField &in) const;
  void Mdag(SpinorField &out, const SpinorField &in) const;
  // source preparation / solution reconstruction for the even-odd preconditioned system
  // x, b: full fields.  On return src/sol are (views of) the fields the solver works on.
  virtual void prepare(SpinorField &src, SpinorField &sol, SpinorField &x, SpinorField &b, SolutionType sol_type) const = 0;
  virtual void reconstruct(SpinorField &x, const SpinorField &b, SolutionType sol_type) const = 0;
  // build the coarse operator of this operator through the transfer T (multigrid setup).  preconditioned = true: the Galerkin product
  // of S^-1 M (S = site-diagonal term: A on the fine grid, X on a coarse grid) whose even-odd Schur complement is the symmetric
  // preconditioned operator -- what Dirac*PC::createCoarseOp builds in the reference (lib/dirac_twisted_mass.cpp:580,
  // lib/dirac_coarse.cpp:377, lib/coarse_op.cuh:1349-1440)
  virtual void create_coarse_op(CoarseOperator &coarse, const Transfer &T, bool preconditioned = false) const;
  // out = S^-1 in on every site of a full or parity field (S as above)
  virtual void DiagInv(SpinorField &out, const SpinorField &in) const;
  // out = S in on the single-parity field `in` of parity `parity` (S = A on the fine grid, X on a coarse grid)
  virtual void Diag(SpinorField &out, const SpinorField &in, int parity) const;
};

// Wilson and degenerate twisted-mass operator (lib/dirac_wilson.cpp, lib/dirac_twisted_mass.cpp)
class DiracTM : public Dirac {
 public:
  Lattice *lat;
  const GaugeField *gauge;      // links in the operator's working precision
  const GaugeField *gauge_vec;  // links in the precision of the solver vectors (prepare / reconstruct when they differ), may be null
  double kappa, mu;
  double epsilon = 0.0;   // flavour splitting of the non-degenerate doublet
  int flavor;   // +-1 twisted mass, 0 = plain Wilson, 2 = non-degenerate doublet (fields with nflavor = 2)
  CloverSet *clover = nullptr;  // set: Wilson-clover / twisted-clover, the site-local term is C + i a gamma5 (clover.h)
  bool pc;
  int matpc_type;
  mutable std::unique_ptr<SpinorField> tmp1, tmp2, tmp3, conv_in, conv_out;

  DiracTM(Lattice *lat, const GaugeField *gauge, double kappa, double mu, int flavor, bool pc, int matpc_type, bool dagger);
  DiracType type() const override { return flavor == 0 ? (pc ? DIRAC_WILSON_PC : DIRAC_WILSON) : (pc ? DIRAC_TM_PC : DIRAC_TM); }
  bool is_pc() const override { return pc; }
  int matpc() const override { return matpc_type; }
  Prec precision() const override { return gauge->prec; }
  int nflavor() const { return flavor == 2 ? 2 : 1; }
  SpinorField *new_field(Prec prec) const override { return new SpinorField(lat->geom.Vh, pc ? 1 : 2, prec, 4, 3, 1, nflavor()); }
  SpinorField *new_parity_field(Prec prec) const override { return new SpinorField(lat->geom.Vh, 1, prec, 4, 3, 1, nflavor()); }

  bool symmetric() const { return matpc_type == MATPC_EVEN_EVEN || matpc_type == MATPC_ODD_ODD; }
  double twist_a() const { return 2.0 * kappa * mu * flavor; }   // A = 1 + i a gamma5
  TwistCoef A() const { return TwistCoef(1.0, dagger ? -twist_a() : twist_a()); }
  TwistCoef Ainv(double scale = 1.0) const {
    const double a = dagger ? twist_a() : -twist_a();  // inverse twist: a -> -a, dagger flips again
    const double b = scale / (1.0 + a * a);
    return TwistCoef(b, b * a);
  }

  // plain Wilson hop (no twist): out = D in,  out = x + k D in
  void WilsonDslash(SpinorField &out, const SpinorField &in, int parity) const;
  void WilsonDslashXpay(SpinorField &out, const SpinorField &in, int parity, const SpinorField &x, double k) const;
  void Twist(SpinorField &out, const SpinorField &in) const;      // A in
  void TwistInv(SpinorField &out, const SpinorField &in) const;   // A^-1 in
  // clover operators: out(parity) = [x +] k (C + i a g5) in  /  [x +] k (C + i a g5)^-1 in, daggered as the operator is
  void CloverTwist(SpinorField &out, const SpinorField &in, int parity, bool inverse, const SpinorField *x = nullptr, double k = 1.0) const;
  // out = [x +] k A^-1 D in   (A^-dagger D^dagger when daggered), one launch
  void WilsonDslashCloverInv(SpinorField &out, const SpinorField &in, int parity, const SpinorField *x, double k) const;

  void Dslash(SpinorField &out, const SpinorField &in, int parity) const override;
  // Dslash restricted to the checkerboard range [begin, begin+count) of the output, on stream s (unpartitioned lattice)
  void DslashRange(SpinorField &out, const SpinorField &in, int parity, int begin, int count, cudaStream_t s) const;
  void DslashXpay(SpinorField &out, const SpinorField &in, int parity, const SpinorField &x, double k) const override;
  void M(SpinorField &out, const SpinorField &in) const override;
  void prepare(SpinorField &src, SpinorField &sol, SpinorField &x, SpinorField &b, SolutionType sol_type) const override;
  void reconstruct(SpinorField &x, const SpinorField &b, SolutionType sol_type) const override;
  void create_coarse_op(CoarseOperator &coarse, const Transfer &T, bool preconditioned = false) const override;
  void DiagInv(SpinorField &out, const SpinorField &in) const override;
  void Diag(SpinorField &out, const SpinorField &in, int parity) const override;

 private:
  // non-degenerate doublet (wilson_dslash_reference.cpp:412-587; dirac_twisted_mass.cpp handles it through the same class)
  void NdegTwist(SpinorField &out, const SpinorField &in, bool inverse, double c1 = 1.0, const SpinorField *x = nullptr, double c2 = 0.0) const;
  void NdegDslash(SpinorField &out, const SpinorField &in, int parity) const;
  void NdegM(SpinorField &out, const SpinorField &in) const;
  void NdegPrepare(SpinorField &src, SpinorField &sol, SpinorField &x, SpinorField &b) const;
  void NdegReconstruct(SpinorField &x, const SpinorField &b) const;
  SpinorField &tmp(std::unique_ptr<SpinorField> &t, const SpinorField &like) const;
  const GaugeField &links_for(const SpinorField &f) const;
};

// functors handed to solvers (cf. DiracM / DiracMdagM, include/dirac_quda.h:869-1030)
struct DiracMatrix {
  const Dirac *d;
  bool normal;  // false: M, true: MdagM
  DiracMatrix(const Dirac *d_, bool normal_ = false) : d(d_), normal(normal_) {}
  void operator()(SpinorField &out, const SpinorField &in) const {
    if (normal) d->MdagM(out, in);
    else d->M(out, in);
  }
};

}  // namespace qb
