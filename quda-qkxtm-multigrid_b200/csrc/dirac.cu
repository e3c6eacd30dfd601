// Wilson / twisted-mass operators on resident fields.
// Operator algebra follows /root/reference/lib/dirac_twisted_mass.cpp:129-174 (M), :246-294 (Dslash),
// :297-344 (DslashXpay), :346-403 (PC M), :418-520 (prepare), :522-570 (reconstruct); every
// composite below is issued as ONE fused hop kernel where the reference needs a hop plus separate
// twist / xpay kernels.
#include "dirac.h"
#include "blas.h"

namespace qb {

void Dirac::Mdag(SpinorField &out, const SpinorField &in) const {
  Dirac *self = const_cast<Dirac *>(this);
  self->dagger = !self->dagger;
  M(out, in);
  self->dagger = !self->dagger;
}

void Dirac::MdagM(SpinorField &out, const SpinorField &in) const {
  std::unique_ptr<SpinorField> t(new SpinorField(in.Vh, in.nparity, in.prec, in.nspin, in.ncolor, 1, in.nflavor));
  M(*t, in);
  Mdag(out, *t);
}

void Dirac::create_coarse_op(CoarseOperator &, const Transfer &, bool) const { QB_ERROR("create_coarse_op not implemented for this operator"); }
void Dirac::DiagInv(SpinorField &, const SpinorField &) const { QB_ERROR("DiagInv not implemented for this operator"); }
void Dirac::Diag(SpinorField &, const SpinorField &, int) const { QB_ERROR("Diag not implemented for this operator"); }

void DiracTM::Diag(SpinorField &out, const SpinorField &in, int parity) const {
  if (flavor == 2) QB_ERROR("Diag: the non-degenerate doublet is not supported");
  if (clover) { CloverTwist(out, in, parity, false); return; }
  if (flavor == 0) { if (out.v != in.v) blas::copy(out, in); return; }
  Twist(out, in);
}

// A^-1 on every site of the field (full or single parity); with a clover term (C + i a gamma5)^-1 parity by parity
void DiracTM::DiagInv(SpinorField &out, const SpinorField &in) const {
  if (flavor == 2) QB_ERROR("DiagInv: the non-degenerate doublet is not supported");
  if (clover) {
    if (in.nparity != 2) QB_ERROR("DiagInv with a clover term needs a full field");
    for (int p = 0; p < 2; p++) {
      SpinorField o, i;
      out.view_parity(o, p); in.view_parity(i, p);
      CloverTwist(o, i, p, true);
    }
    return;
  }
  if (flavor == 0) { if (out.v != in.v) blas::copy(out, in); return; }
  TwistInv(out, in);
}

DiracTM::DiracTM(Lattice *lat_, const GaugeField *gauge_, double kappa_, double mu_, int flavor_, bool pc_, int matpc_, bool dagger_)
    : lat(lat_), gauge(gauge_), gauge_vec(nullptr), kappa(kappa_), mu(mu_), flavor(flavor_), pc(pc_), matpc_type(matpc_) {
  dagger = dagger_;
  if (flavor != 0 && flavor != 1 && flavor != -1 && flavor != 2) QB_ERROR("twist flavor must be +-1 (degenerate) or 2 (non-degenerate doublet), got %d", flavor);
}

SpinorField &DiracTM::tmp(std::unique_ptr<SpinorField> &t, const SpinorField &like) const {
  if (!t || t->prec != like.prec || t->Vh != like.Vh || t->nparity != like.nparity || t->nbatch != like.nbatch || t->nflavor != like.nflavor)
    t.reset(new SpinorField(like.Vh, like.nparity, like.prec, 4, 3, like.nbatch, like.nflavor));
  return *t;
}

const GaugeField &DiracTM::links_for(const SpinorField &f) const {
  if (f.prec == gauge->prec) return *gauge;
  if (gauge_vec && f.prec == gauge_vec->prec) return *gauge_vec;
  QB_ERROR("operator works in precision %d but was handed a field of precision %d and no matching gauge field", (int)gauge->prec, (int)f.prec);
}

void DiracTM::WilsonDslash(SpinorField &out, const SpinorField &in, int parity) const {
  apply_hop(*lat, *gauge, out, in, parity, dagger, TwistCoef(), TwistCoef(), nullptr, TwistCoef());
  flops += 1320ll * in.Vh;
}

void DiracTM::WilsonDslashXpay(SpinorField &out, const SpinorField &in, int parity, const SpinorField &x, double k) const {
  apply_hop(*lat, *gauge, out, in, parity, dagger, TwistCoef(), TwistCoef(k, 0.0), &x, TwistCoef());
  flops += 1368ll * in.Vh;
}

void DiracTM::CloverTwist(SpinorField &out, const SpinorField &in, int parity, bool inverse, const SpinorField *x, double k) const {
  const double a0 = twist_a();
  const CloverField &cl = clover->get(in.prec, a0);
  if (!inverse) clover_apply(out, in, cl, parity, CLOVER_DIRECT, dagger ? -a0 : a0, x, k);
  else clover_apply(out, in, cl, parity, dagger ? CLOVER_INVERSE_ADJ : CLOVER_INVERSE, a0, x, k);
  flops += (inverse ? 576ll : 552ll) * in.Vh;
}

// out = [x +] k A^-1 D in (A^-dagger with dagger) in one launch: the inverse clover block is the hop kernel's epilogue.
// QB_FUSE_CLOVER=0: hop and clover_apply as two launches (measurement switch)
static bool fuse_clover() {
  static int f = -1;
  if (f < 0) { const char *e = getenv("QB_FUSE_CLOVER"); f = (e && atoi(e) == 0) ? 0 : 1; }
  return f == 1;
}
void DiracTM::WilsonDslashCloverInv(SpinorField &out, const SpinorField &in, int parity, const SpinorField *x, double k) const {
  if (!fuse_clover() || out.v == in.v || (x && x->v == in.v)) {
    SpinorField &t = tmp(tmp2, in);
    WilsonDslash(t, in, parity);
    CloverTwist(out, t, parity, true, x, k);
    return;
  }
  const CloverField &cl = clover->get(in.prec, twist_a());
  if (in.prec == PREC_HALF) {
    const_cast<CloverField &>(cl).ainv16();
    apply_hop(*lat, *gauge, out, in, parity, dagger, TwistCoef(), TwistCoef(k, 0.0), x, TwistCoef(), cl.Ainv16, dagger ? 2 : 1, cl.Ainv16_norm);
  } else apply_hop(*lat, *gauge, out, in, parity, dagger, TwistCoef(), TwistCoef(k, 0.0), x, TwistCoef(), cl.Ainv, dagger ? 2 : 1);
  flops += (1320ll + 576ll + (x ? 48ll : 0ll)) * in.Vh;
}

void DiracTM::Twist(SpinorField &out, const SpinorField &in) const { apply_twist_field(out, in, A()); }
void DiracTM::TwistInv(SpinorField &out, const SpinorField &in) const { apply_twist_field(out, in, Ainv()); }

// PC hop:  A^-1 D  (no dagger, or asymmetric)   |   D A^-1  (dagger & symmetric: twist on the input)
void DiracTM::Dslash(SpinorField &out, const SpinorField &in, int parity) const {
  if (flavor == 2) return NdegDslash(out, in, parity);
  if (clover) {
    // A^-1 D  |  D A^-1 (dagger & symmetric)   with A = C + i a gamma5   (dirac_twisted_clover.cpp:191-227, clover_reference.cpp:234-255)
    if (!dagger || !symmetric()) WilsonDslashCloverInv(out, in, parity, nullptr, 1.0);
    else { SpinorField &t = tmp(tmp2, in); CloverTwist(t, in, 1 - parity, true); WilsonDslash(out, t, parity); }
    return;
  }
  if (flavor == 0) return WilsonDslash(out, in, parity);
  if (!dagger || !symmetric()) apply_hop(*lat, *gauge, out, in, parity, dagger, TwistCoef(), Ainv(), nullptr, TwistCoef());
  else apply_hop(*lat, *gauge, out, in, parity, dagger, Ainv(), TwistCoef(), nullptr, TwistCoef());
  flops += 1392ll * in.Vh * in.nbatch;
}

void DiracTM::DslashRange(SpinorField &out, const SpinorField &in, int parity, int begin, int count, cudaStream_t s) const {
  if (flavor == 0) apply_hop_range(*lat, *gauge, out, in, parity, dagger, TwistCoef(), TwistCoef(), nullptr, TwistCoef(), begin, count, s);
  else if (!dagger || !symmetric()) apply_hop_range(*lat, *gauge, out, in, parity, dagger, TwistCoef(), Ainv(), nullptr, TwistCoef(), begin, count, s);
  else apply_hop_range(*lat, *gauge, out, in, parity, dagger, Ainv(), TwistCoef(), nullptr, TwistCoef(), begin, count, s);
  flops += 1392ll * count;
}

// out = x + k (A^-1 D | D A^-1) in     (dirac_twisted_mass.cpp:297-344: dagger alone selects the input twist)
void DiracTM::DslashXpay(SpinorField &out, const SpinorField &in, int parity, const SpinorField &x, double k) const {
  if (flavor == 2) {  // out = x + k (T^-1 D | D T^-1) in
    SpinorField &t = tmp(tmp2, in);
    if (!dagger) { WilsonDslash(t, in, parity); NdegTwist(out, t, true, k, &x, 1.0); }
    else { NdegTwist(t, in, true); WilsonDslashXpay(out, t, parity, x, k); }
    return;
  }
  if (clover) {
    if (!dagger) WilsonDslashCloverInv(out, in, parity, &x, k);
    else { SpinorField &t = tmp(tmp2, in); CloverTwist(t, in, 1 - parity, true); WilsonDslashXpay(out, t, parity, x, k); }
    return;
  }
  if (flavor == 0) return WilsonDslashXpay(out, in, parity, x, k);
  if (!dagger) apply_hop(*lat, *gauge, out, in, parity, dagger, TwistCoef(), Ainv(k), &x, TwistCoef());
  else apply_hop(*lat, *gauge, out, in, parity, dagger, Ainv(), TwistCoef(k, 0.0), &x, TwistCoef());
  flops += 1416ll * in.Vh * in.nbatch;
}

void DiracTM::M(SpinorField &out, const SpinorField &in) const {
  if (in.nbatch > 1 && (clover || flavor == 2 || in.prec != gauge->prec || !pc)) {
    // batch fields: only the plain even-odd twisted-mass / Wilson operator in the fields' own precision runs on all members at once
    // (batched hop, links read once per group); everything else member by member
    if (out.nbatch != in.nbatch) QB_ERROR("DiracTM::M: batch sizes differ");
    for (int c = 0; c < in.nbatch; c++) {
      SpinorField o, i;
      out.member(o, c); in.member(i, c);
      M(o, i);
    }
    return;
  }
  if (flavor == 2 && in.prec == gauge->prec) return NdegM(out, in);
  if (in.prec != gauge->prec) {
    // vectors live in another precision than the operator (e.g. fp32 smoother vectors, int16 operator):
    // convert, apply in the operator's precision, convert back
    if (!conv_in || conv_in->prec != gauge->prec || conv_in->Vh != in.Vh || conv_in->nparity != in.nparity || conv_in->nflavor != in.nflavor) {
      conv_in.reset(new SpinorField(in.Vh, in.nparity, gauge->prec, 4, 3, 1, in.nflavor));
      conv_out.reset(new SpinorField(in.Vh, in.nparity, gauge->prec, 4, 3, 1, in.nflavor));
    }
    copy_spinor(*conv_in, in, rt().compute);
    M(*conv_out, *conv_in);
    copy_spinor(out, *conv_out, rt().compute);
    return;
  }
  if (!pc) {
    // full operator on [even | odd]:  out_p = A in_p - kappa D_{p,1-p} in_{1-p}
    if (in.nparity != 2 || out.nparity != 2) QB_ERROR("full operator needs full fields");
    SpinorField oe, oo, ie, io;
    out.view_parity(oe, 0); out.view_parity(oo, 1);
    in.view_parity(ie, 0); in.view_parity(io, 1);
    if (clover) {
      // out_p = (C + i a g5) in_p - kappa D in_q   (dirac_twisted_clover.cpp M, clover_reference.cpp:257-282)
      SpinorField &t = tmp(tmp2, ie);
      CloverTwist(t, io, 1, false);
      apply_hop(*lat, *gauge, oo, ie, 1, dagger, TwistCoef(), TwistCoef(-kappa, 0.0), &t, TwistCoef());
      CloverTwist(t, ie, 0, false);
      apply_hop(*lat, *gauge, oe, io, 0, dagger, TwistCoef(), TwistCoef(-kappa, 0.0), &t, TwistCoef());
      flops += 1368ll * 2 * in.Vh;
      return;
    }
    apply_hop(*lat, *gauge, oo, ie, 1, dagger, TwistCoef(), TwistCoef(-kappa, 0.0), &io, flavor ? A() : TwistCoef());
    apply_hop(*lat, *gauge, oe, io, 0, dagger, TwistCoef(), TwistCoef(-kappa, 0.0), &ie, flavor ? A() : TwistCoef());
    flops += (1320ll + 72ll) * 2 * in.Vh;
    return;
  }
  if (in.nparity != 1 || out.nparity != 1) QB_ERROR("preconditioned operator needs single-parity fields");
  const double kappa2 = -kappa * kappa;
  const int p_out = (matpc_type == MATPC_EVEN_EVEN || matpc_type == MATPC_EVEN_EVEN_ASYM) ? 0 : 1;
  SpinorField &t = tmp(tmp1, in);
  if (clover) {
    if (symmetric()) {
      if (!dagger) {
        Dslash(t, in, 1 - p_out);                 // A^-1 D in
        DslashXpay(out, t, p_out, in, kappa2);    // in - kappa^2 A^-1 D t
      } else {
        // in - kappa^2 D A^-1 D A^-1 in
        Dslash(t, in, 1 - p_out);                 // D A^-1 in
        DslashXpay(out, t, p_out, in, kappa2);    // in + kappa2 D A^-1 t
      }
    } else {
      // (C + i a g5) in - kappa^2 D A^-1 D in, both daggers
      SpinorField &u = tmp(tmp3, in);
      WilsonDslashCloverInv(t, in, 1 - p_out, nullptr, 1.0);
      CloverTwist(u, in, p_out, false);
      WilsonDslashXpay(out, t, p_out, u, kappa2);
    }
    return;
  }
  Dslash(t, in, 1 - p_out);
  if (flavor == 0 || symmetric()) {
    DslashXpay(out, t, p_out, in, kappa2);
  } else {
    // asymmetric: out = A in - kappa^2 D t
    apply_hop(*lat, *gauge, out, t, p_out, dagger, TwistCoef(), TwistCoef(kappa2, 0.0), &in, A());
    flops += (1320ll + 96ll) * in.Vh;
  }
}

// Solve M x = b through the Schur complement on one parity (dirac_twisted_mass.cpp:418-520):
//   symmetric :  src = A^-1 (b_p + kappa D A^-1 b_q),  asymmetric:  src = b_p + kappa D A^-1 b_q
// with p the preconditioned parity and q the other one.  `src` aliases the q-half of x as in the reference.
void DiracTM::prepare(SpinorField &src, SpinorField &sol, SpinorField &x, SpinorField &b, SolutionType sol_type) const {
  if (flavor == 2 && pc && sol_type != SOL_MATPC && sol_type != SOL_MATPCDAG_MATPC) return NdegPrepare(src, sol, x, b);
  if (!pc) {
    if (sol_type == SOL_MATPC || sol_type == SOL_MATPCDAG_MATPC) QB_ERROR("Preconditioned solution requires a preconditioned solve_type");
    b.view_parity(src, 0); src.nparity = b.nparity; src.parity_bytes = b.parity_bytes;
    x.view_parity(sol, 0); sol.nparity = x.nparity; sol.parity_bytes = x.parity_bytes;
    return;
  }
  if (sol_type == SOL_MATPC || sol_type == SOL_MATPCDAG_MATPC) {
    b.view_parity(src, 0);
    x.view_parity(sol, 0);
    return;
  }
  if (x.nparity != 2 || b.nparity != 2) QB_ERROR("prepare: full-lattice source and solution required for a MAT solution");
  const int p = (matpc_type == MATPC_EVEN_EVEN || matpc_type == MATPC_EVEN_EVEN_ASYM) ? 0 : 1, q = 1 - p;
  SpinorField bp, bq;
  b.view_parity(bp, p); b.view_parity(bq, q);
  x.view_parity(src, q);
  x.view_parity(sol, p);
  if (clover) {
    // symmetric: src = A_p^-1 (b_p + kappa D A_q^-1 b_q);  asymmetric: src = b_p + kappa D A_q^-1 b_q   (dirac_twisted_clover.cpp prepare)
    if (links_for(bq).prec != gauge->prec) QB_ERROR("clover operators need the solver vectors in the operator's precision");
    SpinorField &t = tmp(tmp1, bq);
    CloverTwist(t, bq, q, true);
    if (symmetric()) {
      SpinorField &u = tmp(tmp3, bq);
      WilsonDslashXpay(u, t, p, bp, kappa);
      CloverTwist(src, u, p, true);
    } else {
      WilsonDslashXpay(src, t, p, bp, kappa);
    }
    return;
  }
  if (flavor == 0) {
    apply_hop(*lat, links_for(bq), src, bq, p, dagger, TwistCoef(), TwistCoef(kappa, 0.0), &bp, TwistCoef());
  } else if (symmetric()) {
    const TwistCoef ai = Ainv();
    apply_hop(*lat, links_for(bq), src, bq, p, dagger, ai, TwistCoef(kappa * ai.p, kappa * ai.q), &bp, ai);
  } else {
    apply_hop(*lat, links_for(bq), src, bq, p, dagger, Ainv(), TwistCoef(kappa, 0.0), &bp, TwistCoef());
  }
  flops += 1440ll * b.Vh;
}

// x_q = A^-1 (b_q + kappa D x_p)   (dirac_twisted_mass.cpp:522-570)
void DiracTM::reconstruct(SpinorField &x, const SpinorField &b, SolutionType sol_type) const {
  if (!pc) return;
  if (sol_type == SOL_MATPC || sol_type == SOL_MATPCDAG_MATPC) return;
  if (flavor == 2) return NdegReconstruct(x, b);
  const int p = (matpc_type == MATPC_EVEN_EVEN || matpc_type == MATPC_EVEN_EVEN_ASYM) ? 0 : 1, q = 1 - p;
  SpinorField xp, xq, bq;
  x.view_parity(xp, p); x.view_parity(xq, q);
  b.view_parity(bq, q);
  if (clover) {
    // x_q = A_q^-1 (b_q + kappa D x_p)
    SpinorField &t = tmp(tmp1, xp);
    WilsonDslashXpay(t, xp, q, bq, kappa);
    CloverTwist(xq, t, q, true);
    return;
  }
  if (flavor == 0) {
    apply_hop(*lat, links_for(xp), xq, xp, q, dagger, TwistCoef(), TwistCoef(kappa, 0.0), &bq, TwistCoef());
  } else {
    const TwistCoef ai = Ainv();
    apply_hop(*lat, links_for(xp), xq, xp, q, dagger, TwistCoef(), TwistCoef(kappa * ai.p, kappa * ai.q), &bq, ai);
  }
  flops += 1416ll * b.Vh;
}

// ---- non-degenerate twisted-mass doublet ---------------------------------------------------------------------------------------
// T = 1 + i a gamma5 tau3 + b tau1 with a = 2 kappa mu, b = -2 kappa epsilon; T^-1 = (1 - i a gamma5 tau3 - b tau1) / (1 + a^2 - b^2);
// the dagger flips a (wilson_dslash_reference.cpp:412-445).  The two flavours of a field go through ONE batched Wilson hop (links read
// once for both), the flavour mixing is the site-local apply_ndeg_twist.
void DiracTM::NdegTwist(SpinorField &out, const SpinorField &in, bool inverse, double c1, const SpinorField *x, double c2) const {
  double a = 2.0 * kappa * mu, b = -2.0 * kappa * epsilon, d = 1.0;
  if (inverse) { a = -a; b = -b; d = 1.0 / (1.0 + a * a - b * b); }
  if (dagger) a = -a;
  apply_ndeg_twist(out, in, a, b, d, c1, x, c2);
  flops += 96ll * in.Vh * in.nparity;
}

// T^-1 D  (no dagger, or asymmetric)   |   D T^-1  (dagger & symmetric)        [tm_ndeg_dslash :461-473]
void DiracTM::NdegDslash(SpinorField &out, const SpinorField &in, int parity) const {
  if (in.nflavor != 2 || out.nflavor != 2) QB_ERROR("the non-degenerate doublet operator needs flavour-doublet fields");
  if (!dagger || !symmetric()) {
    WilsonDslash(out, in, parity);
    NdegTwist(out, out, true);
  } else {
    SpinorField &t = tmp(tmp2, in);
    NdegTwist(t, in, true);
    WilsonDslash(out, t, parity);
  }
}

void DiracTM::NdegM(SpinorField &out, const SpinorField &in) const {
  if (in.nflavor != 2 || out.nflavor != 2) QB_ERROR("the non-degenerate doublet operator needs flavour-doublet fields");
  if (!pc) {
    // out_p = T in_p - kappa D in_q        [tm_ndeg_mat :544-587]
    if (in.nparity != 2 || out.nparity != 2) QB_ERROR("full operator needs full fields");
    SpinorField oe, oo, ie, io;
    out.view_parity(oe, 0); out.view_parity(oo, 1);
    in.view_parity(ie, 0); in.view_parity(io, 1);
    WilsonDslash(oo, ie, 1);
    WilsonDslash(oe, io, 0);
    NdegTwist(oo, io, false, 1.0, &oo, -kappa);
    NdegTwist(oe, ie, false, 1.0, &oe, -kappa);
    return;
  }
  if (in.nparity != 1 || out.nparity != 1) QB_ERROR("preconditioned operator needs single-parity fields");
  const double kappa2 = -kappa * kappa;
  const int p = (matpc_type == MATPC_EVEN_EVEN || matpc_type == MATPC_EVEN_EVEN_ASYM) ? 0 : 1, q = 1 - p;
  SpinorField &t = tmp(tmp1, in);
  if (!symmetric()) {
    // T in - kappa^2 D T^-1 D in        [tm_ndeg_matpc, asymmetric branch, both daggers]
    SpinorField &u = tmp(tmp3, in);
    WilsonDslash(t, in, q);
    NdegTwist(t, t, true);
    WilsonDslash(u, t, p);
    NdegTwist(out, in, false, 1.0, &u, kappa2);
  } else if (!dagger) {
    // in - kappa^2 T^-1 D T^-1 D in
    SpinorField &u = tmp(tmp3, in);
    WilsonDslash(t, in, q);
    NdegTwist(t, t, true);
    WilsonDslash(u, t, p);
    NdegTwist(out, u, true, kappa2, &in, 1.0);
  } else {
    // in - kappa^2 D T^-1 D T^-1 in   (T daggered)
    SpinorField &u = tmp(tmp3, in);
    NdegTwist(t, in, true);
    WilsonDslash(u, t, q);
    NdegTwist(t, u, true);
    WilsonDslashXpay(out, t, p, in, kappa2);
  }
}

// symmetric: src = T^-1 (b_p + kappa D T^-1 b_q);  asymmetric: src = b_p + kappa D T^-1 b_q     (as the degenerate prepare above)
void DiracTM::NdegPrepare(SpinorField &src, SpinorField &sol, SpinorField &x, SpinorField &b) const {
  if (x.nparity != 2 || b.nparity != 2 || x.nflavor != 2 || b.nflavor != 2) QB_ERROR("prepare: full-lattice doublet source and solution required");
  if (b.prec != gauge->prec) QB_ERROR("the doublet operator needs the solver vectors in the operator's precision");
  const int p = (matpc_type == MATPC_EVEN_EVEN || matpc_type == MATPC_EVEN_EVEN_ASYM) ? 0 : 1, q = 1 - p;
  SpinorField bp, bq;
  b.view_parity(bp, p); b.view_parity(bq, q);
  x.view_parity(src, q);
  x.view_parity(sol, p);
  SpinorField &t = tmp(tmp1, bq);
  NdegTwist(t, bq, true);
  if (symmetric()) {
    SpinorField &u = tmp(tmp3, bq);
    WilsonDslashXpay(u, t, p, bp, kappa);
    NdegTwist(src, u, true);
  } else {
    WilsonDslashXpay(src, t, p, bp, kappa);
  }
}

// x_q = T^-1 (b_q + kappa D x_p)
void DiracTM::NdegReconstruct(SpinorField &x, const SpinorField &b) const {
  const int p = (matpc_type == MATPC_EVEN_EVEN || matpc_type == MATPC_EVEN_EVEN_ASYM) ? 0 : 1, q = 1 - p;
  SpinorField xp, xq, bq;
  x.view_parity(xp, p); x.view_parity(xq, q);
  b.view_parity(bq, q);
  SpinorField &t = tmp(tmp1, xp);
  WilsonDslashXpay(t, xp, q, bq, kappa);
  NdegTwist(xq, t, true);
}

}  // namespace qb
