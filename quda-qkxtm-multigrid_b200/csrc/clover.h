// Clover term for the Wilson-clover / twisted-clover operators (reference:
// lib/clover_field.cpp, lib/clover_invert.cu, lib/dslash_core/tmc_core.h, lib/dirac_twisted_clover.cpp).
//
// The site-local part of the operator is  A = C + i a gamma5  with C the clover matrix (1 + kappa-normalised
// sigma.F term as the application supplies it), block diagonal in chirality: two Hermitian 6 x 6 blocks per site.
// In the internal DeGrand-Rossi basis gamma5 = diag(+1, +1, -1, -1), so block chi is  C_chi + i s_chi a  (s = +1, -1).
// Resident data, structure-of-arrays of 16-byte planes like every other field:
//   C    [parity][plane][cb]  the packed Hermitian blocks exactly as the host hands them (72 reals per site)
//   Ainv [parity][plane][cb]  (C + i s a)^-1 as two full complex 6 x 6 blocks, row major (144 reals per site), computed
//                             on the device in fp64 (the reference stores (C^2 + a^2)^-1 and applies C, the twist and that
//                             inverse in turn: three matrix-vector products where this needs one)
// The daggered operator uses the conjugate transpose of the same blocks on the fly.
#pragma once
#include <memory>
#include "field.h"

namespace qb {

struct CloverField {
  Prec prec = PREC_DOUBLE;  // PREC_DOUBLE or PREC_SINGLE (int16 spinors are multiplied with the fp32 copy)
  long Vh = 0;
  void *C = nullptr;
  void *Ainv = nullptr;
  // fp32 copies only: the inverse blocks once more as int16 fixed point + one float norm per site ([parity][36 planes of 4 x int16][cb],
  // norm [parity][cb]; the reference keeps its half-precision clover as short4 + norm too, lib/clover_field.cpp) for the int16 hop's
  // fused epilogue: 292 instead of 576 B per site.  Built by ainv16() on first use.
  void *Ainv16 = nullptr;
  float *Ainv16_norm = nullptr;
  double a16 = 0.0;
  bool have16 = false;
  void ainv16();
  double a = 0.0;           // twist the inverse was built for
  CloverField(long Vh, Prec prec);
  ~CloverField();
  CloverField(const CloverField &) = delete;
  int reals_per_plane() const { return prec == PREC_DOUBLE ? 2 : 4; }
};

// The loaded clover term: fp64 master copy in host order on the device + working copies per precision.
struct CloverSet {
  long Vh = 0;
  double *master = nullptr;   // [V][72] packed, even sites first (QUDA_PACKED_CLOVER_ORDER)
  // per precision up to CLOVER_CACHE working copies, each with the inverse built for one twist `a`: the outer operator and the multigrid
  // operators (kappa, mu scaled by delta_kappaPR / delta_muPR) may use different twists in turn and must not evict each other
  static constexpr int CLOVER_CACHE = 3;
  struct Slot { std::unique_ptr<CloverField> f; unsigned long last_use = 0; };
  Slot d64[CLOVER_CACHE], f32[CLOVER_CACHE];
  unsigned long use_clock = 0;
  bool loaded = false;
  // working copy in the precision the operator computes in (PREC_HALF -> fp32), with the inverse built for twist a
  const CloverField &get(Prec prec, double a);
  void load(const void *h_clover, Prec host_prec, long Vh);
  void inverse_to_host(void *h_clovinv, Prec host_prec, double a2);  // (C^2 + a2)^-1 in packed order (what loadCloverQuda returns)
  void release();
  // site-major fp32 copy [V][72] (pool memory; caller frees with pool_free): the coarse-link build reads one record per site
  float *site_major_f32() const;
};

enum CloverMode { CLOVER_DIRECT = 0, CLOVER_INVERSE = 1, CLOVER_INVERSE_ADJ = 2 };

// out(parity) = [x +] k * S in   with S = C + i a gamma5 (DIRECT; a carries the dagger sign) or (C + i a0 gamma5)^-1 (INVERSE)
// or its conjugate transpose (INVERSE_ADJ).  out may alias in.
void clover_apply(SpinorField &out, const SpinorField &in, const CloverField &cl, int parity, CloverMode mode, double a, const SpinorField *x, double k);

}  // namespace qb
