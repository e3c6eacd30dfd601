// Inter-grid transfer operator of the multigrid (replaces /root/reference/lib/transfer.cpp,
// lib/transfer_util.cu, lib/prolongator.cu, lib/restrictor.cu; include/transfer.h).
//
//   P : fine(x, k)      = sum_j V(x, k, j) coarse(X(x), chi(k), j)                  (prolongator.cu:41-56)
//   R : coarse(X, S, j) = sum_{x in X} sum_{k: chi(k)=S} conj(V(x, k, j)) fine(x, k)  (restrictor.cu:49-125)
// with k = (spin, colour) of the fine level, chi(k) = spin / spin_block_size the chirality, X(x) the
// aggregate (coarse site) containing x, j the null-vector index (= coarse colour).
//
// Everything lives in HBM in fp32: V as [parity][k][j-pair][cb site] float4 planes (so that P streams V
// with 128-bit coalesced loads), the geometry maps as int arrays.  V is block-orthonormalised on the
// device with fp64 accumulation (the reference does this on the CPU, transfer_util.cu:327-363).
#pragma once
#include <vector>
#include "field.h"

namespace qb {

struct LevelGeom {
  int X[4];        // local extents of this level
  int Xh;
  long Vh;
  int part[4];     // dimension partitioned over ranks (or forced self-exchange): neighbours across it live in ghost buffers
  int faceVh[4];   // checkerboard face volume per dimension
  long V() const { return 2 * Vh; }
  void set(const int *x) {
    for (int d = 0; d < 4; d++) X[d] = x[d];
    Xh = x[0] / 2;
    Vh = (long)x[0] * x[1] * x[2] * x[3] / 2;
    for (int d = 0; d < 4; d++) {
      part[d] = (rt().part_mask >> d) & 1;
      faceVh[d] = (int)(Vh / x[d]);
    }
  }
  bool partitioned() const { return part[0] || part[1] || part[2] || part[3]; }
};

class Transfer {
 public:
  int nvec = 0;
  int fine_nspin = 4, fine_ncolor = 3, Nf = 12;  // Nf = fine complex components per site
  int spin_bs = 2;                                // fine spins per chirality block
  int geo_bs[4] = {1, 1, 1, 1};
  int block_sites = 1;
  LevelGeom fine, coarse;
  float *V = nullptr;   // [parity][k][nvec/2][Vh_f] float4
  void *V16 = nullptr;  // optional fp16 copy (same indexing, 4 halves per element pair) used by P / R when set
  void enable_half_v();
  int *f2c = nullptr;   // [parity*Vh_f + cb]      -> coarse full index (parity_c*Vh_c + cb_c)
  int *c2f = nullptr;   // [coarse full index][block_sites] -> fine full index
  // ghost copies of V on the neighbours' boundary slices, [d][0]: backward neighbour's x_d = X_d-1 slice, [d][1]: forward
  // neighbour's x_d = 0 slice; layout [parity][k][nvec/2][faceVh] float4.  Needed by the coarse-link build only.
  float *Vghost[4][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
  mutable long long flops = 0;

  // B: nvec near-null vectors (full fine fields, fp32).  geo_bs is adjusted in place exactly as the
  // reference does (halved until it divides the lattice and leaves an even coarse extent, transfer.cpp:31-44).
  Transfer(const std::vector<SpinorField *> &B, int nvec, int *geo_bs, int spin_bs, const int *fine_X);
  ~Transfer();
  Transfer(const Transfer &) = delete;

  SpinorField *new_coarse_field() const { return new SpinorField(coarse.Vh, 2, PREC_SINGLE, 2, nvec); }
  void P(SpinorField &fine_out, const SpinorField &coarse_in) const;
  // parity >= 0 (preconditioned coarsening, lib/transfer.cpp:270-348 with a parity site subset): the fine field lives on that parity only
  // (a single-parity field, or that block of a full field); the other parity counts as zero and its half of V is never read
  void R(SpinorField &coarse_out, const SpinorField &fine_in, int parity = -1) const;
  // several vectors per pass over V (block multigrid); accumulate: fine_out += P coarse_in
  void P_multi(SpinorField *const *fine_out, const SpinorField *const *coarse_in, int n, bool accumulate, int parity = -1) const;
  void R_multi(SpinorField *const *coarse_out, const SpinorField *const *fine_in, int n, int parity = -1) const;
  size_t v_bytes() const { return (size_t)2 * fine.Vh * Nf * nvec * 8; }
 private:
  void exchange_v_ghost();
};

}  // namespace qb
