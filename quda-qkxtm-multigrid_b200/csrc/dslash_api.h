// Host-side driver of the fine-grid hopping kernels: halo exchange + interior/boundary overlap.
#pragma once
#include "field.h"

namespace qb {

struct DslashParam;
struct PackParam;
struct StoreD; struct StoreS; struct StoreH;

template <typename Store> void launch_dslash_T(const DslashParam &p, int recon, bool twist_in, bool has_x, bool ghost, int block, cudaStream_t s, int clover = 0);
template <typename Store> void launch_pack_T(const PackParam &p, bool twist_in, cudaStream_t s);
template <typename Store> void launch_twist_T(void *out, float *out_norm, const void *in, const float *in_norm, long stride, int n, double pr, double qr, cudaStream_t s);

// (p, q) <-> p + i q gamma5
struct TwistCoef {
  double p = 1.0, q = 0.0;
  TwistCoef() {}
  TwistCoef(double p_, double q_) : p(p_), q(q_) {}
  bool trivial() const { return p == 1.0 && q == 0.0; }
};

// Lattice context: geometry, comm pattern, halo buffers, interior/boundary site lists.
struct Lattice {
  Geom geom;
  long V = 0;
  // partitioned-lattice bookkeeping (built lazily by setup_partition)
  int *interior_list[2] = {nullptr, nullptr};  // device, cb sites of parity p with no partitioned-boundary hop
  int *boundary_list[2] = {nullptr, nullptr};
  int n_interior[2] = {0, 0}, n_boundary[2] = {0, 0};
  bool interior_contiguous[2] = {false, false};
  int interior_begin[2] = {0, 0};
  // halo arenas per storage precision index (0: double, 1: single, 2: half)
  void *send_arena[3] = {nullptr, nullptr, nullptr};
  void *recv_arena[3] = {nullptr, nullptr, nullptr};
  size_t arena_bytes[3] = {0, 0, 0};
  size_t face_off[3][4][2];   // byte offset of the (d, dir) half-spinor block inside an arena
  size_t norm_off[3][4][2];   // byte offset of the norm block (half precision only)
  // Direct halo delivery (several ranks, receive arenas mapped into each other with CUDA IPC): the pack kernel stores the projected
  // faces straight into the NEIGHBOUR's ghost zone over NVLink, a one-warp kernel raises a sequence flag there, and the receiver's
  // boundary launch is preceded by a one-warp kernel that waits for the flags of its two neighbours -- no NCCL kernel, no copy of the
  // faces.  The receive arena is double buffered by sequence parity (a neighbour can be one hop ahead, never two: it needs my flag of
  // hop n + 1, raised after my boundary launch of hop n, before it may overwrite the buffer of hop n with hop n + 2).
  bool peer_halo[3] = {false, false, false};
  void *peer_recv[3][16];                    // receive arena of every rank as seen from this device
  unsigned long long *halo_flags[3] = {nullptr, nullptr, nullptr};   // my [4][2] arrival flags (inside my receive arena, behind the two buffers)
  unsigned long long halo_seq[3] = {0, 0, 0};
  int block_size = 0;         // 0: per-precision default

  void init(const int *X, int t_boundary_sign, double anisotropy);
  void setup_partition();
  void release();
};

// out(parity) = Cx x + Co D(parity <- 1-parity) Cin in     (see dslash.cuh)
// x may be nullptr.  Runs on rt().compute; when dimensions are partitioned the face pack + exchange
// run on rt().halo concurrently with the interior kernel, then the boundary sites are completed.
// clover_inv != nullptr: out = Cx x + S Co D in with S = (C + i a gamma5)^-1 of the output parity (clover_mode 1) or its conjugate
// transpose (2), CloverField::Ainv in the arithmetic type of the fields (int16 fields: Ainv16 + clover_norm) -- the twisted-clover even-odd
// hop in one launch
void apply_hop(Lattice &lat, const GaugeField &gauge, SpinorField &out, const SpinorField &in, int parity, bool dagger,
               TwistCoef cin, TwistCoef co, const SpinorField *x, TwistCoef cx, const void *clover_inv = nullptr, int clover_mode = 0,
               const float *clover_norm = nullptr);

// same on the contiguous checkerboard range [site_begin, site_begin + site_count) only, on stream s (unpartitioned lattices;
// used by the pipelined host path of dslashQuda)
void apply_hop_range(Lattice &lat, const GaugeField &gauge, SpinorField &out, const SpinorField &in, int parity, bool dagger,
                     TwistCoef cin, TwistCoef co, const SpinorField *x, TwistCoef cx, int site_begin, int site_count, cudaStream_t s);

void face_index_map(const Lattice &lat, int mu, int face_num, int parity, int *h_out);

// non-degenerate doublet: out = c1 * d (1 + i a gamma5 tau3 + b tau1) in + c2 * x   (x may be null or alias out)
void apply_ndeg_twist(SpinorField &out, const SpinorField &in, double a, double b, double d, double c1, const SpinorField *x, double c2);

// out = (p + i q gamma5) in on every site of the field
void apply_twist_field(SpinorField &out, const SpinorField &in, TwistCoef c);

}  // namespace qb
