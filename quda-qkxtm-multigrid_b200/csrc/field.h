// Device-resident lattice fields and their HBM layout.
//
// Layout rule (all fields): structure-of-arrays of 16-byte "planes", so that one warp reading one
// plane of 32 consecutive checkerboard sites issues a single fully coalesced 512-byte request with
// 128-bit loads per thread:
//     spinor  : [parity][plane][cb site]      plane = 16 B = 1 complex<double> | 2 complex<float> | 4 complex<int16>
//     gauge   : [parity][mu][plane][cb site]  recon 12/8: 16-B planes (fp64/fp32), 8-B planes (half);
//                                             recon 18: 9 planes of one complex each
//     half    : additionally one float norm per site, [parity][cb site]
// Internal gamma basis is DeGrand-Rossi (chiral): the twist is diagonal, chirality blocks coincide
// with the multigrid coarse spin, and the host test basis needs no rotation.  (The reference keeps
// UKQCD internally and rotates on every copy: /root/reference/include/color_spinor_field.h:174,
// lib/copy_color_spinor.cuh:49-92.)
#pragma once
#include "common.h"

namespace qb {

// Local lattice geometry, passed by value to kernels.
struct Geom {
  int X[4];       // local extents (x,y,z,t), all even
  int Xh;         // X[0]/2
  int Vh;         // checkerboard volume
  int part[4];    // dimension partitioned -> boundary hops read the ghost zone
  int faceVh[4];  // checkerboard face volume per dimension
  int tb_fwd;     // -1 if forward T links on the local slice t=T-1 carry the antiperiodic sign
  int tb_bwd;     // -1 if backward T links used by sites at t=0 carry it
  double aniso;   // anisotropy (recon 12/8: scale of the reconstructed row for spatial links); double: fp64 fields must see the exact value
  float aniso_f;  // the same rounded once on the host, for the fp32 / int16 kernels (no fp64 conversion in their inner code)
};

struct SpinorField {
  Prec prec = PREC_DOUBLE;
  int nparity = 1;       // 1: single-parity field, 2: full field [even | odd]
  int ncomplex = 12;     // complex components per site (nspin * ncolor)
  int nspin = 4, ncolor = 3;
  long Vh = 0;           // sites per parity
  void *v = nullptr;     // planes
  float *norm = nullptr; // half precision only
  size_t parity_bytes = 0;
  bool owner = true;
  // batch of nbatch fields of identical shape in ONE allocation, batch_bytes apart ([rhs][parity][plane][site]); the members above
  // describe member 0.  Only the fine-grid operators (apply_hop and everything built on it) understand batches: one launch covers all
  // members and the links are fetched from HBM once.  BLAS works on member views (member()).
  int nbatch = 1;
  size_t batch_bytes = 0;
  // flavour doublet (non-degenerate twisted mass): nflavor = 2 flavours side by side INSIDE every parity block,
  // [parity][flavour][plane][site] (the host layout of the reference's doublet fields); parity_bytes covers both flavours, so BLAS
  // sees one flat vector and parity views stay contiguous; the hop kernel treats the flavours as a batch of two.
  int nflavor = 1;
  size_t flavor_bytes() const { return parity_bytes / nflavor; }

  SpinorField() {}
  SpinorField(long Vh, int nparity, Prec prec, int nspin = 4, int ncolor = 3, int nbatch = 1, int nflavor = 1);
  ~SpinorField();
  SpinorField(const SpinorField &) = delete;
  SpinorField &operator=(const SpinorField &) = delete;

  int planes() const { return ncomplex * 2 * (int)(prec == PREC_HALF ? 2 : prec) / 16; }
  size_t bytes() const { return parity_bytes * nparity; }
  long reals() const { return (long)Vh * ncomplex * 2 * nparity * nflavor; }
  void *parity_ptr(int p) const { return (char *)v + parity_bytes * (nparity == 2 ? p : 0); }
  float *parity_norm(int p) const { return norm ? norm + Vh * (nparity == 2 ? p : 0) : nullptr; }
  // non-owning view of one parity of a full field (or the field itself if single parity)
  void view_parity(SpinorField &dst, int p) const;
  // non-owning view of member c of a batch (nbatch = 1)
  void member(SpinorField &dst, int c) const;
  void zero(cudaStream_t s);
};

struct GaugeField {
  Prec prec = PREC_DOUBLE;
  int recon = 18;
  long Vh = 0;
  void *data = nullptr;          // [parity][mu][plane][Vh]
  void *ghost[4] = {nullptr, nullptr, nullptr, nullptr};  // [parity][plane][faceVh[d]] : U_d at x_d = X_d-1 of the backward neighbour
  int store_bytes() const { return prec == PREC_HALF ? 2 : (int)prec; }
  int reals_per_plane() const { return recon == 18 ? 2 : (prec == PREC_DOUBLE ? 2 : 4); }
  int planes() const { return recon / reals_per_plane(); }
  size_t plane_elem_bytes() const { return (size_t)reals_per_plane() * store_bytes(); }
  size_t dir_bytes() const { return (size_t)recon * store_bytes() * Vh; }  // one (parity, mu) block
  size_t bytes() const { return 8 * dir_bytes(); }
  void *dir_ptr(int parity, int mu) const { return (char *)data + (size_t)(parity * 4 + mu) * dir_bytes(); }
  GaugeField(long Vh, Prec prec, int recon);
  ~GaugeField();
  GaugeField(const GaugeField &) = delete;
  GaugeField &operator=(const GaugeField &) = delete;
};

// host <-> device marshalling (interface_quda.cpp:521-692 loadGaugeQuda; cuda_color_spinor_field.cu:513-552)
enum HostBasis { BASIS_DEGRAND_ROSSI = 0, BASIS_UKQCD = 1 };
enum HostSpinorOrder { ORDER_SPIN_COLOR = 0, ORDER_COLOR_SPIN = 1 };

// h_gauge: QDP order void*[4], each [parity][cb][row][col][re,im] in host_prec
// host link orders (include/gauge_field_order.h of the reference): QDP = void*[4], [mu][parity][x_cb][row][col]; MILC = one array
// [parity][x_cb][mu][row][col] (:1028-1070); CPS = MILC indexing with the colour matrix transposed and multiplied by the anisotropy (:1076-1135)
enum HostGaugeOrder { GAUGE_ORDER_QDP = 0, GAUGE_ORDER_MILC = 1, GAUGE_ORDER_CPS = 2 };
void import_gauge(GaugeField &g, void *const *h_gauge, Prec host_prec, const Geom &geom, cudaStream_t s, HostGaugeOrder order = GAUGE_ORDER_QDP);
void export_gauge(void *const *h_gauge, const GaugeField &g, Prec host_prec, const Geom &geom, cudaStream_t s, HostGaugeOrder order = GAUGE_ORDER_QDP);
// fills g.ghost[d] for the partitioned dimensions from the local field (self-exchange) or the neighbour (NCCL)
void exchange_gauge_ghost(GaugeField &g, const Geom &geom, cudaStream_t s);

// h: [parity][cb][spin][color][re,im] (or color-spin) in host_prec; nparity taken from the field
void import_spinor(SpinorField &f, const void *h, Prec host_prec, HostBasis basis, HostSpinorOrder order, cudaStream_t s);
void export_spinor(void *h, const SpinorField &f, Prec host_prec, HostBasis basis, HostSpinorOrder order, cudaStream_t s);
void import_spinor_range(SpinorField &f, const void *stage_dev, Prec host_prec, HostBasis basis, HostSpinorOrder order, long begin, long count, cudaStream_t s);
void export_spinor_range(void *stage_dev, const SpinorField &f, Prec host_prec, HostBasis basis, HostSpinorOrder order, long begin, long count, cudaStream_t s);
void *staging(size_t bytes);
// precision change / copy between resident fields of identical geometry
void copy_spinor(SpinorField &dst, const SpinorField &src, cudaStream_t s);

// host <-> device in the generic order [parity][x_cb][component][re, im] (fp32, any number of components)
void import_generic(SpinorField &f, const float *h, cudaStream_t s);
void export_generic(float *h, const SpinorField &f, cudaStream_t s);

}  // namespace qb
