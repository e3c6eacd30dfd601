// Host driver for the fine-grid hop: geometry, halo buffers, pack -> exchange -> interior/boundary.
// Replaces the reference's policy layer (/root/reference/lib/dslash_policy.cuh:148-297, :838-997) and the
// ghost bookkeeping of lib/cuda_color_spinor_field.cu:592-676, :1465-1860.
//
// Design difference (B200-first): the reference's interior kernel computes *partial* sums for
// boundary sites and per-dimension exterior kernels read them back and finish them.  Here the
// interior launch covers only the sites that need no remote data (complete results, written
// once), and after the halo has landed one boundary launch computes the remaining sites wholly.
// No read-modify-write of `out`, no "incomplete" bookkeeping, two launches regardless of how
// many dimensions are partitioned.
#include <vector>
#include "dslash.cuh"
#include "dslash_api.h"
#include "comm.h"

namespace qb {

static int prec_index(Prec p) { return p == PREC_DOUBLE ? 0 : (p == PREC_SINGLE ? 1 : 2); }
// measurement hook (timeHaloQudaB200): on partitioned lattices run only pack + exchange of a hop, no interior / boundary kernel
// Phases of a hop on a partitioned lattice.  Normal operation runs all three; the timing hook runs the exchange alone; the pipelined
// host path of dslashQuda (interface.cu) starts the exchange as soon as the boundary slices of the input have landed, runs the interior
// in slabs (apply_hop_range) while later slabs are still arriving, and the boundary sites at the end.
enum { PH_EXCHANGE = 1, PH_BOUNDARY = 2, PH_INTERIOR = 4, PH_SYNC = 8, PH_ALL = 7 };
static int g_phase_mask = PH_ALL;
void set_hop_phase(int mask) { g_phase_mask = mask; }
void set_halo_only(bool on) { g_phase_mask = on ? (PH_EXCHANGE | PH_SYNC) : PH_ALL; }
static int prec_store_bytes(Prec p) { return p == PREC_HALF ? 2 : (int)p; }

void Lattice::init(const int *X, int t_boundary_sign, double anisotropy) {
  release();
  V = 1;
  for (int d = 0; d < 4; d++) {
    if (X[d] < 2 || (X[d] & 1)) QB_ERROR("local lattice extent X[%d]=%d must be even and >= 2", d, X[d]);
    geom.X[d] = X[d];
    V *= X[d];
  }
  geom.Xh = X[0] / 2;
  geom.Vh = (int)(V / 2);
  Runtime &r = rt();
  for (int d = 0; d < 4; d++) {
    geom.faceVh[d] = (int)(V / X[d] / 2);
    geom.part[d] = (r.part_mask >> d) & 1;
  }
  // antiperiodic sign lives on the links of the last global time slice
  const bool last_t = r.coord[3] == r.grid[3] - 1, first_t = r.coord[3] == 0;
  geom.tb_fwd = (t_boundary_sign < 0 && last_t) ? -1 : 1;
  geom.tb_bwd = (t_boundary_sign < 0 && first_t) ? -1 : 1;
  geom.aniso = anisotropy;
  geom.aniso_f = (float)anisotropy;
  for (int i = 0; i < 3; i++) arena_bytes[i] = 0;
  if (r.part_mask) setup_partition();
}

void Lattice::release() {
  for (int p = 0; p < 2; p++) {
    if (interior_list[p]) cudaFree(interior_list[p]);
    if (boundary_list[p]) cudaFree(boundary_list[p]);
    interior_list[p] = boundary_list[p] = nullptr;
    n_interior[p] = n_boundary[p] = 0;
  }
  for (int i = 0; i < 3; i++) {
    if (peer_halo[i]) {
      // nobody may still be storing into (or reading flags from) an arena that is about to go away
      cudaDeviceSynchronize();
      comm_barrier();
      comm_ipc_unmap(peer_recv[i]);
      comm_barrier();
      peer_halo[i] = false;
    }
    if (send_arena[i]) comm_free_halo(send_arena[i]);
    if (recv_arena[i]) comm_free_halo(recv_arena[i]);
    send_arena[i] = recv_arena[i] = nullptr;
    halo_flags[i] = nullptr;
    arena_bytes[i] = 0;
  }
}

void Lattice::setup_partition() {
  const Geom &g = geom;
  // site lists (host build, once per lattice)
  for (int p = 0; p < 2; p++) {
    std::vector<int> in_l, bd_l;
    in_l.reserve(g.Vh);
    for (int cb = 0; cb < g.Vh; cb++) {
      const int za = cb / g.Xh, zb = za / g.X[1];
      int x[4];
      x[1] = za - zb * g.X[1];
      x[3] = zb / g.X[2];
      x[2] = zb - x[3] * g.X[2];
      x[0] = 2 * cb + ((x[1] + x[2] + x[3] + p) & 1) - za * g.X[0];
      bool bd = false;
      for (int d = 0; d < 4; d++)
        if (g.part[d] && (x[d] == 0 || x[d] == g.X[d] - 1)) bd = true;
      (bd ? bd_l : in_l).push_back(cb);
    }
    n_interior[p] = (int)in_l.size();
    n_boundary[p] = (int)bd_l.size();
    // a T-only (or any slowest-dimension-only) split leaves the interior as one contiguous cb range: no index list needed
    interior_contiguous[p] = !in_l.empty() && in_l.back() - in_l.front() + 1 == (int)in_l.size();
    interior_begin[p] = in_l.empty() ? 0 : in_l.front();
    if (n_interior[p]) {
      QB_CUDA(cudaMalloc((void **)&interior_list[p], sizeof(int) * n_interior[p]));
      QB_CUDA(cudaMemcpy(interior_list[p], in_l.data(), sizeof(int) * n_interior[p], cudaMemcpyHostToDevice));
    }
    if (n_boundary[p]) {
      QB_CUDA(cudaMalloc((void **)&boundary_list[p], sizeof(int) * n_boundary[p]));
      QB_CUDA(cudaMemcpy(boundary_list[p], bd_l.data(), sizeof(int) * n_boundary[p], cudaMemcpyHostToDevice));
    }
  }
}

// halo arena for one precision: per partitioned (d, dir) a [plane][faceVh] half-spinor block
// (12 reals per face site) and, for half precision, faceVh float norms.  256-byte aligned blocks.
static void ensure_arena(Lattice &lat, Prec prec) {
  const int pi = prec_index(prec);
  if (lat.arena_bytes[pi]) return;
  const Geom &g = lat.geom;
  size_t off = 0;
  auto align = [](size_t x) { return (x + 255) & ~(size_t)255; };
  for (int d = 0; d < 4; d++)
    for (int dir = 0; dir < 2; dir++) {
      lat.face_off[pi][d][dir] = off;
      lat.norm_off[pi][d][dir] = off;
      if (!g.part[d]) continue;
      off = align(off + (size_t)g.faceVh[d] * 12 * prec_store_bytes(prec));
      if (prec == PREC_HALF) {
        lat.norm_off[pi][d][dir] = off;
        off = align(off + (size_t)g.faceVh[d] * sizeof(float));
      }
    }
  lat.arena_bytes[pi] = off;
  lat.send_arena[pi] = comm_alloc_halo(off);
  // receive side: two buffers (sequence parity) + the arrival flags; mapped into the neighbours when there are real ranks
  const size_t recv_total = 2 * off + 256;
  lat.recv_arena[pi] = comm_alloc_halo(recv_total);
  QB_CUDA(cudaMemset(lat.recv_arena[pi], 0, recv_total));
  lat.halo_flags[pi] = (unsigned long long *)((char *)lat.recv_arena[pi] + 2 * off);
  lat.halo_seq[pi] = 0;
  lat.peer_halo[pi] = false;
  if (comm_peer_halo_wanted()) {
    lat.peer_halo[pi] = comm_ipc_map(lat.recv_arena[pi], lat.peer_recv[pi]);
    if (!lat.peer_halo[pi]) log_msg(1, "halo exchange: receive arenas could not be mapped between the ranks, faces travel through NCCL send / recv\n");
  }
}

template <typename Store>
static void hop_T(Lattice &lat, const GaugeField &gauge, SpinorField &out, const SpinorField &in, int parity, bool dagger,
                  TwistCoef cin, TwistCoef co, const SpinorField *x, TwistCoef cx, int range_begin = 0, int range_count = -1,
                  cudaStream_t range_stream = nullptr, const void *clover_inv = nullptr, int clover_mode = 0, const float *clover_norm = nullptr) {
  Runtime &r = rt();
  const Geom &g = lat.geom;
  DslashParam p;
  memset(&p, 0, sizeof(p));
  p.g = g;
  p.parity = parity;
  p.out = out.v; p.out_norm = out.norm;
  p.in = in.v; p.in_norm = in.norm;
  p.x = x ? x->v : nullptr; p.x_norm = x ? x->norm : nullptr;
  p.gauge_fwd = gauge.dir_ptr(parity, 0);
  p.gauge_bwd = gauge.dir_ptr(1 - parity, 0);
  p.stride = g.Vh;
  p.cin[0] = cin.p; p.cin[1] = cin.q;
  p.co[0] = co.p; p.co[1] = co.q;
  p.cx[0] = cx.p; p.cx[1] = cx.q;
  p.sgn_fwd = dagger ? 1.0 : -1.0;
  p.clover_inv = clover_inv;
  p.clover_norm = clover_norm;
  if (clover_inv && Store::prec == PREC_HALF && !clover_norm) QB_ERROR("apply_hop: int16 fields take the int16 + norm copy of the inverse clover blocks");
  const int clover = clover_inv ? clover_mode : 0;
  const bool twist_in = !cin.trivial();
  const bool has_x = x != nullptr;

  // default launch geometry: 128 threads (fp32 / int16), 64 threads (fp64); lat.block_size > 0 overrides
  const int block = lat.block_size > 0 ? lat.block_size : (Store::prec == PREC_DOUBLE ? 64 : 128);
  const bool partitioned = g.part[0] || g.part[1] || g.part[2] || g.part[3];
  if (out.nbatch != in.nbatch || (x && x->nbatch != out.nbatch)) QB_ERROR("apply_hop: batch sizes differ (out %d, in %d)", out.nbatch, in.nbatch);
  if (out.nflavor != in.nflavor || (x && x->nflavor != out.nflavor)) QB_ERROR("apply_hop: flavour counts differ");
  // members handled by one launch: the fields of a batch, or the two flavours of a doublet (which sit inside the parity block)
  const int nmember = out.nflavor == 2 ? 2 : out.nbatch;
  if (clover && (nmember > 1 || twist_in)) QB_ERROR("apply_hop: the fused clover epilogue takes single fields without input twist");
  const size_t st_in = out.nflavor == 2 ? in.flavor_bytes() : in.batch_bytes, st_out = out.nflavor == 2 ? out.flavor_bytes() : out.batch_bytes;
  const size_t st_x = x ? (out.nflavor == 2 ? x->flavor_bytes() : x->batch_bytes) : 0;
  if (nmember > 1 && partitioned) {
    // partitioned lattice: the halo arena and the pack kernel hold one field, so the members go through the overlapped
    // pack / exchange / interior / boundary path one after the other (no link sharing between members here)
    auto member_view = [](SpinorField &w, const SpinorField &f, int c, size_t stride) {
      w.prec = f.prec; w.nparity = 1; w.ncomplex = f.ncomplex; w.nspin = f.nspin; w.ncolor = f.ncolor; w.Vh = f.Vh;
      w.v = (char *)f.v + (size_t)c * stride; w.norm = nullptr; w.parity_bytes = f.flavor_bytes(); w.owner = false; w.nbatch = 1; w.nflavor = 1;
    };
    for (int c = 0; c < nmember; c++) {
      SpinorField oc, ic, xc;
      member_view(oc, out, c, st_out); member_view(ic, in, c, st_in);
      if (x) member_view(xc, *x, c, st_x);
      hop_T<Store>(lat, gauge, oc, ic, parity, dagger, cin, co, x ? &xc : nullptr, cx, range_begin, range_count, range_stream);
    }
    return;
  }
  if (!partitioned) {
    p.site_begin = range_begin; p.site_count = range_count < 0 ? g.Vh : range_count; p.site_list = nullptr;
    cudaStream_t st = range_stream ? range_stream : r.compute;
    if (nmember > 1) {
      // members in groups of DSLASH_BATCH_MAX: one launch per group, links read from HBM once per group
      for (int first = 0; first < nmember; first += 12) {
        DslashParam pb = p;
        pb.nbatch = std::min(12, nmember - first);
        pb.batch_in = (long)st_in; pb.batch_out = (long)st_out; pb.batch_x = (long)st_x;
        pb.in = (const char *)in.v + (size_t)first * st_in;
        pb.out = (char *)out.v + (size_t)first * st_out;
        pb.x = x ? (const char *)x->v + (size_t)first * st_x : nullptr;
        if (pb.nbatch == 1) { pb.nbatch = 0; launch_dslash_T<Store>(pb, gauge.recon, twist_in, has_x, false, block, st); }
        else launch_dslash_T<Store>(pb, gauge.recon, twist_in, has_x, false, block, st);
      }
      return;
    }
    launch_dslash_T<Store>(p, gauge.recon, twist_in, has_x, false, block, st, clover);
    return;
  }
  if (range_count >= 0) {
    // slab of the interior (pipelined host path): only for a T-only partition, where the interior is the contiguous block of the
    // time slices 1 .. T-2 and a slab needs nothing but its own and the two adjacent slabs of `in`
    if (g.part[0] || g.part[1] || g.part[2] || !lat.interior_contiguous[parity]) QB_ERROR("apply_hop_range on a partitioned lattice needs a T-only partition");
    const int lo = std::max(range_begin, lat.interior_begin[parity]);
    const int hi = std::min(range_begin + range_count, lat.interior_begin[parity] + lat.n_interior[parity]);
    if (hi > lo) {
      p.site_begin = lo; p.site_count = hi - lo; p.site_list = nullptr;
      launch_dslash_T<Store>(p, gauge.recon, twist_in, has_x, false, block, range_stream ? range_stream : r.compute, clover);
    }
    return;
  }

  const int pi = prec_index(Store::prec);
  ensure_arena(lat, Store::prec);
  char *send = (char *)lat.send_arena[pi], *recv = (char *)lat.recv_arena[pi];
  const bool self = comm_self_exchange();
  const bool peer = lat.peer_halo[pi];
  const int mask = g_phase_mask;
  if (peer && (mask & PH_EXCHANGE)) lat.halo_seq[pi]++;   // identical on all ranks: hops are collective
  const unsigned long long seq = lat.halo_seq[pi];
  const size_t buf = peer ? (size_t)(seq & 1) * lat.arena_bytes[pi] : 0;
  HaloFlags sig{}, wt{};
  sig.seq = wt.seq = seq;
  // self-exchange (single rank, forced partitioning): our back face *is* our forward ghost -> alias, no copy
  PackParam pk;
  memset(&pk, 0, sizeof(pk));
  pk.g = g;
  pk.parity = 1 - parity;
  pk.in = in.v; pk.in_norm = in.norm;
  pk.stride = g.Vh;
  pk.cin[0] = cin.p; pk.cin[1] = cin.q;
  pk.sgn_fwd = p.sgn_fwd;
  int off = 0;
  for (int d = 0; d < 4; d++) {
    pk.thread_off[d] = off;
    if (g.part[d]) off += 2 * g.faceVh[d];
    for (int dir = 0; dir < 2; dir++) {
      pk.send[d][dir] = send + lat.face_off[pi][d][dir];
      pk.send_norm[d][dir] = (float *)(send + lat.norm_off[pi][d][dir]);
      // ghost[d][0] comes from the backward neighbour = its forward face (dir 1); ghost[d][1] = nbr's back face
      const char *src = self ? send : recv;
      p.ghost[d][dir] = src + lat.face_off[pi][d][1 - dir];
      p.ghost_norm[d][dir] = (const float *)(src + lat.norm_off[pi][d][1 - dir]);
      if (!self) {  // received blocks are laid out by *receiving* slot: recv[d][0] <- from back, recv[d][1] <- from fwd
        p.ghost[d][dir] = recv + buf + lat.face_off[pi][d][dir];
        p.ghost_norm[d][dir] = (const float *)(recv + buf + lat.norm_off[pi][d][dir]);
      }
      if (peer && g.part[d]) {
        // my face `dir` (0: slice x_d = 0, travels backward; 1: slice X_d - 1, travels forward) goes straight into the neighbour's
        // receive slot (d, 1 - dir) of this hop's buffer; its arrival flag (d, 1 - dir) sits behind the neighbour's two buffers
        char *nb = (char *)lat.peer_recv[pi][comm_neighbor_rank(d, dir)];
        pk.send[d][dir] = nb + buf + lat.face_off[pi][d][1 - dir];
        pk.send_norm[d][dir] = (float *)(nb + buf + lat.norm_off[pi][d][1 - dir]);
        sig.p[sig.n++] = (unsigned long long *)(nb + 2 * lat.arena_bytes[pi]) + (d * 2 + (1 - dir));
        wt.p[wt.n++] = lat.halo_flags[pi] + (d * 2 + dir);
      }
    }
    p.gauge_ghost[d] = gauge.ghost[d] ? (const char *)gauge.ghost[d] + (size_t)(1 - parity) * gauge.recon * gauge.store_bytes() * g.faceVh[d] : nullptr;
    if (g.part[d] && !gauge.ghost[d]) QB_ERROR("gauge field has no ghost links for partitioned dimension %d", d);
  }
  pk.thread_off[4] = off;

  // halo stream (high priority): wait until `in` is complete on the compute stream, pack, exchange, and then
  // compute the boundary sites right there -- they are few (2 faces of X*Y*Z/2 sites for a T split), so
  // the launch is latency-bound and hides completely under the interior kernel running on the compute stream.
  // Interior and boundary launches write disjoint sites of `out`.
  if (mask & PH_EXCHANGE) {
    QB_CUDA(cudaEventRecord(r.ev_in_ready, r.compute));
    QB_CUDA(cudaStreamWaitEvent(r.halo, r.ev_in_ready, 0));
    launch_pack_T<Store>(pk, twist_in, r.halo);
    if (peer) {
      comm_halo_signal(sig, r.halo);
      if (mask & PH_SYNC) comm_halo_wait(wt, r.halo);   // timing hook: until the neighbours' faces are here
    } else if (!self) comm_exchange_halo(lat, pi, r.halo);
    if (!(mask & PH_BOUNDARY)) {
      QB_CUDA(cudaEventRecord(r.ev_halo_done, r.halo));
      if (mask & PH_SYNC) QB_CUDA(cudaStreamWaitEvent(r.compute, r.ev_halo_done, 0));
    }
  }

  const int np = parity;
  if (mask & PH_BOUNDARY) {
    DslashParam pb = p;
    // with the exchange in this call: on the halo stream right behind it; alone (pipelined host path): on the compute stream, after
    // the exchange started earlier has finished
    cudaStream_t bs = (mask & PH_EXCHANGE) ? r.halo : r.compute;
    if (!(mask & PH_EXCHANGE)) QB_CUDA(cudaStreamWaitEvent(r.compute, r.ev_halo_done, 0));
    if (peer) comm_halo_wait(wt, bs);   // the neighbours' faces of this hop have landed
    if (lat.n_boundary[np]) {
      pb.site_begin = 0; pb.site_count = lat.n_boundary[np]; pb.site_list = lat.boundary_list[np];
      launch_dslash_T<Store>(pb, gauge.recon, twist_in, has_x, true, block, bs, clover);
    }
    if (mask & PH_EXCHANGE) QB_CUDA(cudaEventRecord(r.ev_halo_done, r.halo));
  }

  // interior sites: everything that needs no remote data, overlapping with pack + exchange + boundary
  if ((mask & PH_INTERIOR) && lat.n_interior[np]) {
    p.site_count = lat.n_interior[np];
    if (lat.interior_contiguous[np]) { p.site_begin = lat.interior_begin[np]; p.site_list = nullptr; }
    else { p.site_begin = 0; p.site_list = lat.interior_list[np]; }
    launch_dslash_T<Store>(p, gauge.recon, twist_in, has_x, false, block, r.compute, clover);
  }
  if ((mask & PH_EXCHANGE) && (mask & PH_BOUNDARY)) QB_CUDA(cudaStreamWaitEvent(r.compute, r.ev_halo_done, 0));
}

void apply_hop(Lattice &lat, const GaugeField &gauge, SpinorField &out, const SpinorField &in, int parity, bool dagger,
               TwistCoef cin, TwistCoef co, const SpinorField *x, TwistCoef cx, const void *clover_inv, int clover_mode, const float *clover_norm) {
  if (out.prec != in.prec || out.prec != gauge.prec || (x && x->prec != out.prec))
    QB_ERROR("apply_hop: precision mismatch (out %d, in %d, gauge %d)", (int)out.prec, (int)in.prec, (int)gauge.prec);
  if (out.Vh != lat.geom.Vh || in.Vh != lat.geom.Vh) QB_ERROR("apply_hop: field volume does not match the lattice");
  if (out.v == in.v) QB_ERROR("apply_hop: out and in must not alias");
  if (out.prec == PREC_DOUBLE) hop_T<StoreD>(lat, gauge, out, in, parity, dagger, cin, co, x, cx, 0, -1, nullptr, clover_inv, clover_mode, clover_norm);
  else if (out.prec == PREC_SINGLE) hop_T<StoreS>(lat, gauge, out, in, parity, dagger, cin, co, x, cx, 0, -1, nullptr, clover_inv, clover_mode, clover_norm);
  else hop_T<StoreH>(lat, gauge, out, in, parity, dagger, cin, co, x, cx, 0, -1, nullptr, clover_inv, clover_mode, clover_norm);
}

// face index -> checkerboard index table of the pack kernel (for the index-parity tests)
__global__ void face_map_kernel(int *out, Geom g, int mu, int face_num, int parity) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= g.faceVh[mu]) return;
  const int slice = face_num ? g.X[mu] - 1 : 0;
  int cb;
  if (mu == 0) cb = face_to_cb<0>(f, slice, parity, g);
  else if (mu == 1) cb = face_to_cb<1>(f, slice, parity, g);
  else if (mu == 2) cb = face_to_cb<2>(f, slice, parity, g);
  else cb = face_to_cb<3>(f, slice, parity, g);
  out[f] = cb;
}

void apply_hop_range(Lattice &lat, const GaugeField &gauge, SpinorField &out, const SpinorField &in, int parity, bool dagger,
                     TwistCoef cin, TwistCoef co, const SpinorField *x, TwistCoef cx, int site_begin, int site_count, cudaStream_t s) {
  if (out.prec != in.prec || out.prec != gauge.prec) QB_ERROR("apply_hop_range: precision mismatch");
  if (out.prec == PREC_DOUBLE) hop_T<StoreD>(lat, gauge, out, in, parity, dagger, cin, co, x, cx, site_begin, site_count, s);
  else if (out.prec == PREC_SINGLE) hop_T<StoreS>(lat, gauge, out, in, parity, dagger, cin, co, x, cx, site_begin, site_count, s);
  else hop_T<StoreH>(lat, gauge, out, in, parity, dagger, cin, co, x, cx, site_begin, site_count, s);
}

void face_index_map(const Lattice &lat, int mu, int face_num, int parity, int *h_out) {
  const int n = lat.geom.faceVh[mu];
  int *d;
  QB_CUDA(cudaMalloc((void **)&d, sizeof(int) * n));
  face_map_kernel<<<div_up(n, 128), 128, 0, rt().compute>>>(d, lat.geom, mu, face_num, parity);
  QB_CHECK_LAUNCH();
  QB_CUDA(cudaMemcpyAsync(h_out, d, sizeof(int) * n, cudaMemcpyDeviceToHost, rt().compute));
  QB_CUDA(cudaStreamSynchronize(rt().compute));
  QB_CUDA(cudaFree(d));
}

// out = c1 * d (1 + i a gamma5 tau3 + b tau1) in + c2 * x on a flavour-doublet field (x may be null, may alias out)
template <typename Store>
__global__ void __launch_bounds__(128) ndeg_twist_kernel(void *out, const void *in, const void *x, long Vh, long nsites, size_t parity_bytes, size_t flavor_bytes,
                                                         double a_, double b_, double d_, double c1_, double c2_) {
  typedef typename Store::real real;
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nsites) return;
  const int parity = (int)(t / Vh);
  const long cb = t - (long)parity * Vh;
  const size_t off = parity_bytes * parity;
  cplx<real> f1[12], f2[12], o1[12], o2[12];
  Store::template load<12, false>(f1, (const char *)in + off, nullptr, Vh, cb);
  Store::template load<12, false>(f2, (const char *)in + off + flavor_bytes, nullptr, Vh, cb);
  const real a = (real)a_, b = (real)b_, s1 = (real)(c1_ * d_);
#pragma unroll
  for (int k = 0; k < 12; k++) {
    const real a5 = k < 6 ? a : -a;   // gamma5 = diag(1, 1, -1, -1) in the internal chiral basis
    o1[k] = cplx<real>(s1 * (f1[k].re - a5 * f1[k].im + b * f2[k].re), s1 * (f1[k].im + a5 * f1[k].re + b * f2[k].im));
    o2[k] = cplx<real>(s1 * (f2[k].re + a5 * f2[k].im + b * f1[k].re), s1 * (f2[k].im - a5 * f2[k].re + b * f1[k].im));
  }
  if (x) {
    const real c2 = (real)c2_;
    cplx<real> x1[12], x2[12];
    Store::template load<12, false>(x1, (const char *)x + off, nullptr, Vh, cb);
    Store::template load<12, false>(x2, (const char *)x + off + flavor_bytes, nullptr, Vh, cb);
#pragma unroll
    for (int k = 0; k < 12; k++) { o1[k].re += c2 * x1[k].re; o1[k].im += c2 * x1[k].im; o2[k].re += c2 * x2[k].re; o2[k].im += c2 * x2[k].im; }
  }
  Store::template store<12>((char *)out + off, nullptr, Vh, cb, o1);
  Store::template store<12>((char *)out + off + flavor_bytes, nullptr, Vh, cb, o2);
}

void apply_ndeg_twist(SpinorField &out, const SpinorField &in, double a, double b, double d, double c1, const SpinorField *x, double c2) {
  if (in.nflavor != 2 || out.nflavor != 2 || (x && x->nflavor != 2)) QB_ERROR("apply_ndeg_twist needs flavour-doublet fields");
  if (out.prec != in.prec || out.Vh != in.Vh || out.nparity != in.nparity || (x && (x->prec != in.prec || x->nparity != in.nparity))) QB_ERROR("apply_ndeg_twist: field mismatch");
  const long n = in.Vh * in.nparity;
  cudaStream_t s = rt().compute;
  if (in.prec == PREC_DOUBLE) ndeg_twist_kernel<StoreD><<<div_up(n, 128), 128, 0, s>>>(out.v, in.v, x ? x->v : nullptr, in.Vh, n, in.parity_bytes, in.flavor_bytes(), a, b, d, c1, c2);
  else if (in.prec == PREC_SINGLE) ndeg_twist_kernel<StoreS><<<div_up(n, 128), 128, 0, s>>>(out.v, in.v, x ? x->v : nullptr, in.Vh, n, in.parity_bytes, in.flavor_bytes(), a, b, d, c1, c2);
  else QB_ERROR("apply_ndeg_twist: fp32 / fp64 fields only");
  QB_CHECK_LAUNCH();
}

void apply_twist_field(SpinorField &out, const SpinorField &in, TwistCoef c) {
  if (out.prec != in.prec || out.Vh != in.Vh || out.nparity != in.nparity) QB_ERROR("apply_twist_field: field mismatch");
  Runtime &r = rt();
  const int n = (int)(in.Vh);
  for (int p = 0; p < in.nparity; p++) {
    void *o = out.parity_ptr(p); float *on = out.parity_norm(p);
    const void *i = in.parity_ptr(p); const float *inn = in.parity_norm(p);
    if (in.prec == PREC_DOUBLE) launch_twist_T<StoreD>(o, on, i, inn, in.Vh, n, c.p, c.q, r.compute);
    else if (in.prec == PREC_SINGLE) launch_twist_T<StoreS>(o, on, i, inn, in.Vh, n, c.p, c.q, r.compute);
    else launch_twist_T<StoreH>(o, on, i, inn, in.Vh, n, c.p, c.q, r.compute);
  }
}

}  // namespace qb
