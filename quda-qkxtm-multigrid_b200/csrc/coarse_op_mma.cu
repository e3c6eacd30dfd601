// Galerkin coarse-link build of the finest level on the tensor cores (tcgen05 / TMEM):
//
//     L^c_d(X) = sum_{x in X} V(x)^dag  W_d(x),      W_d(x) = L_d(x) V(x + e_d)          (reference: CPU only,
//     lib/coarse_op.cuh computeUV :59-125, computeVUV :487-599, computeCoarseLocal :670-711; see coarse_op.cu)
//
// W_d(x) (12 x N complex, the reference's "UV") costs 3.5 kFMA per (site, direction) and is formed by CUDA cores
// straight into the shared-memory operand of the MMA; the contraction V^dag W (110 kflop per site and direction, 94 % of
// the work) runs as  D[2N x N] += A[2N x 24] B[N x 24]^T  with
//   A[2r + ri][(k, re/im)] : row 2r   = ( Re V(x)[k][r],  Im V(x)[k][r])   -> Re of conj(V) * W
//                            row 2r+1 = (-Im V(x)[k][r],  Re V(x)[k][r])   -> Im of conj(V) * W,   zero unless chirality(k) = S'(r)
//   B[n][(k, re/im)]       : ( Re W_d(x)[k][n], Im W_d(x)[k][n] )
// in tf32 with both operands split hi + lo (hi*hi + lo*hi + hi*lo accumulated in fp32), i.e. fp32-grade links.
// The tensor core rounds EVERY accumulation toward zero (whatever the size of the addend), and the sums here are coherent
// (V^dag V is positive), so a long chain carries a systematic relative error of ~ chain_length * 3e-8.  Therefore
//   * the large hi*hi products and the small cross terms go to separate accumulators (the cross terms would otherwise
//     cost the big accumulator two more truncations per K step for nothing),
//   * a chain covers only `chunk` (8) sites = 24 MMAs; the epilogue adds the partial sums to running fp32 totals held in
//     registers (round to nearest), one accumulator row per thread; two TMEM buffers so that draining overlaps the MMAs.
//
// One persistent CTA per SM walks aggregates; warp roles:
//   warp 0    : producer  - bulk copies of V(x), V(x + e_d) and the pre-multiplied link of every (aggregate, direction, site)
//   warp 1    : MMA issue (one elected lane, warp-uniform control flow) + TMEM allocation
//   warps 2-9 : builders  - A operand (into tensor memory) and B tile of every work item, hi / lo split (two groups of 4 warps)
//   warps 10-13: epilogue  - tcgen05.ld -> running totals -> coarse links (plain stores, no read-modify-write)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "coarse.h"
#include "dslash.cuh"
#include "tc05.cuh"

namespace qb {

using namespace tc;

float *decompress_gauge(const GaugeField &gf, const Geom &g);  // coarse_op.cu
float *decompress_ghost_links(const GaugeField &gf, const Geom &g, int mu);  // coarse_op.cu: [parity][faceVh][18], cudaMalloc'ed

// ---- site-major copies of the operands: everything a site needs is one contiguous block -> one bulk copy ------------------
// Vs[fs][k][j] complex  (12 * nvec * 8 bytes per fine site, fs = parity * Vh + cb)
__global__ void pack_v_site_major_kernel(float2 *Vs, const float4 *V, int Nf, int nvec, long Vh) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nvh = nvec / 2;
  const long n = 2 * Vh * Nf * nvh;
  if (t >= n) return;
  // thread order = source order ([parity][k][jp][cb], coalesced reads); writes are 16-byte pieces
  const long cb = t % Vh;
  long u = t / Vh;
  const int jp = (int)(u % nvh); u /= nvh;
  const int k = (int)(u % Nf);
  const int parity = (int)(u / Nf);
  const float4 v = V[t];
  float4 *dst = (float4 *)(Vs + (((size_t)parity * Vh + cb) * Nf + k) * nvec + 2 * jp);
  *dst = v;
}

// index of a site inside the checkerboarded face orthogonal to mu (3-d lexicographic of the remaining coordinates >> 1)
__device__ __forceinline__ int face_index_rt(int mu, const int *x, const Geom &g) {
  const int d0 = mu == 0 ? 1 : 0, d1 = mu <= 1 ? 2 : 1, d2 = mu <= 2 ? 3 : 2;
  return (int)((x[d0] + g.X[d0] * (x[d1] + (long)g.X[d1] * x[d2])) >> 1);
}

// Us[fs][d][10] complex (9 entries + pad: 80-byte records), d = 2 mu: -kappa U_mu(x);  d = 2 mu + 1: -kappa U_mu(x - mu)^dag  (periodic wrap, boundary sign in U)
struct GhostLinks { const float *u[4]; };   // decompressed U_mu at x_mu = X_mu - 1 of the backward neighbour, [parity][faceVh][18] (partitioned dims)
__global__ void pack_u_site_kernel(float2 *Us, const float *U, Geom g, float kappa, GhostLinks gl) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long Vh = g.Vh;
  if (t >= 2 * Vh * 8) return;
  const int d = (int)(t & 7);
  const long fs = t >> 3;
  const int parity = fs >= Vh ? 1 : 0;
  const int cb = (int)(fs - (long)parity * Vh);
  const int mu = d >> 1;
  float2 *dst = Us + ((size_t)fs * 8 + d) * 10;
  dst[9] = make_float2(0.f, 0.f);
  if (!(d & 1)) {
    const float *u = U + (((size_t)parity * 4 + mu) * Vh + cb) * 18;
    for (int e = 0; e < 9; e++) dst[e] = make_float2(-kappa * u[2 * e], -kappa * u[2 * e + 1]);
  } else {
    int x[4], full;
    cb_coords(x, full, cb, parity, g);
    const float *u;
    if (x[mu] == 0 && g.part[mu]) {   // the link lives on the backward neighbour: ghost copy, indexed by the face site
      u = gl.u[mu] + ((size_t)(1 - parity) * g.faceVh[mu] + face_index_rt(mu, x, g)) * 18;
    } else {
      x[mu] = (x[mu] + g.X[mu] - 1) % g.X[mu];
      const long ncb = ((((long)x[3] * g.X[2] + x[2]) * g.X[1] + x[1]) * g.X[0] + x[0]) >> 1;
      u = U + (((size_t)(1 - parity) * 4 + mu) * Vh + ncb) * 18;
    }
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) dst[r * 3 + c] = make_float2(-kappa * u[(c * 3 + r) * 2], kappa * u[(c * 3 + r) * 2 + 1]);
  }
}

// nbr[fs][d] = full index of x + e_d; across a partitioned boundary: index of the ghost record appended to Vs,
// Vf + ghost_off[mu][from forward (d even) : 1, from backward : 0] + neighbour parity * faceVh + face index
struct GhostOffsets { long off[4][2]; };
__global__ void fine_nbr_kernel(int *nbr, Geom g, GhostOffsets go) {
  const long fs = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long Vh = g.Vh;
  if (fs >= 2 * Vh) return;
  const int parity = fs >= Vh ? 1 : 0;
  const int cb = (int)(fs - (long)parity * Vh);
  int x[4], full;
  cb_coords(x, full, cb, parity, g);
  for (int d = 0; d < 8; d++) {
    const int mu = d >> 1;
    int y[4] = {x[0], x[1], x[2], x[3]};
    const bool edge = (d & 1) ? (x[mu] == 0) : (x[mu] == g.X[mu] - 1);
    if (edge && g.part[mu]) {
      nbr[fs * 8 + d] = (int)(2 * Vh + go.off[mu][(d & 1) ? 0 : 1] + (long)(1 - parity) * g.faceVh[mu] + face_index_rt(mu, x, g));
      continue;
    }
    y[mu] = (y[mu] + ((d & 1) ? g.X[mu] - 1 : 1)) % g.X[mu];
    const long ncb = ((((long)y[3] * g.X[2] + y[2]) * g.X[1] + y[1]) * g.X[0] + y[0]) >> 1;
    nbr[fs * 8 + d] = (int)((long)(1 - parity) * Vh + ncb);
  }
}

// per (aggregate, site-in-aggregate) record, in the order the CTAs walk: everything index-like the pipeline needs, so that
// no role chases pointers: {fs, nbr[8], mask (bit d: x + e_d leaves the aggregate), pad}
struct __align__(16) GmMeta { int fs; int nb[8]; int mask; int pad[2]; };
__global__ void galerkin_meta_kernel(GmMeta *meta, const int *c2f, const int *f2c, const int *nbr, long n, int bs, long Vf) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int X = (int)(t / bs);
  GmMeta m;
  m.fs = c2f[t];
  m.mask = 0;
  for (int d = 0; d < 8; d++) {
    m.nb[d] = nbr[(size_t)m.fs * 8 + d];
    if (m.nb[d] >= Vf || f2c[m.nb[d]] != X) m.mask |= 1 << d;   // ghost records (other rank) are never in this aggregate
  }
  m.pad[0] = m.pad[1] = 0;
  meta[t] = m;
}

struct GalerkinMmaArgs {
  const float2 *Vs;   // [Vf][12][nvec]
  const float2 *Us;   // [Vf][8][10]
  const GmMeta *meta; // [Vc][block_sites]
  const float *clover; // site-major packed clover term [Vf][72] or nullptr
  float *Y;           // [Vc][9][N][N/2] float4 viewed as floats: ((X*9 + d)*N + c)*2N + 2r + ri
  long Vc;
  int block_sites;
  float twist_a;
  int chunk;          // sites per accumulation chain
  int variant;        // tuning experiments (QB_GALERKIN_VARIANT)
  long long *dbg;     // timing trace of CTA 0 (QB_GALERKIN_TRACE): [item][8] clock64 stamps
};

constexpr int GM_GROUPS = 2;                       // builder groups of 128 threads (4 warps = the 4 TMEM lane quarters), alternate items
constexpr int GM_THREADS = 64 + GM_GROUPS * 128 + 128;  // producer, MMA, builders, epilogue
constexpr int GM_RAW_SLOTS = 8, GM_TILE_SLOTS = 6;

template <int NV> struct GmCfg {
  static constexpr int N = 2 * NV;                 // coarse components = accumulator columns
  static constexpr int M = 128;                    // MMA rows; 2N live
  static constexpr int V_BYTES = 12 * NV * 8;      // one site of Vs
  static constexpr int U_BYTES = 10 * 8;           // one link, padded to a multiple of 16 bytes (bulk-copy granularity)
  static constexpr int RAW_STRIDE = (2 * V_BYTES + U_BYTES + 127) & ~127;   // V(x), V(x + e_d), link of one work item
  static constexpr int B_HALF = N * 6 * 16;        // hi (or lo) part of the B tile: 6 chunks of 16 B per row
  static constexpr int TILE_BYTES = 2 * B_HALF;
  static constexpr int ZERO_BYTES = N * 2 * 16;    // one K step of zeros
  // tensor memory: two buffers of 4 accumulators (hop main (hi*hi), hop cross (lo*hi + hi*lo), diag main, diag cross), then
  // the A operand of each builder group: 24 columns hi + 24 columns lo (K = 12 complex = 24 tf32 per row)
  static constexpr int A_COL0 = 8 * N, A_COLS = 48;
  static constexpr int TMEM_NEED = A_COL0 + GM_GROUPS * A_COLS;
  static constexpr int TMEM_COLS = TMEM_NEED <= 128 ? 128 : (TMEM_NEED <= 256 ? 256 : 512);
  static_assert(TMEM_NEED <= 512, "accumulators do not fit tensor memory");
  static constexpr size_t smem_bytes() { return (size_t)GM_RAW_SLOTS * RAW_STRIDE + (size_t)GM_TILE_SLOTS * TILE_BYTES + ZERO_BYTES + 512 + 128; }
};

__device__ __forceinline__ void sts64(uint32_t saddr, float a, float b) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(saddr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ float2 lds64(uint32_t saddr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(saddr));
  return v;
}

// Work items are walked in the order (aggregate X, direction d, site i of X).  Per direction the epilogue keeps the running
// fp32 totals of the hopping block Y_d(X) and of the site-diagonal block in registers (one accumulator row per thread); the
// tensor core only ever accumulates `chunk` sites (3 * chunk MMAs per accumulator) before its partial sums are added to them.
// The A operand (V(x) side, 2N rows x 24) is written straight into tensor memory by the builder warps (tcgen05.st).
// Measured (clock64 trace, QB_GALERKIN_TRACE): the kernel is bound by the issue rate of the tcgen05 instructions: one
// elected thread gets one UTCHMMA / UTCBAR through every ~80 cycles whatever the tile size, so the nine 128 x N x 8
// MMAs + commits of a work item cost ~900 cycles although their math is 9 x 26 cycles (tensor pipe 15 % busy).  The next
// step is to batch the directions of a site into the N dimension (N = 240) so that each instruction carries 5x the work.
template <int NV>
__global__ void __launch_bounds__(GM_THREADS, 1) galerkin_mma_kernel(const GalerkinMmaArgs p) {
  using C = GmCfg<NV>;
  constexpr int N = C::N, M = C::M;
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  unsigned char *raw0 = smem;
  unsigned char *tile0 = raw0 + GM_RAW_SLOTS * C::RAW_STRIDE;
  unsigned char *zero0 = tile0 + GM_TILE_SLOTS * C::TILE_BYTES;
  uint64_t *bars = (uint64_t *)(zero0 + C::ZERO_BYTES);
  uint64_t *raw_full = bars, *raw_empty = raw_full + GM_RAW_SLOTS;
  uint64_t *tile_ready = raw_empty + GM_RAW_SLOTS, *tile_free = tile_ready + GM_TILE_SLOTS;
  uint64_t *acc_full = tile_free + GM_TILE_SLOTS, *acc_empty = acc_full + 2;
  uint64_t *a_free = acc_empty + 2;                   // [GM_GROUPS] the MMAs reading a group's A operand in tensor memory are done
  uint32_t *tmem_slot = (uint32_t *)(a_free + GM_GROUPS);
  int *tile_leaves = (int *)(tmem_slot + 1);          // [GM_TILE_SLOTS] 1: the item goes to the hopping block, 0: to the diagonal block
  int *raw_leaves = tile_leaves + GM_TILE_SLOTS;      // [GM_RAW_SLOTS] the same flag as published by the producer

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bs = p.block_sites;
  const int chunk = p.chunk;

  for (int e = tid; e < (GM_TILE_SLOTS * C::TILE_BYTES + C::ZERO_BYTES) / 16; e += GM_THREADS) ((float4 *)tile0)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  fence_proxy_async_smem();
  if (tid == 0) {
    for (int s = 0; s < GM_RAW_SLOTS; s++) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 128); }
    for (int s = 0; s < GM_TILE_SLOTS; s++) { mbar_init(&tile_ready[s], 128); mbar_init(&tile_free[s], 1); }
    for (int b = 0; b < 2; b++) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 128); }
    for (int g = 0; g < GM_GROUPS; g++) mbar_init(&a_free[g], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= producer =================
    // The whole warp takes part: lane l fetches the index record of item i0 + l (32 records in flight at once), then the
    // lanes issue their copies one after the other in item order.
    int slot = 0; uint32_t phase = 0;
    for (long X = blockIdx.x; X < p.Vc; X += gridDim.x) {
      const GmMeta *mrow = p.meta + (size_t)X * bs;
      for (int d = 0; d < 9; d++) {
        for (int i0 = 0; i0 < bs; i0 += 32) {
          const int cnt = bs - i0 < 32 ? bs - i0 : 32;
          int fs = 0, nb = 0, mask = 0;
          if (lane < cnt) { fs = __ldg(&mrow[i0 + lane].fs); nb = d < 8 ? __ldg(&mrow[i0 + lane].nb[d]) : 0; mask = __ldg(&mrow[i0 + lane].mask); }
          for (int l = 0; l < cnt; l++) {
            if (lane == l) {
              mbar_wait(&raw_empty[slot], phase ^ 1);
              unsigned char *st = raw0 + slot * C::RAW_STRIDE;
              raw_leaves[slot] = (d < 8 && ((mask >> d) & 1)) ? 1 : 0;  // published by the release of the arrive below
              if (d < 8) {
                mbar_arrive_expect_tx(&raw_full[slot], 2 * C::V_BYTES + C::U_BYTES);
                bulk_g2s(st, p.Vs + (size_t)fs * 12 * NV, C::V_BYTES, &raw_full[slot]);
                bulk_g2s(st + C::V_BYTES, p.Vs + (size_t)nb * 12 * NV, C::V_BYTES, &raw_full[slot]);
                bulk_g2s(st + 2 * C::V_BYTES, p.Us + ((size_t)fs * 8 + d) * 10, C::U_BYTES, &raw_full[slot]);
              } else if (p.clover) {
                mbar_arrive_expect_tx(&raw_full[slot], C::V_BYTES + 288);
                bulk_g2s(st, p.Vs + (size_t)fs * 12 * NV, C::V_BYTES, &raw_full[slot]);
                bulk_g2s(st + C::V_BYTES, p.clover + (size_t)fs * 72, 288, &raw_full[slot]);  // 72 floats: the site's two packed blocks
              } else {
                mbar_arrive_expect_tx(&raw_full[slot], C::V_BYTES);
                bulk_g2s(st, p.Vs + (size_t)fs * 12 * NV, C::V_BYTES, &raw_full[slot]);
              }
            }
            __syncwarp();
            if (++slot == GM_RAW_SLOTS) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issue (warp-uniform; one elected lane issues) =================
    constexpr uint32_t idesc = idesc_tf32(M, N);
    const uint64_t b_d0 = smem_desc(smem_u32(tile0), N * 16, 128);       // B hi of tile slot 0
    const uint64_t z_d = smem_desc(smem_u32(zero0), N * 16, 128);
    constexpr uint64_t b_step = (2 * N * 16) >> 4;
    int ts = 0; uint32_t tphase = 0;
    int buf = 0; uint32_t bphase[2] = {0, 0};
    int grp = 0;
    long cm = 0;
    for (long X = blockIdx.x; X < p.Vc; X += gridDim.x) {
      for (int d = 0; d < 9; d++) {
        for (int i = 0; i < bs; i++, cm++) {
          const bool mtrace = p.dbg && blockIdx.x == 0 && lane == 0 && cm < 512;
          if (mtrace) p.dbg[cm * 8 + 5] = clock64();
          mbar_wait(&tile_ready[ts], tphase);
          if (mtrace) p.dbg[cm * 8 + 6] = clock64();
          if (i % chunk == 0) {
            // new chain: wait until the epilogue has drained this buffer, then reset its accumulators with a zero product
            mbar_wait(&acc_empty[buf], bphase[buf] ^ 1);
            tc_fence_after_sync();
            if (elect_one()) {
#pragma unroll
              for (int a = 0; a < 4; a++) mma_tf32_ts(tmem_base + (buf * 4 + a) * N, tmem_base + C::A_COL0 + grp * C::A_COLS, z_d, idesc, 0u);
            }
            __syncwarp();
          }
          tc_fence_after_sync();
          const int leaves = tile_leaves[ts];
          const uint32_t d_main = tmem_base + (buf * 4 + (leaves ? 0 : 2)) * N, d_cross = d_main + N;
          const uint32_t ah = tmem_base + C::A_COL0 + grp * C::A_COLS, al = ah + 24;
          const uint64_t bh = b_d0 + (uint64_t)((ts * C::TILE_BYTES) >> 4), bl = bh + (uint64_t)(C::B_HALF >> 4);
          if (elect_one()) {
            if (p.variant == 5) {  // timing experiment: nine independent accumulators (wrong results)
#pragma unroll
              for (int k = 0; k < 3; k++) {
                mma_tf32_ts(tmem_base + ((3 * k) % 8) * N, ah + 8 * k, bh + k * b_step, idesc, 1u);
                mma_tf32_ts(tmem_base + ((3 * k + 1) % 8) * N, al + 8 * k, bh + k * b_step, idesc, 1u);
                mma_tf32_ts(tmem_base + ((3 * k + 2) % 8) * N, ah + 8 * k, bl + k * b_step, idesc, 1u);
              }
            } else if (p.variant == 6) {  // timing experiment: three MMAs only
#pragma unroll
              for (int k = 0; k < 3; k++) mma_tf32_ts(d_main, ah + 8 * k, bh + k * b_step, idesc, 1u);
            } else if (p.variant != 1) {
#pragma unroll
              for (int k = 0; k < 3; k++) {
                mma_tf32_ts(d_main, ah + 8 * k, bh + k * b_step, idesc, 1u);
                mma_tf32_ts(d_cross, al + 8 * k, bh + k * b_step, idesc, 1u);
                mma_tf32_ts(d_cross, ah + 8 * k, bl + k * b_step, idesc, 1u);
              }
            }
            mma_commit(&a_free[grp]);  // also frees the B slot (completion is in order; see the builders)
            if ((i + 1) % chunk == 0 || i + 1 == bs) mma_commit(&acc_full[buf]);
          }
          __syncwarp();
          if (mtrace) p.dbg[cm * 8 + 7] = clock64();
          if (++ts == GM_TILE_SLOTS) { ts = 0; tphase ^= 1; }
          if ((i + 1) % chunk == 0 || i + 1 == bs) { bphase[buf] ^= 1; buf ^= 1; }
          if (++grp == GM_GROUPS) grp = 0;
        }
      }
    }
  } else if (warp < 2 + 4 * GM_GROUPS) {
    // ================= builders: GM_GROUPS groups of 128 threads take the work items round-robin =================
    const int grp = (warp - 2) >> 2, bt = tid - 64 - grp * 128;
    const int q = warp & 3;  // TMEM lane quarter of this warp
    const long nitem = ((p.Vc - blockIdx.x + gridDim.x - 1) / gridDim.x) * 9 * bs;
    int slot = grp % GM_RAW_SLOTS, ts = grp % GM_TILE_SLOTS;
    uint32_t sphase = 0, tphase = 0, aphase = 0;
    int i = grp % bs, d = (grp / bs) % 9;  // item c = (aggregate * 9 + d) * bs + i, advanced incrementally
    // A-operand row of this thread: m = 32 q + lane = 2 r + ri, r = (S', j')
    const int m = q * 32 + lane;
    const int ar = m >> 1, ari = m & 1, aS = ar / NV, aj = ar - aS * NV;
    const bool a_live = m < 2 * N;
    const uint32_t a_taddr = tmem_base + C::A_COL0 + grp * C::A_COLS + ((uint32_t)(q * 32) << 16);
    for (long c = grp; c < nitem; c += GM_GROUPS) {
      const bool trace = p.dbg && blockIdx.x == 0 && bt == 0 && c < 512;
      if (trace) p.dbg[c * 8 + 0] = clock64();
      mbar_wait(&raw_full[slot], sphase);
      if (trace) p.dbg[c * 8 + 1] = clock64();
      // (the B slot is free: MMAs complete in order, and this group already waited for the item two slots of its own back)
      const uint32_t st = smem_u32(raw0 + slot * C::RAW_STRIDE);
      const uint32_t bh = smem_u32(tile0 + ts * C::TILE_BYTES), bl = bh + C::B_HALF;
      if (bt == 0) tile_leaves[ts] = raw_leaves[slot];
      // ---- B tile: role (n = (S, j), si) computes T[c'] = sum_c U[c'][c] V(x + e_d)[(ss, c)][j] for the spin ss = 2 S + si
      for (int role = bt; role < 2 * N; role += 128) {
        const int n = role % N, si = role / N;
        const int S = n / NV, j = n - S * NV;
        const int ss = 2 * S + si;
        cplx<float> T[3];
        int sp_opp;                   // the row spin s' of the other chirality that couples to ss through gamma_mu
        cplx<float> coef(0.f, 0.f);
        if (d < 8) {
          const int mu = d >> 1;
          const uint32_t vn = st + C::V_BYTES, us = st + 2 * C::V_BYTES;
          // all twelve shared-memory loads first (they are volatile asm: issued in this order, one wait for the lot)
          float2 qv[3], qu[9];
#pragma unroll
          for (int cc = 0; cc < 3; cc++) qv[cc] = lds64(vn + (uint32_t)((ss * 3 + cc) * NV + j) * 8);
#pragma unroll
          for (int e = 0; e < 9; e++) qu[e] = lds64(us + (uint32_t)e * 8);
#pragma unroll
          for (int cp = 0; cp < 3; cp++) {
            T[cp] = cplx<float>(0.f, 0.f);
#pragma unroll
            for (int cc = 0; cc < 3; cc++) cmac(T[cp], cplx<float>(qu[cp * 3 + cc].x, qu[cp * 3 + cc].y), cplx<float>(qv[cc].x, qv[cc].y));
          }
          sp_opp = mu < 2 ? 3 - ss : (ss + 2) & 3;
          const float sigma = (d & 1) ? 1.f : -1.f;  // forward: 1 - gamma_mu, backward: 1 + gamma_mu
          const int gr = mu == 1 ? ((sp_opp == 0 || sp_opp == 3) ? -1 : 1) : (mu == 3 ? 1 : 0);
          const int gi = mu == 0 ? (sp_opp < 2 ? 1 : -1) : (mu == 2 ? ((sp_opp == 0 || sp_opp == 3) ? 1 : -1) : 0);
          coef = cplx<float>(sigma * gr, sigma * gi);
        } else {
          // site-local term (1 + i a gamma5) V(x), or (C + i a gamma5) V(x) with a clover term: chirality-diagonal, the other
          // chirality's rows are zero
          sp_opp = (ss + 2) & 3;
          const float tw = S == 0 ? p.twist_a : -p.twist_a;
          if (!p.clover) {
            float2 qv[3];
#pragma unroll
            for (int cp = 0; cp < 3; cp++) qv[cp] = lds64(st + (uint32_t)((ss * 3 + cp) * NV + j) * 8);
#pragma unroll
            for (int cp = 0; cp < 3; cp++) T[cp] = cplx<float>(qv[cp].x - tw * qv[cp].y, qv[cp].y + tw * qv[cp].x);
          } else {
            float2 qv[6];
#pragma unroll
            for (int c2 = 0; c2 < 6; c2++) qv[c2] = lds64(st + (uint32_t)((6 * S + c2) * NV + j) * 8);
            const uint32_t cbk = st + C::V_BYTES + (uint32_t)S * 144;  // packed Hermitian block of chirality S: 6 reals + 15 complex
#pragma unroll
            for (int cp = 0; cp < 3; cp++) {
              const int r = si * 3 + cp;
              float dg;
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(dg) : "r"(cbk + (uint32_t)r * 4));
              const float2 vr = si ? qv[3 + cp] : qv[cp];  // = qv[r] without dynamic register indexing
              cplx<float> w(dg * vr.x - tw * vr.y, dg * vr.y + tw * vr.x);
#pragma unroll
              for (int c2 = 0; c2 < 6; c2++) {
                if (c2 == r) continue;
                const int lo = c2 < r ? c2 : r, hi = c2 < r ? r : c2;
                const int kk = 15 - (6 - lo) * (5 - lo) / 2 + hi - lo - 1;
                const float2 l = lds64(cbk + 24 + (uint32_t)kk * 8);
                if (r > c2) cmac(w, cplx<float>(l.x, l.y), cplx<float>(qv[c2].x, qv[c2].y));
                else cmac_conj(w, cplx<float>(l.x, l.y), cplx<float>(qv[c2].x, qv[c2].y));
              }
              T[cp] = w;
            }
          }
        }
#pragma unroll
        for (int cp = 0; cp < 3; cp++) {
          const cplx<float> w0 = T[cp], w1 = coef * T[cp];
          const int k0 = ss * 3 + cp, k1 = sp_opp * 3 + cp;
          const uint32_t o0 = (uint32_t)(n + (k0 >> 1) * N) * 16 + (k0 & 1) * 8, o1 = (uint32_t)(n + (k1 >> 1) * N) * 16 + (k1 & 1) * 8;
          const float h0r = tf32_hi(w0.re), h0i = tf32_hi(w0.im), h1r = tf32_hi(w1.re), h1i = tf32_hi(w1.im);
          sts64(bh + o0, h0r, h0i); sts64(bl + o0, w0.re - h0r, w0.im - h0i);
          sts64(bh + o1, h1r, h1i); sts64(bl + o1, w1.re - h1r, w1.im - h1i);
        }
      }
      // the A operand last: its slot in tensor memory is free only when the MMAs of this group's previous item are done,
      // and that latency is now hidden behind the B tile above
      if (trace) p.dbg[c * 8 + 2] = clock64();
      mbar_wait(&a_free[grp], aphase ^ 1);
      if (trace) p.dbg[c * 8 + 3] = clock64();
      tc_fence_after_sync();
      // ---- A operand row from V(x): columns (k, re/im), k = 0 .. 11; only chirality(k) = S' is non-zero
      {
        float hi[24], lo[24];
#pragma unroll
        for (int e = 0; e < 24; e++) { hi[e] = 0.f; lo[e] = 0.f; }
        if (a_live) {
          float2 av[6];
#pragma unroll
          for (int kk = 0; kk < 6; kk++) av[kk] = lds64(st + (uint32_t)((6 * aS + kk) * NV + aj) * 8);
#pragma unroll
          for (int kk = 0; kk < 6; kk++) {
            const float2 v = av[kk];
            const float x0 = ari ? -v.y : v.x, x1 = ari ? v.x : v.y;
            const float h0 = tf32_hi(x0), h1 = tf32_hi(x1);
#pragma unroll
            for (int S = 0; S < 2; S++)
              if (S == aS) { hi[12 * S + 2 * kk] = h0; hi[12 * S + 2 * kk + 1] = h1; lo[12 * S + 2 * kk] = x0 - h0; lo[12 * S + 2 * kk + 1] = x1 - h1; }
          }
        }
#pragma unroll
        for (int k8 = 0; k8 < 3; k8++) { tmem_st8(a_taddr + 8 * k8, hi + 8 * k8); tmem_st8(a_taddr + 24 + 8 * k8, lo + 8 * k8); }
      }
      tmem_st_wait();
      tc_fence_before_sync();
      fence_proxy_async_smem();
      if (trace) p.dbg[c * 8 + 4] = clock64();
      mbar_arrive(&tile_ready[ts]);
      mbar_arrive(&raw_empty[slot]);
      slot += GM_GROUPS; if (slot >= GM_RAW_SLOTS) { slot -= GM_RAW_SLOTS; sphase ^= 1; }
      ts += GM_GROUPS; if (ts >= GM_TILE_SLOTS) { ts -= GM_TILE_SLOTS; tphase ^= 1; }
      aphase ^= 1;
      i += GM_GROUPS;
      while (i >= bs) { i -= bs; if (++d == 9) d = 0; }
    }
  } else {
    // ================= epilogue: chunk accumulators -> running fp32 totals (registers) -> coarse links =================
    const int q = warp & 3;
    const int m = q * 32 + lane;  // accumulator row = 2 r + ri
    int buf = 0; uint32_t bphase[2] = {0, 0};
    for (long X = blockIdx.x; X < p.Vc; X += gridDim.x) {
      float diag[N];
#pragma unroll
      for (int c = 0; c < N; c++) diag[c] = 0.f;
      for (int d = 0; d < 9; d++) {
        float hop[N];
#pragma unroll
        for (int c = 0; c < N; c++) hop[c] = 0.f;
        for (int i0 = 0; i0 < bs; i0 += chunk) {
          mbar_wait(&acc_full[buf], bphase[buf]);
          bphase[buf] ^= 1;
          tc_fence_after_sync();
          const uint32_t tb = tmem_base + (buf * 4) * N + ((uint32_t)(q * 32) << 16);
          float v[N];
#pragma unroll
          for (int a = 0; a < 4; a++) {
#pragma unroll
            for (int c = 0; c < N / 16; c++) tmem_ld16(tb + a * N + c * 16, v + c * 16);
            tmem_ld_wait();
            if (a < 2) {
#pragma unroll
              for (int c = 0; c < N; c++) hop[c] += v[c];
            } else {
#pragma unroll
              for (int c = 0; c < N; c++) diag[c] += v[c];
            }
          }
          tc_fence_before_sync();
          mbar_arrive(&acc_empty[buf]);
          buf ^= 1;
        }
        if (d < 8 && m < 2 * N) {
          float *y = p.Y + ((size_t)X * 9 + d) * N * 2 * N + m;
#pragma unroll
          for (int c = 0; c < N; c++) y[(size_t)c * 2 * N] = hop[c];
        }
      }
      if (m < 2 * N) {
        float *y = p.Y + ((size_t)X * 9 + 8) * N * 2 * N + m;
#pragma unroll
        for (int c = 0; c < N; c++) y[(size_t)c * 2 * N] = diag[c];
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

template <int NV> static void launch_gm(const GalerkinMmaArgs &a) {
  using C = GmCfg<NV>;
  int dev = 0, nsm = 0;
  QB_CUDA(cudaGetDevice(&dev));
  QB_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  const size_t sm = C::smem_bytes();
  QB_CUDA(cudaFuncSetAttribute(galerkin_mma_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  const unsigned grid = (unsigned)(a.Vc < nsm ? a.Vc : nsm);
  galerkin_mma_kernel<NV><<<grid, GM_THREADS, sm, rt().compute>>>(a);
  QB_CHECK_LAUNCH();
}

bool galerkin_mma_supported(const Transfer &T) {
  if (getenv("QB_GALERKIN_MMA") && atoi(getenv("QB_GALERKIN_MMA")) == 0) return false;
  if (T.Nf != 12) return false;
  if (T.fine.partitioned() && getenv("QB_GALERKIN_MMA_PARTITIONED") && atoi(getenv("QB_GALERKIN_MMA_PARTITIONED")) == 0) return false;
  return T.nvec == 8 || T.nvec == 16 || T.nvec == 24;
}

// out.Y must be allocated and zero (CoarseOperator::allocate)
void build_coarse_from_fine_mma(CoarseOperator &out, const Transfer &T, const GaugeField &gauge, const Geom &fine_geom, double kappa, double twist_a,
                                const float *clover_site) {
  cudaStream_t s = rt().compute;
  const long Vh = T.fine.Vh, Vf = 2 * Vh;
  float *U = decompress_gauge(gauge, fine_geom);
  float2 *Vs, *Us;
  int *nbr;
  // partitioned dimensions: the neighbours' boundary slices of V (Transfer::Vghost) become extra site-major records behind the Vf local
  // ones and the neighbour table points at them; backward links across the boundary come from the gauge field's ghost links.
  // The tensor-core kernel itself only follows the tables.
  GhostOffsets go{};
  GhostLinks gl{};
  float *ug[4] = {nullptr, nullptr, nullptr, nullptr};
  long nghost = 0;
  for (int d = 0; d < 4; d++) {
    gl.u[d] = nullptr;
    for (int k = 0; k < 2; k++) go.off[d][k] = 0;
    if (!fine_geom.part[d]) continue;
    if (!T.Vghost[d][0] || !T.Vghost[d][1]) QB_ERROR("transfer operator has no ghost V for partitioned dimension %d", d);
    if (!gauge.ghost[d]) QB_ERROR("gauge field has no ghost links for partitioned dimension %d", d);
    for (int k = 0; k < 2; k++) { go.off[d][k] = nghost; nghost += 2L * fine_geom.faceVh[d]; }
    ug[d] = decompress_ghost_links(gauge, fine_geom, d);
    gl.u[d] = ug[d];
  }
  Vs = (float2 *)pool_malloc((size_t)(Vf + nghost) * 12 * T.nvec * sizeof(float2));
  Us = (float2 *)pool_malloc((size_t)Vf * 80 * sizeof(float2));
  nbr = (int *)pool_malloc((size_t)Vf * 8 * sizeof(int));
  {
    const long n = Vf * 12 * (T.nvec / 2);
    pack_v_site_major_kernel<<<div_up(n, 256), 256, 0, s>>>(Vs, (const float4 *)T.V, 12, T.nvec, Vh);
    QB_CHECK_LAUNCH();
    for (int d = 0; d < 4; d++)
      for (int k = 0; k < 2 && fine_geom.part[d]; k++) {
        const long fv = fine_geom.faceVh[d], ng = 2 * fv * 12 * (T.nvec / 2);
        pack_v_site_major_kernel<<<div_up(ng, 256), 256, 0, s>>>(Vs + (size_t)(Vf + go.off[d][k]) * 12 * T.nvec, (const float4 *)T.Vghost[d][k], 12, T.nvec, fv);
        QB_CHECK_LAUNCH();
      }
    pack_u_site_kernel<<<div_up(Vf * 8, 256), 256, 0, s>>>(Us, U, fine_geom, (float)kappa, gl);
    QB_CHECK_LAUNCH();
    fine_nbr_kernel<<<div_up(Vf, 256), 256, 0, s>>>(nbr, fine_geom, go);
    QB_CHECK_LAUNCH();
  }
  GmMeta *meta;
  const long nmeta = T.coarse.V() * T.block_sites;
  meta = (GmMeta *)pool_malloc((size_t)nmeta * sizeof(GmMeta));
  galerkin_meta_kernel<<<div_up(nmeta, 256), 256, 0, s>>>(meta, T.c2f, T.f2c, nbr, nmeta, T.block_sites, Vf);
  QB_CHECK_LAUNCH();
  GalerkinMmaArgs a{};
  a.Vs = Vs; a.Us = Us; a.meta = meta; a.clover = clover_site; a.Y = out.Y; a.Vc = T.coarse.V(); a.block_sites = T.block_sites;
  a.twist_a = (float)twist_a;
  a.chunk = getenv("QB_GALERKIN_CHUNK") ? atoi(getenv("QB_GALERKIN_CHUNK")) : 8;
  if (a.chunk < 1) a.chunk = 1;
  a.variant = getenv("QB_GALERKIN_VARIANT") ? atoi(getenv("QB_GALERKIN_VARIANT")) : 0;
  a.dbg = nullptr;
  if (getenv("QB_GALERKIN_TRACE")) {
    QB_CUDA(cudaMalloc((void **)&a.dbg, 512 * 8 * sizeof(long long)));
    QB_CUDA(cudaMemsetAsync(a.dbg, 0, 512 * 8 * sizeof(long long), s));
  }
  cudaEvent_t e0, e1;
  QB_CUDA(cudaEventCreate(&e0)); QB_CUDA(cudaEventCreate(&e1));
  QB_CUDA(cudaEventRecord(e0, s));
  switch (T.nvec) {
    case 8: launch_gm<8>(a); break;
    case 16: launch_gm<16>(a); break;
    case 24: launch_gm<24>(a); break;
    default: QB_ERROR("tensor-core coarse-link build: n_vec = %d is not instantiated", T.nvec);
  }
  QB_CUDA(cudaEventRecord(e1, s));
  QB_CUDA(cudaStreamSynchronize(s));
  float ms = 0;
  QB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  const double flops = (double)Vf * 9 * 8.0 * 6 * (2.0 * T.nvec) * (2.0 * T.nvec);  // V^dag W: rank-6 update of an N x N complex block per (site, direction)
  log_msg(1, "coarse-link build on the tensor cores: %.3f ms (chains of %d sites), %.1f TFLOP/s useful\n", ms, a.chunk, flops / ms / 1e9);
  if (a.dbg) {
    std::vector<long long> h(512 * 8);
    QB_CUDA(cudaMemcpy(h.data(), a.dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    FILE *f = fopen(getenv("QB_GALERKIN_TRACE"), "w");
    if (f) {
      for (int c = 0; c < 512; c++) {
        for (int e = 0; e < 8; e++) fprintf(f, "%lld ", h[c * 8 + e] - h[0]);
        fprintf(f, "\n");
      }
      fclose(f);
    }
    cudaFree(a.dbg);
  }
  pool_free(U); pool_free(Vs); pool_free(Us); pool_free(nbr); pool_free(meta);
  for (int d = 0; d < 4; d++)
    if (ug[d]) QB_CUDA(cudaFree(ug[d]));
}

}  // namespace qb
