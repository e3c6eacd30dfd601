// Multi-right-hand-side coarse Dslash on the 5th-generation tensor cores (tcgen05 / TMEM).
//
//     out_r(x) = [X or Xinv] in_r(x) + sum_{d<8} L_d(x) in_r(x + e_d),      r = 0 .. R-1 right-hand sides at once
//
// (reference: one right-hand side at a time, lib/dslash_coarse.cu:49-333; the multi-RHS request is BASELINE config 5).
// With R right-hand sides the N x N complex links of a site (9 N^2 x 8 B = 166 KB for N = 48) are read ONCE for
// R matrix-vector products: the arithmetic intensity grows from 1 flop/B to ~R flop/B and the fp32 FMA pipe
// (74 TFLOP/s) would become the limit at R ~ 12; the tensor pipe (tf32: 1.1 PFLOP/s) keeps the kernel on the HBM
// roofline up to R ~ 64.
//
// Mapping to the MMA  D[M x Nmma] += A[M x K] * B[Nmma x K]^T  (both operands K-major in shared memory):
//   * B = the link matrix exactly as stored in HBM ("Ymma" layout below): row n = output colour i,
//     K index = (input colour c, re/im) -> K = 2N per direction, 9 directions chained into one accumulation.
//   * A = the right-hand sides, two rows per vector:  row 2r = (Re b, -Im b) -> Re(out),  row 2r+1 = (Im b, Re b) -> Im(out).
//     Built in shared memory by the builder warps from the neighbour site's block of vectors.
//   * D lives in tensor memory: lane = (r, re/im), column = output colour; double buffered so that the epilogue
//     of one site overlaps the MMAs of the next.
//   * tf32 has 11 significant bits; MODE 3 recovers fp32 accuracy by splitting both operands x = hi + lo
//     (hi = the bits the tensor core reads, lo = x - hi exactly): the A tile carries hi rows and lo rows, the
//     B tile is used twice (hi, lo), and the epilogue adds the hi-row and lo-row results = all four partial products.
//
// One persistent CTA per SM walks sites; warp roles:
//   warp 0    : producer  - one 1-D bulk copy (cp.async.bulk, mbarrier complete_tx) per (site, direction) link matrix
//   warp 1    : MMA issue - one thread, tcgen05.mma.kind::tf32, tcgen05.commit releases stages / publishes accumulators
//   warps 2-9  : builders  - A tile (and in MODE 3 the hi / lo split of the B tile): two groups of 128 threads on alternate stages
//   warps 10-13: epilogue - tcgen05.ld -> shared staging -> a * out + b * xpay -> coalesced 128-bit stores
#include <cstdlib>
#include "coarse.h"
#include "comm.h"
#include "tc05.cuh"

namespace qb {

using namespace tc;

// ---- layouts ------------------------------------------------------------------------------------------------
// Ymma[site][d][kc][i] float4 = (L_d[i][2kc].re, .im, L_d[i][2kc+1].re, .im): the canonical no-swizzle K-major
// UMMA operand (8-row x 16-byte core matrices, SBO = 128 B, LBO = 16 N B) stored verbatim in HBM, so that one
// contiguous 8 N^2-byte bulk copy per (site, direction) lands ready for the tensor core.
__global__ void ymma_from_y_kernel(float4 *dst, const float4 *src, int N, long nmat) {
  const int NRP = N / 2;
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nmat * N * NRP) return;
  const long mat = t / (N * NRP);
  const int e = (int)(t - mat * N * NRP);
  const int kc = e / N, i = e - kc * N;
  // source layout: [col c][row pair rp] float4 = (L[2rp][c], L[2rp+1][c])
  const float4 *m = src + mat * N * NRP;
  const float4 c0 = m[(size_t)(2 * kc) * NRP + (i >> 1)], c1 = m[(size_t)(2 * kc + 1) * NRP + (i >> 1)];
  dst[t] = (i & 1) ? make_float4(c0.z, c0.w, c1.z, c1.w) : make_float4(c0.x, c0.y, c1.x, c1.y);
}

struct GhostOff { long off[4][2]; };

// neighbour table: full-site index (parity * Vh + cb) of x + e_d for d < 8; hops across a partitioned boundary point at ghost sites,
// numbered from 2 Vh (CoarseOperator::mrhs_goff)
__global__ void coarse_nbr_kernel(int *nbr, LevelGeom g, GhostOff go) {
  const long fs = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (fs >= 2L * g.Vh) return;
  const int parity = fs >= g.Vh ? 1 : 0;
  const long cb = fs - (long)parity * g.Vh;
  const long za = cb / g.Xh, zb = za / g.X[1];
  int x[4];
  x[1] = (int)(za - zb * g.X[1]);
  x[3] = (int)(zb / g.X[2]);
  x[2] = (int)(zb - (long)x[3] * g.X[2]);
  x[0] = (int)(2 * cb + ((x[1] + x[2] + x[3] + parity) & 1) - za * g.X[0]);
  for (int d = 0; d < 8; d++) {
    const int mu = d >> 1;
    int y[4] = {x[0], x[1], x[2], x[3]};
    const bool wrap = (d & 1) ? x[mu] == 0 : x[mu] == g.X[mu] - 1;
    if (wrap && g.part[mu]) {
      // face index of the neighbour on its slice: 3-d lexicographic index of the other coordinates >> 1 (coarse_face_to_cb)
      const int d0 = mu == 0 ? 1 : 0, d1 = mu <= 1 ? 2 : 1, d2 = mu <= 2 ? 3 : 2;
      const long fidx = (((long)x[d2] * g.X[d1] + x[d1]) * g.X[d0] + x[d0]) >> 1;
      nbr[fs * 8 + d] = (int)(2L * g.Vh + go.off[mu][(d & 1) ? 0 : 1] + (long)(1 - parity) * g.faceVh[mu] + fidx);
      continue;
    }
    y[mu] = (y[mu] + ((d & 1) ? g.X[mu] - 1 : 1)) % g.X[mu];
    const long ncb = ((((long)y[3] * g.X[2] + y[2]) * g.X[1] + y[1]) * g.X[0] + y[0]) >> 1;
    nbr[fs * 8 + d] = (int)((long)(1 - parity) * g.Vh + ncb);
  }
}

void CoarseOperator::prepare_mrhs() {
  cudaStream_t s = rt().compute;
  const long V = geom.V();
  if (mrhs_ready && Ymma && nbr && (!Xinv || Xinv_mma)) return;   // links are immutable once the level is set up
  mrhs_ready = true;
  if (!Ymma) Ymma = (float *)pool_malloc(link_bytes());
  {
    const long n = V * 9 * N * (N / 2);
    ymma_from_y_kernel<<<div_up(n, 256), 256, 0, s>>>((float4 *)Ymma, (const float4 *)Y, N, V * 9);
    QB_CHECK_LAUNCH();
  }
  if (Xinv) {
    if (!Xinv_mma) Xinv_mma = (float *)pool_malloc((size_t)V * N * N * 8);
    const long n = V * N * (N / 2);
    ymma_from_y_kernel<<<div_up(n, 256), 256, 0, s>>>((float4 *)Xinv_mma, (const float4 *)Xinv, N, V);
    QB_CHECK_LAUNCH();
  }
  if (!nbr) {
    GhostOff go;
    long off = 0;
    for (int d = 0; d < 4; d++)
      for (int dir = 0; dir < 2; dir++) {
        go.off[d][dir] = mrhs_goff[d][dir] = off;
        if (geom.part[d]) off += 2L * geom.faceVh[d];
      }
    mrhs_ghost_sites = off;
    QB_CUDA(cudaMalloc((void **)&nbr, sizeof(int) * 8 * V));
    coarse_nbr_kernel<<<div_up(V, 256), 256, 0, s>>>(nbr, geom, go);
    QB_CHECK_LAUNCH();
  }
}

// ---- ghost zone of block fields ------------------------------------------------------------------------------------------------
struct BlockPackArgs {
  float4 *dst[4][2];   // where face (d, dir) goes: a block of the send arena, or of the neighbour's ghost zone itself (peer delivery)
  const float4 *field;
  long poff[2];
  int X[4], faceVh[4], part[4];
  long off[5];   // prefix sums of the float4 counts per partitioned dimension
  int nelem, parity_mask;
};
__device__ __forceinline__ long block_face_to_cb(int mu, int fidx, int slice, int parity, const int *X) {
  const int d0 = mu == 0 ? 1 : 0, d1 = mu <= 1 ? 2 : 1, d2 = mu <= 2 ? 3 : 2;
  const int L0 = X[d0], L1 = X[d1];
  const int f2 = 2 * fidx;
  const int row = f2 / L0;
  const int c = row / L1, b = row - c * L1;
  int a = f2 - row * L0;
  a += (slice + b + c + parity + a) & 1;
  int x[4];
  x[mu] = slice; x[d0] = a; x[d1] = b; x[d2] = c;
  return ((((long)x[3] * X[2] + x[2]) * X[1] + x[1]) * X[0] + x[0]) >> 1;
}
// send arena, same site numbering as the receiver's ghost zone seen from the other side: block [d][0] = my slice x_d = 0 (travels
// backward, becomes the neighbour's [d][1]), block [d][1] = my slice x_d = X_d - 1
__global__ void block_ghost_pack_kernel(const BlockPackArgs a) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.off[4]) return;
  int d = 0;
  while (d < 3 && t >= a.off[d + 1]) d++;
  long r = t - a.off[d];
  const int e = (int)(r % a.nelem); r /= a.nelem;
  const int fv = a.faceVh[d];
  const int fidx = (int)(r % fv); r /= fv;
  const int parity = (int)(r & 1), dir = (int)(r >> 1);
  if (!((a.parity_mask >> parity) & 1)) return;
  const long cb = block_face_to_cb(d, fidx, dir ? a.X[d] - 1 : 0, parity, a.X);
  a.dst[d][dir][(size_t)((long)parity * fv + fidx) * a.nelem + e] = a.field[a.poff[parity] + (size_t)cb * a.nelem + e];
}

void CoarseOperator::exchange_block_ghost(const float *field, const long *poff, int parity_mask, int R) const {
  if (!geom.partitioned()) return;
  Runtime &r = rt();
  const int nelem = (N / 2) * R;
  const size_t need = (size_t)mrhs_ghost_sites * nelem * sizeof(float4);
  if (need > mrhs_arena_bytes) {
    // (collective: `need` depends on the operator and R only, alike on all ranks)
    QB_CUDA(cudaStreamSynchronize(r.compute));
    if (mrhs_send) cudaFree(mrhs_send);
    QB_CUDA(cudaMalloc((void **)&mrhs_send, need));
    mrhs_arena.create(need);
    mrhs_arena_bytes = need;
  }
  const bool peer = mrhs_arena.peer;
  if (peer) mrhs_arena.seq++;
  mrhs_recv = (float *)mrhs_arena.recv_base();
  HaloFlags sig{}, wt{};
  sig.seq = wt.seq = mrhs_arena.seq;
  const size_t site_bytes = (size_t)nelem * sizeof(float4);
  BlockPackArgs a;
  a.field = (const float4 *)field;
  a.poff[0] = poff[0]; a.poff[1] = poff[1];
  a.nelem = nelem; a.parity_mask = parity_mask;
  long off = 0;
  for (int d = 0; d < 4; d++) {
    a.X[d] = geom.X[d]; a.faceVh[d] = geom.faceVh[d]; a.part[d] = geom.part[d];
    for (int dir = 0; dir < 2; dir++) {
      a.dst[d][dir] = (float4 *)((char *)mrhs_send + (size_t)mrhs_goff[d][dir] * site_bytes);
      if (peer && geom.part[d]) {
        // my slice 0 (dir 0) -> the backward neighbour's "from forward" block [d][1]; my last slice -> the forward neighbour's [d][0]
        const int nb = comm_neighbor_rank(d, dir);
        a.dst[d][dir] = (float4 *)(mrhs_arena.send_base(nb) + (size_t)mrhs_goff[d][1 - dir] * site_bytes);
        sig.p[sig.n++] = mrhs_arena.flag_of(nb, d * 2 + (1 - dir));
        wt.p[wt.n++] = mrhs_arena.flag_mine(d * 2 + dir);
      }
    }
    a.off[d] = off;
    if (geom.part[d]) off += 4L * geom.faceVh[d] * nelem;
  }
  a.off[4] = off;
  block_ghost_pack_kernel<<<div_up(off, 256), 256, 0, r.compute>>>(a);
  QB_CHECK_LAUNCH();
  if (peer) {
    comm_halo_signal(sig, r.compute);
    comm_halo_wait(wt, r.compute);
    return;
  }
  const void *sb[8]; void *rb[8]; int to[8], from[8]; size_t nb[8];
  int n = 0;
  for (int d = 0; d < 4; d++) {
    if (!geom.part[d]) continue;
    const size_t blk = (size_t)2 * geom.faceVh[d] * nelem * sizeof(float4);
    char *s0 = (char *)mrhs_send + (size_t)mrhs_goff[d][0] * nelem * sizeof(float4), *s1 = (char *)mrhs_send + (size_t)mrhs_goff[d][1] * nelem * sizeof(float4);
    char *r0 = (char *)mrhs_recv + (size_t)mrhs_goff[d][0] * nelem * sizeof(float4), *r1 = (char *)mrhs_recv + (size_t)mrhs_goff[d][1] * nelem * sizeof(float4);
    // my slice 0 -> the backward neighbour's "from forward" block [d][1]; my last slice -> the forward neighbour's "from backward" block [d][0]
    sb[n] = s0; to[n] = comm_neighbor_rank(d, 0); rb[n] = r1; from[n] = comm_neighbor_rank(d, 1); nb[n++] = blk;
    sb[n] = s1; to[n] = comm_neighbor_rank(d, 1); rb[n] = r0; from[n] = comm_neighbor_rank(d, 0); nb[n++] = blk;
  }
  comm_sendrecv_group(n, sb, to, rb, from, nb, r.compute);
}

// ---- block fields: [parity][cb][kc][r] float4 <-> R single fields [parity][plane kc][cb] float4 ----------------------
__global__ void block_pack_kernel(float4 *blk, const float4 *const *fields, int R, int NKC, long Vh, int nparity, int to_block) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long n = (long)nparity * Vh * NKC * R;
  if (t >= n) return;
  const int r = (int)(t % R);
  long u = t / R;
  const int kc = (int)(u % NKC);
  u /= NKC;  // parity * Vh + cb
  const int p = (int)(u / Vh);
  const long cb = u - (long)p * Vh;
  float4 *f = (float4 *)fields[r] + ((size_t)p * NKC + kc) * Vh + cb;
  if (to_block) blk[t] = *f;
  else *f = blk[t];
}

CoarseBlockField::CoarseBlockField(long Vh_, int nparity_, int N_, int R_) : Vh(Vh_), nparity(nparity_), N(N_), R(R_) {
  v = (float *)pool_malloc(bytes());
  QB_CUDA(cudaMalloc((void **)&ptrs, sizeof(void *) * R));
}
CoarseBlockField::CoarseBlockField(long Vh_, int nparity_, int N_, int R_, float *storage) : Vh(Vh_), nparity(nparity_), N(N_), R(R_), v(storage), owner(false) {
  QB_CUDA(cudaMalloc((void **)&ptrs, sizeof(void *) * R));
}
CoarseBlockField::~CoarseBlockField() {
  if (v && owner) pool_free(v);
  if (ptrs) cudaFree(ptrs);
}
static void block_copy(CoarseBlockField &b, SpinorField *const *f, int to_block) {
  std::vector<const void *> h(b.R);
  for (int r = 0; r < b.R; r++) {
    if (f[r]->prec != PREC_SINGLE || f[r]->ncomplex != b.N || f[r]->Vh != b.Vh || f[r]->nparity != b.nparity) QB_ERROR("block field: member %d does not match", r);
    h[r] = f[r]->v;
  }
  cudaStream_t s = rt().compute;
  QB_CUDA(cudaMemcpyAsync(b.ptrs, h.data(), sizeof(void *) * b.R, cudaMemcpyHostToDevice, s));
  QB_CUDA(cudaStreamSynchronize(s));  // h goes out of scope
  const long n = (long)b.nparity * b.Vh * (b.N / 2) * b.R;
  block_pack_kernel<<<div_up(n, 256), 256, 0, s>>>((float4 *)b.v, (const float4 *const *)b.ptrs, b.R, b.N / 2, b.Vh, b.nparity, to_block);
  QB_CHECK_LAUNCH();
}
static void block_copy_ptrs(CoarseBlockField &b, const void *const *src, int to_block) {
  cudaStream_t s = rt().compute;
  QB_CUDA(cudaMemcpyAsync(b.ptrs, src, sizeof(void *) * b.R, cudaMemcpyHostToDevice, s));
  QB_CUDA(cudaStreamSynchronize(s));
  const long n = (long)b.nparity * b.Vh * (b.N / 2) * b.R;
  block_pack_kernel<<<div_up(n, 256), 256, 0, s>>>((float4 *)b.v, (const float4 *const *)b.ptrs, b.R, b.N / 2, b.Vh, b.nparity, to_block);
  QB_CHECK_LAUNCH();
}
void CoarseBlockField::pack_ptrs(const void *const *p) { block_copy_ptrs(*this, p, 1); }
void CoarseBlockField::unpack_ptrs(const void *const *p) const { block_copy_ptrs(const_cast<CoarseBlockField &>(*this), p, 0); }
void CoarseBlockField::pack(SpinorField *const *f) { block_copy(*this, f, 1); }
void CoarseBlockField::unpack(SpinorField *const *f) const { block_copy(const_cast<CoarseBlockField &>(*this), f, 0); }

// ---- the kernel -------------------------------------------------------------------------------------------------
struct MrhsKernelArgs {
  const float4 *Ymma, *Xinv_mma;
  const float4 *ghost;   // ghost zone of in_hop (partitioned lattices): one block per ghost site, numbered from 2 Vh in nbr
  const int *nbr;
  float4 *out;
  const float4 *in_hop, *in_diag, *xpay;
  long out_poff[2], hop_poff[2], diag_poff[2], xpay_poff[2];  // float4 offsets of the parity blocks
  long Vh, nsites;
  int parity;  // -1: all sites
  int use_y, diag_kind;  // diag_kind: 0 none, 1 X (= L_8), 2 Xinv
  int R, nstages, nwork;  // depths of the link ring and of the work ring
  float a, b;
  int has_xpay;
  int variant;  // tuning experiments (QB_MRHS_VARIANT): 0 = production
};

constexpr int MRHS_BUILD_GROUPS = 2;     // builder groups of 128 threads on alternate stages (a third group gained nothing: 0.417 vs 0.411 ms at 12 rhs)
constexpr int MRHS_THREADS = 64 + 128 * MRHS_BUILD_GROUPS + 128;
constexpr int MRHS_MAX_STAGES = 10;  // depth limit of the link ring
constexpr int MRHS_MAX_WORK = 4;     // depth limit of the work ring (A tile, lo part of the link tile)

// Shared memory plan (all offsets multiples of 128 B):
//   work ring  [3 or 4]     : A tile (rows x NKC x 16 B, rows = live rows rounded up to 8) | MODE 3: lo part of the link tile
//   link ring  [S]          : link tile as copied from HBM (8 N^2 B) | block of R source vectors as copied (NKC x R x 16 B)
//   staging, barriers
// The link ring is as deep as shared memory allows: a B200 SM needs ~100 KB in flight to pull its share of the HBM
// bandwidth, and the (site, direction) stream is perfectly predictable, so the producer simply runs S tiles ahead.
// The MMA reads M (64 / 128) rows of the A tile although only `rows` are live: the rows beyond alias the following
// shared memory; they only produce garbage in accumulator rows nobody reads (TMEM rows are independent).
template <int N, int M, int MODE> struct MrhsCfg {
  static constexpr int NKC = N / 2;          // 16-byte chunks (2 complex numbers) along K per direction
  static constexpr int KSTEPS = NKC / 2;     // tf32 MMAs (K = 8) per direction
  static constexpr int Y_BYTES = N * NKC * 16;
  // threads per builder group; the two groups take alternate stages.  Round-2 measurement (tools/mrhs_bench.py): the kernel time grows
  // linearly with the number of right-hand sides although the MMA count per site does not -- the builders' shared-memory work (and not the
  // tcgen05 issue rate: a second issuing warp for disjoint accumulator groups changed nothing) is what binds MODE 3, so they get 8 warps
  static constexpr int BGROUP = 128;
  static constexpr int YPT = (N * NKC + BGROUP - 1) / BGROUP;  // B-tile float4 per builder thread (MODE 3 split)
  // Accumulator groups.  The tensor core rounds every accumulation toward zero, so the error of a long chain grows
  // linearly with its length; MODE 3 therefore keeps one TMEM accumulator per pair of directions and adds them in
  // fp32 (round to nearest) in the epilogue: chains of 2 * KSTEPS * 2 MMAs instead of 9 * KSTEPS * 2.
  static constexpr int G = MODE == 3 ? (N <= 48 ? 5 : 4) : 1;
  static constexpr int TMEM_NEED = 2 * G * N;
  static constexpr int TMEM_COLS = TMEM_NEED <= 32 ? 32 : TMEM_NEED <= 64 ? 64 : TMEM_NEED <= 128 ? 128 : TMEM_NEED <= 256 ? 256 : 512;
  static_assert(TMEM_NEED <= 512, "accumulators do not fit tensor memory");
  static constexpr int STG_LD = N + 1;
  static constexpr int BAR_BYTES = 512;
  __host__ __device__ static int rows(int R) { return ((MODE == 3 ? 4 : 2) * R + 7) & ~7; }
  __host__ __device__ static int a_bytes(int R) { return rows(R) * NKC * 16; }
  __host__ __device__ static int work_bytes(int R) { return ((a_bytes(R) + 127) & ~127) + (MODE == 3 ? Y_BYTES : 0); }
  __host__ __device__ static int raw_bytes(int R) { return ((NKC * R * 16) + 127) & ~127; }
  __host__ __device__ static int link_bytes(int R) { return Y_BYTES + raw_bytes(R); }
  __host__ __device__ static int stg_bytes(int R) { return (rows(R) * STG_LD * 4 + 127) & ~127; }
  static size_t smem_bytes(int S, int W, int R) { return (size_t)W * work_bytes(R) + (size_t)S * link_bytes(R) + stg_bytes(R) + BAR_BYTES + 128; }
  __host__ __device__ static int group_of(int d) { return d < 8 ? (G >= 4 ? d >> 1 : 0) : (G == 5 ? 4 : 0); }
};

template <int N, int M, int MODE>
__global__ void __launch_bounds__(MRHS_THREADS, 1) coarse_mrhs_kernel(const MrhsKernelArgs p) {
  using C = MrhsCfg<N, M, MODE>;
  constexpr int NKC = C::NKC;
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  const int S = p.nstages, NW = p.nwork;
  const int R = p.R;
  const int ROWS = C::rows(R);
  const int WORK = C::work_bytes(R), LINK = C::link_bytes(R);
  const int YLO_OFF = (C::a_bytes(R) + 127) & ~127;  // inside a work slot
  unsigned char *work0 = smem;
  unsigned char *link0 = smem + (size_t)NW * WORK;
  float *stg = (float *)(link0 + (size_t)S * LINK);
  uint64_t *bars = (uint64_t *)((unsigned char *)stg + C::stg_bytes(R));
  uint64_t *full = bars, *empty = bars + MRHS_MAX_STAGES;
  uint64_t *ab_ready = bars + 2 * MRHS_MAX_STAGES, *work_free = ab_ready + MRHS_MAX_WORK;
  uint64_t *acc_full = work_free + MRHS_MAX_WORK, *acc_empty = acc_full + 2;
  uint32_t *tmem_slot = (uint32_t *)(acc_empty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ndir = (p.use_y ? 8 : 0) + (p.diag_kind ? 1 : 0);
  const int dir0 = p.use_y ? 0 : 8;  // direction of stage j is dir0 + j
  const int nelem = NKC * R;         // float4 per block of R vectors

  if (tid == 0) {
    for (int s = 0; s < S; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int w = 0; w < NW; w++) { mbar_init(&ab_ready[w], C::BGROUP); mbar_init(&work_free[w], 1); }
    for (int a = 0; a < 2; a++) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 128); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  auto site_of = [&](long it, int &parity, long &cb) {
    if (p.parity < 0) { parity = it >= p.Vh ? 1 : 0; cb = it - (long)parity * p.Vh; }
    else { parity = p.parity; cb = it; }
  };

  if (warp == 0) {
    // ================= producer: link matrix + the block of R neighbour vectors of every (site, direction) =================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const uint32_t raw_bytes = (uint32_t)nelem * 16;
      int4 n0 = make_int4(0, 0, 0, 0), n1 = n0;
      auto load_nbr = [&](long it) {
        if (!p.use_y || it >= p.nsites) return;
        int parity; long cb;
        site_of(it, parity, cb);
        const int4 *q = (const int4 *)(p.nbr + ((long)parity * p.Vh + cb) * 8);
        n0 = __ldg(q); n1 = __ldg(q + 1);
      };
      load_nbr(blockIdx.x);
      for (long it = blockIdx.x; it < p.nsites; it += gridDim.x) {
        int parity; long cb;
        site_of(it, parity, cb);
        const long fs = (long)parity * p.Vh + cb;
        const int nb[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
        load_nbr(it + gridDim.x);  // in flight while this site's copies are issued
        for (int j = 0; j < ndir; j++) {
          const int d = dir0 + j;
          const float4 *ysrc = (d == 8 && p.diag_kind == 2) ? p.Xinv_mma + (size_t)fs * N * NKC : p.Ymma + ((size_t)fs * 9 + d) * N * NKC;
          const float4 *vsrc;
          if (d == 8) vsrc = p.in_diag + p.diag_poff[parity] + (size_t)cb * nelem;
          else {
            const int nfs = nb[d];
            if (nfs >= 2 * p.Vh) {
              vsrc = p.ghost + (size_t)(nfs - 2 * p.Vh) * nelem;
            } else {
              const int np = nfs >= p.Vh ? 1 : 0;
              vsrc = p.in_hop + p.hop_poff[np] + (size_t)(nfs - (long)np * p.Vh) * nelem;
            }
          }
          mbar_wait(&empty[stage], phase ^ 1);
          unsigned char *st = link0 + (size_t)stage * LINK;
          mbar_arrive_expect_tx(&full[stage], C::Y_BYTES + raw_bytes);
          bulk_g2s(st, ysrc, C::Y_BYTES, &full[stage]);
          bulk_g2s(st + C::Y_BYTES, vsrc, raw_bytes, &full[stage]);
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issue =================
    // The whole warp walks the loop (warp-uniform control flow, operands in uniform registers); one elected lane
    // issues.  Descriptors of consecutive K steps differ by a constant in their address field.
    constexpr uint32_t idesc = idesc_tf32(M, N);
    int stage = 0, w = 0; uint32_t wphase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    const uint32_t a_lbo = (uint32_t)ROWS * 16;
    const uint64_t a_step = (uint64_t)((2 * a_lbo) >> 4), y_step = (uint64_t)((2 * N * 16) >> 4);
    const uint64_t a_desc0 = smem_desc(smem_u32(work0), a_lbo, 128), y_desc0 = smem_desc(smem_u32(link0), N * 16, 128);
    const uint64_t l_desc0 = smem_desc(smem_u32(work0) + YLO_OFF, N * 16, 128);
    for (long it = blockIdx.x; it < p.nsites; it += gridDim.x) {
      mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tc_fence_after_sync();
      uint32_t started = 0;  // accumulator groups already written for this site
      for (int j = 0; j < ndir; j++) {
        const int g = C::group_of(dir0 + j);
        const uint32_t d_tmem = tmem_base + (acc * C::G + g) * N;
        mbar_wait(&ab_ready[w], wphase);  // implies full[stage]
        tc_fence_after_sync();
        const uint64_t ad0 = a_desc0 + (uint64_t)((w * WORK) >> 4);
        const uint64_t bd0 = y_desc0 + (uint64_t)((stage * LINK) >> 4);
        const uint64_t bl0 = l_desc0 + (uint64_t)((w * WORK) >> 4);
        const uint32_t first = (started >> g) & 1;
        if (elect_one()) {
          if (p.variant != 2) {
#pragma unroll
            for (int k = 0; k < C::KSTEPS; k++) {
              mma_tf32(d_tmem, ad0 + k * a_step, bd0 + k * y_step, idesc, k > 0 ? 1u : first);
              if (MODE == 3) mma_tf32(d_tmem, ad0 + k * a_step, bl0 + k * y_step, idesc, 1u);
            }
          }
          mma_commit(&empty[stage]);
          mma_commit(&work_free[w]);
        }
        __syncwarp();
        started |= 1u << g;
        if (++stage == S) stage = 0;
        if (++w == NW) { w = 0; wphase ^= 1; }
      }
      if (elect_one()) mma_commit(&acc_full[acc]);
      __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp < 2 + 4 * MRHS_BUILD_GROUPS) {
    // ================= builders: raw vectors -> A tile rows, (MODE 3) hi / lo split of the link tile =================
    // Two groups of 64 threads take alternate stages, so that two stages are always under construction.  A thread owns
    // fixed 16-byte slots of the A tile: slot = (chunk kc, row m), m = 2r + h (+ 2R for the lo rows), read from
    // raw[kc][r]; consecutive threads write consecutive slots (no bank conflicts) and the index math is done once.
    const int grp = (warp - 2) >> 2, bt = tid - 64 - grp * C::BGROUP;
    const int live = (MODE == 3 ? 4 : 2) * R;  // live rows
    const int nslot = NKC * live;
    // slot e = bt + u * BGROUP <-> (kc, m) advanced incrementally: no division in the loop
    const int dk = C::BGROUP / live, dm = C::BGROUP - dk * live;
    const int kc0 = bt / live, m0 = bt - kc0 * live;
    // this group's stage counter advances by 2; ring indices and phases are kept incrementally (no division)
    int stage = grp % S, w = grp % NW;
    uint32_t phase = (uint32_t)((grp / S) & 1), wphase = (uint32_t)((grp / NW) & 1);
    const long nstage_total = ((p.nsites - blockIdx.x + gridDim.x - 1) / gridDim.x) * ndir;
    const uint32_t link_s = smem_u32(link0), work_s = smem_u32(work0);
    constexpr int BATCH = 4;
    for (long c = grp; c < nstage_total; c += MRHS_BUILD_GROUPS) {
      mbar_wait(&full[stage], phase);
      mbar_wait(&work_free[w], wphase ^ 1);
      const uint32_t ls = link_s + (uint32_t)stage * LINK, ws = work_s + (uint32_t)w * WORK;
      const uint32_t raw = ls + C::Y_BYTES;
      if (p.variant != 3) {
        int kc = kc0, m = m0;
        for (int e0 = bt; e0 < nslot; e0 += BATCH * C::BGROUP) {
          float4 v[BATCH];
          int dst[BATCH], mrow[BATCH];
#pragma unroll
          for (int u = 0; u < BATCH; u++) {  // all loads of the batch first
            const int part = m >= 2 * R ? 1 : 0, mm = m - part * 2 * R;
            mrow[u] = m; dst[u] = kc * ROWS + m;
            if (e0 + u * C::BGROUP < nslot) v[u] = lds128(raw + (uint32_t)(kc * R + (mm >> 1)) * 16);
            m += dm; kc += dk;
            if (m >= live) { m -= live; kc++; }
          }
#pragma unroll
          for (int u = 0; u < BATCH; u++) {
            if (e0 + u * C::BGROUP < nslot) {
              float4 x = v[u];
              const int part = mrow[u] >= 2 * R ? 1 : 0, mm = mrow[u] - part * 2 * R;
              if (MODE == 3) {
                const float4 h = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
                if (part) x = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
                else x = h;
              }
              sts128(ws + (uint32_t)dst[u] * 16, (mm & 1) ? make_float4(x.y, x.x, x.w, x.z) : make_float4(x.x, -x.y, x.z, -x.w));
            }
          }
        }
      }
      if (MODE == 3) {
        const uint32_t yl = ws + YLO_OFF;
        for (int e0 = bt; e0 < N * NKC; e0 += BATCH * C::BGROUP) {
          float4 v[BATCH];
#pragma unroll
          for (int u = 0; u < BATCH; u++)
            if (e0 + u * C::BGROUP < N * NKC) v[u] = lds128(ls + (uint32_t)(e0 + u * C::BGROUP) * 16);
#pragma unroll
          for (int u = 0; u < BATCH; u++) {
            const int e = e0 + u * C::BGROUP;
            if (e < N * NKC) {
              const float4 x = v[u];
              // (measured, round 2: leaving the raw tile as the hi operand -- the tensor core truncates the low 13 bits, the parity tests pass
              // with lo = x - trunc(x) -- saves this store but not time: 0.437 vs 0.411 ms at 12 rhs)
              const float4 h = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
              sts128(ls + (uint32_t)e * 16, h);
              sts128(yl + (uint32_t)e * 16, make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w));
            }
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(&ab_ready[w]);
      stage += MRHS_BUILD_GROUPS; if (stage >= S) { stage -= S; phase ^= 1; }
      w += MRHS_BUILD_GROUPS; while (w >= NW) { w -= NW; wphase ^= 1; }
    }
  } else {
    // ================= epilogue =================
    const int et = tid - (64 + 128 * MRHS_BUILD_GROUPS);
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    int acc = 0; uint32_t acc_phase = 0;
    // accumulator row held by this thread: M = 128: lane i <-> row i; M = 64: rows 16 q' .. 16 q' + 15 sit in lanes 32 q' .. 32 q' + 15
    const int row = M == 128 ? q * 32 + lane : q * 16 + lane;
    const bool row_ok = (M == 128 ? true : lane < 16) && row < ROWS;
    uint32_t groups = 0;
    for (int j = 0; j < ndir; j++) groups |= 1u << C::group_of(dir0 + j);
    for (long it = blockIdx.x; it < p.nsites; it += gridDim.x) {
      int parity; long cb;
      site_of(it, parity, cb);
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after_sync();
      float v[N];
#pragma unroll
      for (int i = 0; i < N; i++) v[i] = 0.f;
#pragma unroll
      for (int g = 0; g < C::G; g++) {
        if (!((groups >> g) & 1)) continue;
        const uint32_t taddr = tmem_base + (acc * C::G + g) * N + ((uint32_t)(q * 32) << 16);
        float t[N];
#pragma unroll
        for (int c = 0; c < N / 16; c++) tmem_ld16(taddr + c * 16, t + c * 16);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < N; i++) v[i] += t[i];
      }
      tc_fence_before_sync();
      mbar_arrive(&acc_empty[acc]);
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < N; i++) stg[row * C::STG_LD + i] = v[i];
      }
      named_bar_sync(1, 128);
      float4 *out = p.out + p.out_poff[parity] + (size_t)cb * nelem;
      const float4 *xp = p.has_xpay ? p.xpay + p.xpay_poff[parity] + (size_t)cb * nelem : nullptr;
      for (int e = et; e < (p.variant == 4 ? 0 : nelem); e += 128) {
        const int kc = e / R, r = e - kc * R;
        const float *re = stg + (2 * r) * C::STG_LD + 2 * kc, *im = re + C::STG_LD;
        float4 o = make_float4(re[0], im[0], re[1], im[1]);
        if (MODE == 3) {
          const float *rl = re + 2 * R * C::STG_LD, *il = rl + C::STG_LD;
          o.x += rl[0]; o.y += il[0]; o.z += rl[1]; o.w += il[1];
        }
        if (xp) {
          const float4 x = xp[e];
          o.x = p.a * o.x + p.b * x.x; o.y = p.a * o.y + p.b * x.y; o.z = p.a * o.z + p.b * x.z; o.w = p.a * o.w + p.b * x.w;
        } else if (p.a != 1.f) { o.x *= p.a; o.y *= p.a; o.z *= p.a; o.w *= p.a; }
        out[e] = o;
      }
      named_bar_sync(1, 128);  // staging is reused by the next site
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

template <int N, int M, int MODE> static void launch_mrhs(MrhsKernelArgs &k) {
  using C = MrhsCfg<N, M, MODE>;
  int dev = 0, max_smem = 0, nsm = 0;
  QB_CUDA(cudaGetDevice(&dev));
  QB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  QB_CUDA(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  // the A tile is over-read up to M rows: keep M x NKC x 16 bytes of shared memory behind the last work slot (the link ring)
  // ring depths: at least 2 work slots and 3 link slots; then as many link slots as fit (they hide the HBM latency),
  // with a third / fourth work slot when the link ring is already deep
  int W = 2, S = 3;
  if (C::smem_bytes(S, W, k.R) > (size_t)max_smem) QB_ERROR("multi-RHS coarse kernel: %d right-hand sides do not fit shared memory", k.R);
  // MODE 3: the builders do the hi / lo split of the link tile and need a third work slot more than a deep link ring
  if (MODE == 3 && C::smem_bytes(4, 3, k.R) <= (size_t)max_smem) { W = 3; S = 4; }
  while (S < MRHS_MAX_STAGES && C::smem_bytes(S + 1, W, k.R) <= (size_t)max_smem) S++;
  while (W < MRHS_MAX_WORK && S > 6 && C::smem_bytes(S - 1, W + 1, k.R) <= (size_t)max_smem) { W++; S--; }
  while (W < MRHS_MAX_WORK && C::smem_bytes(S, W + 1, k.R) <= (size_t)max_smem) W++;
  if (getenv("QB_MRHS_S")) S = atoi(getenv("QB_MRHS_S"));
  if (getenv("QB_MRHS_W")) W = atoi(getenv("QB_MRHS_W"));
  if (C::smem_bytes(S, W, k.R) > (size_t)max_smem) QB_ERROR("multi-RHS coarse kernel: ring depths %d / %d do not fit", S, W);
  k.nstages = S; k.nwork = W;
  const size_t sm = C::smem_bytes(S, W, k.R);
  QB_CUDA(cudaFuncSetAttribute(coarse_mrhs_kernel<N, M, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  const unsigned grid = (unsigned)(k.nsites < nsm ? k.nsites : nsm);
  coarse_mrhs_kernel<N, M, MODE><<<grid, MRHS_THREADS, sm, rt().compute>>>(k);
  QB_CHECK_LAUNCH();
}

template <int N> static bool mrhs_fits(int R, int mode, size_t max_smem) {
  const int rows = R * (mode == 3 ? 4 : 2);
  if (R < 1 || rows > 128) return false;
  if (rows <= 64) return (mode == 3 ? MrhsCfg<N, 64, 3>::smem_bytes(3, 2, R) : MrhsCfg<N, 64, 1>::smem_bytes(3, 2, R)) <= max_smem;
  return (mode == 3 ? MrhsCfg<N, 128, 3>::smem_bytes(3, 2, R) : MrhsCfg<N, 128, 1>::smem_bytes(3, 2, R)) <= max_smem;
}
// largest number of right-hand sides one launch takes (MMA tile rows and shared memory), 0 if N is not instantiated
int coarse_mrhs_max_rhs(int N, int mode) {
  int dev = 0, max_smem = 0;
  QB_CUDA(cudaGetDevice(&dev));
  QB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  for (int R = 64; R >= 1; R--) {
    bool ok = false;
    switch (N) {
      case 16: ok = mrhs_fits<16>(R, mode, max_smem); break;
      case 32: ok = mrhs_fits<32>(R, mode, max_smem); break;
      case 48: ok = mrhs_fits<48>(R, mode, max_smem); break;
      case 64: ok = mrhs_fits<64>(R, mode, max_smem); break;
      default: return 0;
    }
    if (ok) return R;
  }
  return 0;
}

void coarse_apply_mrhs(const CoarseMrhsArgs &a) {
  const CoarseOperator &op = *a.op;
  if (!op.Ymma || !op.nbr) QB_ERROR("coarse_apply_mrhs: call CoarseOperator::prepare_mrhs() first");
  if (a.use_xinv && !op.Xinv_mma) QB_ERROR("coarse_apply_mrhs: Xinv is not available in the tensor-core layout");
  if (a.mode != 1 && a.mode != 3) QB_ERROR("coarse_apply_mrhs: mode must be 1 (tf32) or 3 (split tf32, fp32-accurate)");
  MrhsKernelArgs k{};
  k.Ymma = (const float4 *)op.Ymma; k.Xinv_mma = (const float4 *)op.Xinv_mma; k.nbr = op.nbr;
  k.out = (float4 *)a.out; k.in_hop = (const float4 *)a.in_hop; k.in_diag = (const float4 *)a.in_diag; k.xpay = (const float4 *)a.xpay;
  for (int p = 0; p < 2; p++) { k.out_poff[p] = a.out_poff[p]; k.hop_poff[p] = a.hop_poff[p]; k.diag_poff[p] = a.diag_poff[p]; k.xpay_poff[p] = a.xpay_poff[p]; }
  k.Vh = op.geom.Vh; k.nsites = a.parity < 0 ? op.geom.V() : op.geom.Vh;
  k.parity = a.parity; k.use_y = a.use_y ? 1 : 0; k.diag_kind = a.use_xinv ? 2 : (a.use_x ? 1 : 0);
  k.R = a.R; k.a = a.a; k.b = a.b; k.has_xpay = a.xpay ? 1 : 0;
  k.variant = getenv("QB_MRHS_VARIANT") ? atoi(getenv("QB_MRHS_VARIANT")) : 0;
  if (!k.use_y && !k.diag_kind) QB_ERROR("coarse_apply_mrhs: nothing to apply");
  if (k.use_y && op.geom.partitioned()) {
    // halo of the hop input: the parities the output sites read from (both for a full-lattice apply)
    op.exchange_block_ghost(a.in_hop, a.hop_poff, a.parity < 0 ? 3 : (1 << (1 - a.parity)), a.R);
    k.ghost = (const float4 *)op.mrhs_recv;
  }
  const int rows = a.R * (a.mode == 3 ? 4 : 2);
  if (a.R < 1 || a.R > coarse_mrhs_max_rhs(op.N, a.mode))
    QB_ERROR("coarse_apply_mrhs: %d right-hand sides do not fit one launch in mode %d (N = %d: at most %d)", a.R, a.mode, op.N, coarse_mrhs_max_rhs(op.N, a.mode));
#define MRHS_BY_MODE(NN, MM)                              \
  if (a.mode == 3) launch_mrhs<NN, MM, 3>(k);             \
  else launch_mrhs<NN, MM, 1>(k)
#define MRHS_BY_M(NN)                                     \
  if (rows <= 64 && !getenv("QB_MRHS_FORCE_M128")) { MRHS_BY_MODE(NN, 64); }               \
  else { MRHS_BY_MODE(NN, 128); }
  switch (op.N) {
    case 16: MRHS_BY_M(16); break;
    case 32: MRHS_BY_M(32); break;
    case 48: MRHS_BY_M(48); break;
    case 64: MRHS_BY_M(64); break;
    default: QB_ERROR("multi-RHS coarse operator with n_vec = %d is not instantiated (supported: 8, 16, 24, 32)", op.nvec);
  }
#undef MRHS_BY_M
#undef MRHS_BY_MODE
}

}  // namespace qb
