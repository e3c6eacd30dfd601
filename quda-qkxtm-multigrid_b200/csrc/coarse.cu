// Coarse-level operator: storage, the coarse Dslash kernel, batched site-block inversion, DiracCoarse.
#include "blas.h"
#include <cuda_fp16.h>
#include "coarse.h"
#include "comm.h"
#include "layout.cuh"

namespace qb {

void CoarseOperator::allocate(const LevelGeom &g, int nvec_) {
  geom = g;
  nvec = nvec_;
  N = 2 * nvec;
  if (Y) pool_free(Y);
  Y = (float *)pool_malloc(link_bytes());   // GB-sized link arrays go through the caching allocator
  QB_CUDA(cudaMemsetAsync(Y, 0, link_bytes(), rt().compute));
  size_t off = 0;
  for (int d = 0; d < 4; d++) {
    if (!geom.part[d]) continue;
    const size_t bytes = (size_t)2 * (N / 2) * geom.faceVh[d] * sizeof(float4);
    for (int k = 0; k < 2; k++) {
      QB_CUDA(cudaMalloc((void **)&send[d][k], bytes));
      if (comm_self_exchange()) recv[d][1 - k] = send[d][k];  // my back face is my own forward ghost and vice versa
      recv_off[d][k] = off;
      off += (bytes + 255) & ~(size_t)255;
    }
  }
  if (off && !comm_self_exchange()) {
    ghost_arena.create(off);   // collective: every rank builds its coarse operators in the same order
    for (int d = 0; d < 4; d++)
      for (int k = 0; k < 2; k++)
        if (geom.part[d]) recv[d][k] = (float *)(ghost_arena.local + recv_off[d][k]);
  }
}

CoarseOperator::~CoarseOperator() {
  if (Y) pool_free(Y);
  if (Xinv) pool_free(Xinv);
  if (Y16) pool_free(Y16);
  if (Xinv16) pool_free(Xinv16);
  if (Yhat) pool_free(Yhat);
  if (Yhat16) pool_free(Yhat16);
  if (Ymma) pool_free(Ymma);
  if (Xinv_mma) pool_free(Xinv_mma);
  if (nbr) cudaFree(nbr);
  if (mrhs_send) cudaFree(mrhs_send);
  mrhs_arena.destroy();
  ghost_arena.destroy();
  for (int d = 0; d < 4; d++)
    for (int k = 0; k < 2; k++)
      if (send[d][k]) cudaFree(send[d][k]);
}

// ---- coarse halo: plain copies of the boundary slices (no spin projection on coarse levels,
// cf. GenericPackGhostKernel, lib/color_spinor_pack.cu) ------------------------------------------------
struct CoarsePackArgs {
  float4 *send[4][2];
  const float4 *field;
  long poff[2];
  int X[4];
  long Vh;
  int faceVh[4], part[4];
  int nplanes, parity_mask;
  long off[5];  // prefix sums of 2 * 2 * nplanes * faceVh[d] over partitioned dims
};

__device__ __forceinline__ long coarse_face_to_cb(int mu, int fidx, int slice, int parity, const int *X) {
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

__global__ void coarse_pack_kernel(const CoarsePackArgs a) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.off[4]) return;
  int d = 0;
  while (d < 3 && t >= a.off[d + 1]) d++;
  long r = t - a.off[d];
  const int fv = a.faceVh[d];
  const int fidx = (int)(r % fv); r /= fv;
  const int pl = (int)(r % a.nplanes); r /= a.nplanes;
  const int parity = (int)(r & 1), dir = (int)(r >> 1);
  if (!((a.parity_mask >> parity) & 1)) return;
  const long cb = coarse_face_to_cb(d, fidx, dir ? a.X[d] - 1 : 0, parity, a.X);
  a.send[d][dir][((size_t)parity * a.nplanes + pl) * fv + fidx] = a.field[a.poff[parity] + (size_t)pl * a.Vh + cb];
}

void CoarseOperator::exchange_ghost(const float *field, const long *poff, int parity_mask) const {
  if (!geom.partitioned()) return;
  Runtime &r = rt();
  const bool peer = ghost_arena.peer;
  if (peer) ghost_arena.seq++;
  HaloFlags sig{}, wt{};
  sig.seq = wt.seq = ghost_arena.seq;
  CoarsePackArgs a;
  a.field = (const float4 *)field;
  a.poff[0] = poff[0]; a.poff[1] = poff[1];
  a.Vh = geom.Vh; a.nplanes = N / 2; a.parity_mask = parity_mask;
  long off = 0;
  for (int d = 0; d < 4; d++) {
    a.X[d] = geom.X[d]; a.faceVh[d] = geom.faceVh[d]; a.part[d] = geom.part[d];
    a.send[d][0] = (float4 *)send[d][0]; a.send[d][1] = (float4 *)send[d][1];
    if (peer && geom.part[d])
      for (int dir = 0; dir < 2; dir++) {
        // my face `dir` (0: slice 0, travels backward) lands in the receive block (d, 1 - dir) of that neighbour's current buffer
        const int nb = comm_neighbor_rank(d, dir);
        a.send[d][dir] = (float4 *)(ghost_arena.send_base(nb) + recv_off[d][1 - dir]);
        sig.p[sig.n++] = ghost_arena.flag_of(nb, d * 2 + (1 - dir));
        wt.p[wt.n++] = ghost_arena.flag_mine(d * 2 + dir);
      }
    a.off[d] = off;
    if (geom.part[d]) off += 2L * 2 * a.nplanes * geom.faceVh[d];
  }
  a.off[4] = off;
  coarse_pack_kernel<<<div_up(off, 256), 256, 0, r.compute>>>(a);
  QB_CHECK_LAUNCH();
  if (comm_self_exchange()) return;
  if (peer) {
    comm_halo_signal(sig, r.compute);
    comm_halo_wait(wt, r.compute);
    return;
  }
  // all faces of all partitioned dimensions in one NCCL group (one launch; the reference posts them all before waiting too,
  // lib/dslash_coarse.cu:707).  my back face -> backward neighbour's "from forward" ghost; my forward face -> forward neighbour's
  // "from backward" ghost
  const void *sb[8]; void *rb[8]; int to[8], from[8]; size_t nb[8];
  int n = 0;
  for (int d = 0; d < 4; d++) {
    if (!geom.part[d]) continue;
    const size_t bytes = (size_t)2 * a.nplanes * geom.faceVh[d] * sizeof(float4);
    sb[n] = send[d][0]; to[n] = comm_neighbor_rank(d, 0); rb[n] = recv[d][1]; from[n] = comm_neighbor_rank(d, 1); nb[n++] = bytes;
    sb[n] = send[d][1]; to[n] = comm_neighbor_rank(d, 1); rb[n] = recv[d][0]; from[n] = comm_neighbor_rank(d, 0); nb[n++] = bytes;
  }
  comm_sendrecv_group(n, sb, to, rb, from, nb, r.compute);
}

// -----------------------------------------------------------------------------------------------------
// Coarse Dslash.  One CTA per coarse site; thread t = (direction d = t / (N/2), row pair rp = t % (N/2)).
// Every thread streams one column-major strip of its direction's matrix: N float4 loads (two rows x
// one column each), 2 complex FMAs per load against the neighbour vector staged in shared memory, no
// cross-thread traffic until the final 9-way sum.  Per site the CTA reads 9 N^2 x 8 B contiguous bytes
// exactly once (166 KB for N = 48): the kernel is a pure HBM stream (AI ~ 1 flop/B, dslash_coarse.cu:346-354).
// -----------------------------------------------------------------------------------------------------
struct CoarseKernelArgs {
  const float4 *Y;
  const float4 *Xinv;
  const uint2 *Y16, *Xinv16;   // the same matrices as 4 x fp16 per element pair (CoarseOperator::enable_half_links)
  float4 *out;
  const float4 *in_hop, *in_diag, *xpay;
  long out_poff[2], hop_poff[2], diag_poff[2], xpay_poff[2];
  int X[4];
  int Xh;
  long Vh;
  int parity;      // -1: all sites (blockIdx.x = full index), else cb index of that parity
  int use_y, use_x, use_xinv;
  float a, b;
  int part[4], faceVh[4];
  const float4 *ghost[4][2];  // [d][0]: from the backward, [d][1]: from the forward neighbour; [parity][plane][faceVh]
};

template <int N, bool H16>
__global__ void __launch_bounds__(((9 * (N / 2) + 31) / 32) * 32) coarse_dslash_kernel(const CoarseKernelArgs p) {
  constexpr int NRP = N / 2;
  __shared__ float2 xin[9][N + 1];
  __shared__ float4 red[9][NRP];
  __shared__ int nbr_par[9];
  __shared__ int nbr_ghost[9];
  __shared__ long nbr_cb[9];

  const long Vh = p.Vh;
  int parity;
  long cb;
  if (p.parity < 0) { parity = blockIdx.x >= Vh ? 1 : 0; cb = blockIdx.x - (long)parity * Vh; }
  else { parity = p.parity; cb = blockIdx.x; }

  const int t = threadIdx.x;
  const int d = t / NRP, rp = t - d * NRP;
  const bool active = d < 9;

  if (t < 9) {
    const long za = cb / p.Xh, zb = za / p.X[1];
    int x[4];
    x[1] = (int)(za - zb * p.X[1]);
    x[3] = (int)(zb / p.X[2]);
    x[2] = (int)(zb - (long)x[3] * p.X[2]);
    x[0] = (int)(2 * cb + ((x[1] + x[2] + x[3] + parity) & 1) - za * p.X[0]);
    nbr_ghost[t] = 0;
    if (t < 8) {
      const int mu = t >> 1;
      const bool edge = (t & 1) ? (x[mu] == 0) : (x[mu] == p.X[mu] - 1);
      nbr_par[t] = 1 - parity;
      if (edge && p.part[mu]) {
        nbr_ghost[t] = 1;
        const int d0 = mu == 0 ? 1 : 0, d1 = mu <= 1 ? 2 : 1, d2 = mu <= 2 ? 3 : 2;
        nbr_cb[t] = (x[d0] + p.X[d0] * (x[d1] + (long)p.X[d1] * x[d2])) >> 1;
      } else {
        x[mu] = (x[mu] + ((t & 1) ? p.X[mu] - 1 : 1)) % p.X[mu];
        nbr_cb[t] = ((((long)x[3] * p.X[2] + x[2]) * p.X[1] + x[1]) * p.X[0] + x[0]) >> 1;
      }
    } else {
      nbr_par[t] = parity;
      nbr_cb[t] = cb;
    }
  }
  __syncthreads();

  // stage the nine input vectors (N complex each)
  if (active) {
    const bool diag = d == 8;
    const bool need = diag ? (p.use_x || p.use_xinv) : p.use_y;
    if (need) {
      float4 v;
      if (!diag && nbr_ghost[d]) {
        const int mu = d >> 1;
        v = __ldg(p.ghost[mu][(d & 1) ? 0 : 1] + ((size_t)nbr_par[d] * NRP + rp) * p.faceVh[mu] + nbr_cb[d]);
      } else {
        const float4 *src = diag ? p.in_diag + p.diag_poff[nbr_par[8]] : p.in_hop + p.hop_poff[nbr_par[d]];
        v = __ldg(src + (size_t)rp * Vh + nbr_cb[d]);
      }
      xin[d][2 * rp] = make_float2(v.x, v.y);
      xin[d][2 * rp + 1] = make_float2(v.z, v.w);
    }
  }
  __syncthreads();

  cplx<float> a0(0.f, 0.f), a1(0.f, 0.f);
  if (active) {
    const bool diag = d == 8;
    const bool need = diag ? (p.use_x || p.use_xinv) : p.use_y;
    if (need) {
      const long site = (long)parity * Vh + cb;
      const size_t moff = (diag && p.use_xinv) ? (size_t)site * N * NRP + rp : ((size_t)site * 9 + d) * N * NRP + rp;
      const float4 *M = ((diag && p.use_xinv) ? p.Xinv : p.Y) + moff;
      const uint2 *M16 = ((diag && p.use_xinv) ? p.Xinv16 : p.Y16) + moff;
#pragma unroll 8
      for (int c = 0; c < N; c++) {
        float4 y;
        if (H16) {
          const int2 h = ld_stream((const int2 *)(M16 + (size_t)c * NRP));
          const float2 lo = __half22float2(*reinterpret_cast<const __half2 *>(&h.x)), hi = __half22float2(*reinterpret_cast<const __half2 *>(&h.y));
          y = make_float4(lo.x, lo.y, hi.x, hi.y);
        } else {
          y = ld_stream(M + (size_t)c * NRP);
        }
        const float2 xv = xin[d][c];
        const cplx<float> xc(xv.x, xv.y);
        cmac(a0, cplx<float>(y.x, y.y), xc);
        cmac(a1, cplx<float>(y.z, y.w), xc);
      }
    }
    red[d][rp] = make_float4(a0.re, a0.im, a1.re, a1.im);
  }
  __syncthreads();
  if (t < NRP) {
    float4 s = red[0][t];
#pragma unroll
    for (int k = 1; k < 9; k++) { const float4 r = red[k][t]; s.x += r.x; s.y += r.y; s.z += r.z; s.w += r.w; }
    const size_t o = (size_t)t * Vh + cb;
    if (p.xpay) {
      const float4 xp = p.xpay[p.xpay_poff[parity] + o];
      s.x = p.a * s.x + p.b * xp.x; s.y = p.a * s.y + p.b * xp.y; s.z = p.a * s.z + p.b * xp.z; s.w = p.a * s.w + p.b * xp.w;
    }
    p.out[p.out_poff[parity] + o] = s;
  }
}

template <int N> static void launch_coarse(const CoarseKernelArgs &k, long nsites) {
  constexpr int threads = ((9 * (N / 2) + 31) / 32) * 32;
  if (k.Y16) coarse_dslash_kernel<N, true><<<(unsigned)nsites, threads, 0, rt().compute>>>(k);
  else coarse_dslash_kernel<N, false><<<(unsigned)nsites, threads, 0, rt().compute>>>(k);
  QB_CHECK_LAUNCH();
}

// fp16 copies of the link matrices for the single-RHS kernel (preconditioner storage precision, as later QUDA versions do for the
// coarse links): the kernel is bound by streaming Y, so halving its bytes halves its time; arithmetic stays fp32
__global__ void links_to_half_kernel(uint2 *dst, const float4 *src, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 v = src[i];
  const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
  dst[i] = make_uint2(*reinterpret_cast<const unsigned *>(&lo), *reinterpret_cast<const unsigned *>(&hi));
}
void CoarseOperator::enable_half_links() {
  const size_t ny = (size_t)geom.V() * 9 * N * (N / 2), nx = (size_t)geom.V() * N * (N / 2);
  if (!Y16) Y16 = pool_malloc(ny * sizeof(uint2));
  links_to_half_kernel<<<(unsigned)div_up((long)ny, 256), 256, 0, rt().compute>>>((uint2 *)Y16, (const float4 *)Y, ny);
  if (Xinv) {
    if (!Xinv16) Xinv16 = pool_malloc(nx * sizeof(uint2));
    links_to_half_kernel<<<(unsigned)div_up((long)nx, 256), 256, 0, rt().compute>>>((uint2 *)Xinv16, (const float4 *)Xinv, nx);
  }
  if (Yhat) {
    if (!Yhat16) Yhat16 = pool_malloc(ny * sizeof(uint2));
    links_to_half_kernel<<<(unsigned)div_up((long)ny, 256), 256, 0, rt().compute>>>((uint2 *)Yhat16, (const float4 *)Yhat, ny);
  }
  QB_CHECK_LAUNCH();
}

void coarse_apply(const CoarseApplyArgs &a) {
  const CoarseOperator &op = *a.op;
  CoarseKernelArgs k;
  if (a.use_yhat && !op.Yhat) QB_ERROR("coarse_apply: Yhat has not been computed");
  k.Y = (const float4 *)(a.use_yhat ? op.Yhat : op.Y); k.Xinv = (const float4 *)op.Xinv;
  const void *y16 = a.use_yhat ? op.Yhat16 : op.Y16;
  const bool h16 = y16 && (!a.use_xinv || op.Xinv16) && !a.force_fp32;
  k.Y16 = h16 ? (const uint2 *)y16 : nullptr; k.Xinv16 = h16 ? (const uint2 *)op.Xinv16 : nullptr;
  k.out = (float4 *)a.out; k.in_hop = (const float4 *)a.in_hop; k.in_diag = (const float4 *)a.in_diag; k.xpay = (const float4 *)a.xpay;
  for (int p = 0; p < 2; p++) { k.out_poff[p] = a.out_poff[p]; k.hop_poff[p] = a.hop_poff[p]; k.diag_poff[p] = a.diag_poff[p]; k.xpay_poff[p] = a.xpay_poff[p]; }
  for (int d = 0; d < 4; d++) k.X[d] = op.geom.X[d];
  k.Xh = op.geom.Xh; k.Vh = op.geom.Vh;
  k.parity = a.parity; k.use_y = a.use_y; k.use_x = a.use_x; k.use_xinv = a.use_xinv; k.a = a.a; k.b = a.b;
  for (int d = 0; d < 4; d++) {
    k.part[d] = op.geom.part[d]; k.faceVh[d] = op.geom.faceVh[d];
  }
  if (a.use_y && op.geom.partitioned()) {
    // halo of the hop input: the parities the output sites read from (both for a full-lattice apply)
    op.exchange_ghost(a.in_hop, a.hop_poff, a.parity < 0 ? 3 : (1 << (1 - a.parity)));
  }
  // (after the exchange: with direct peer delivery the receive buffer alternates with the exchange sequence number)
  const size_t gbuf = op.ghost_arena.buf();
  for (int d = 0; d < 4; d++)
    for (int kdir = 0; kdir < 2; kdir++) k.ghost[d][kdir] = op.recv[d][kdir] ? (const float4 *)((const char *)op.recv[d][kdir] + gbuf) : nullptr;
  if (a.use_xinv && !op.Xinv) QB_ERROR("coarse_apply: Xinv has not been computed");
  const long nsites = a.parity < 0 ? op.geom.V() : op.geom.Vh;
  switch (op.N) {
    case 4: launch_coarse<4>(k, nsites); break;
    case 8: launch_coarse<8>(k, nsites); break;
    case 12: launch_coarse<12>(k, nsites); break;
    case 16: launch_coarse<16>(k, nsites); break;
    case 24: launch_coarse<24>(k, nsites); break;
    case 32: launch_coarse<32>(k, nsites); break;
    case 40: launch_coarse<40>(k, nsites); break;
    case 48: launch_coarse<48>(k, nsites); break;
    case 64: launch_coarse<64>(k, nsites); break;
    default: QB_ERROR("coarse operator with n_vec = %d is not instantiated (supported: 2, 4, 6, 8, 12, 16, 20, 24, 32)", op.nvec);
  }
}

// -----------------------------------------------------------------------------------------------------
// Xinv: batched inverse of the N x N site-diagonal blocks.  One CTA per site, Gauss-Jordan with partial
// pivoting on [A | 1] held in shared memory, fp32 storage with the pivot search / scaling in fp32
// (the blocks are 1 + O(kappa) perturbations, condition number O(1)).  The reference delegates this to
// MAGMA 1.7 getrf/getri_batched (lib/blas_magma.cu:1016-1140); no reference test pins it, our tests
// check |X Xinv - 1|.
// -----------------------------------------------------------------------------------------------------
template <int N> __global__ void __launch_bounds__(256) xinv_kernel(float4 *Xinv, const float4 *Y, long nsites) {
  constexpr int NRP = N / 2;
  extern __shared__ float2 smem[];
  float2 *A = smem;              // [row][col], N x N
  float2 *B = smem + N * N;      // inverse being built
  __shared__ int piv;
  const long site = blockIdx.x;
  const float4 *src = Y + ((size_t)site * 9 + 8) * N * NRP;
  for (int e = threadIdx.x; e < N * NRP; e += blockDim.x) {
    const int c = e / NRP, rp = e - c * NRP;
    const float4 v = src[e];
    A[(2 * rp) * N + c] = make_float2(v.x, v.y);
    A[(2 * rp + 1) * N + c] = make_float2(v.z, v.w);
  }
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) B[e] = make_float2((e / N) == (e % N) ? 1.f : 0.f, 0.f);
  __syncthreads();
  for (int k = 0; k < N; k++) {
    if (threadIdx.x == 0) {
      int best = k;
      float bm = A[k * N + k].x * A[k * N + k].x + A[k * N + k].y * A[k * N + k].y;
      for (int r = k + 1; r < N; r++) {
        const float m = A[r * N + k].x * A[r * N + k].x + A[r * N + k].y * A[r * N + k].y;
        if (m > bm) { bm = m; best = r; }
      }
      piv = best;
    }
    __syncthreads();
    const int pr = piv;
    if (pr != k) {
      for (int c = threadIdx.x; c < N; c += blockDim.x) {
        float2 t = A[k * N + c]; A[k * N + c] = A[pr * N + c]; A[pr * N + c] = t;
        t = B[k * N + c]; B[k * N + c] = B[pr * N + c]; B[pr * N + c] = t;
      }
      __syncthreads();
    }
    const float2 pv = A[k * N + k];
    const float inv = 1.0f / (pv.x * pv.x + pv.y * pv.y);
    const float2 ip = make_float2(pv.x * inv, -pv.y * inv);
    __syncthreads();
    for (int c = threadIdx.x; c < N; c += blockDim.x) {
      float2 a = A[k * N + c], b = B[k * N + c];
      A[k * N + c] = make_float2(a.x * ip.x - a.y * ip.y, a.x * ip.y + a.y * ip.x);
      B[k * N + c] = make_float2(b.x * ip.x - b.y * ip.y, b.x * ip.y + b.y * ip.x);
    }
    __syncthreads();
    // eliminate column k from every other row; factors are read before the row is modified
    for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
      const int r = e / N, c = e - r * N;
      if (r == k) continue;
      const float2 f = A[r * N + k];
      if (c == k) continue;  // handled after the sweep (needs the old A[r][k])
      const float2 ak = A[k * N + c];
      A[e].x -= f.x * ak.x - f.y * ak.y;
      A[e].y -= f.x * ak.y + f.y * ak.x;
    }
    for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
      const int r = e / N, c = e - r * N;
      if (r == k) continue;
      const float2 f = A[r * N + k];
      const float2 bk = B[k * N + c];
      B[e].x -= f.x * bk.x - f.y * bk.y;
      B[e].y -= f.x * bk.y + f.y * bk.x;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < N; r += blockDim.x)
      if (r != k) A[r * N + k] = make_float2(0.f, 0.f);
    __syncthreads();
  }
  float4 *dst = Xinv + (size_t)site * N * NRP;
  for (int e = threadIdx.x; e < N * NRP; e += blockDim.x) {
    const int c = e / NRP, rp = e - c * NRP;
    const float2 r0 = B[(2 * rp) * N + c], r1 = B[(2 * rp + 1) * N + c];
    dst[e] = make_float4(r0.x, r0.y, r1.x, r1.y);
  }
}

template <int N> static void launch_xinv(float *Xinv, const float *Y, long nsites) {
  const size_t sm = (size_t)2 * N * N * sizeof(float2);
  QB_CUDA(cudaFuncSetAttribute(xinv_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  xinv_kernel<N><<<(unsigned)nsites, 256, sm, rt().compute>>>((float4 *)Xinv, (const float4 *)Y, nsites);
  QB_CHECK_LAUNCH();
}

void CoarseOperator::compute_xinv() {
  if (!Xinv) Xinv = (float *)pool_malloc((size_t)geom.V() * N * N * 8);
  switch (N) {
    case 4: launch_xinv<4>(Xinv, Y, geom.V()); break;
    case 8: launch_xinv<8>(Xinv, Y, geom.V()); break;
    case 12: launch_xinv<12>(Xinv, Y, geom.V()); break;
    case 16: launch_xinv<16>(Xinv, Y, geom.V()); break;
    case 24: launch_xinv<24>(Xinv, Y, geom.V()); break;
    case 32: launch_xinv<32>(Xinv, Y, geom.V()); break;
    case 40: launch_xinv<40>(Xinv, Y, geom.V()); break;
    case 48: launch_xinv<48>(Xinv, Y, geom.V()); break;
    case 64: launch_xinv<64>(Xinv, Y, geom.V()); break;
    default: QB_ERROR("Xinv for n_vec = %d is not instantiated", nvec);
  }
}

// -----------------------------------------------------------------------------------------------------
// Yhat_d(x) = Xinv(x) L_d(x), d < 8; slot 8 = identity.  One CTA per (site, d): both N x N matrices staged in shared memory
// as [row][col], thread = output element.  8 N^3 flops per (site, d): milliseconds once per setup.
// -----------------------------------------------------------------------------------------------------
template <int N> __global__ void __launch_bounds__(256) yhat_kernel(float4 *Yhat, const float4 *Y, const float4 *Xinv) {
  constexpr int NRP = N / 2;
  extern __shared__ float2 smem[];
  float2 *A = smem;          // Xinv [row][col]
  float2 *B = smem + N * N;  // Y_d  [row][col]
  const long site = blockIdx.x;
  const int d = blockIdx.y;
  float2 *dst = (float2 *)(Yhat + ((size_t)site * 9 + d) * N * NRP);
  if (d == 8) {
    for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
      const int r = e / N, c = e - r * N;
      dst[((size_t)c * NRP + (r >> 1)) * 2 + (r & 1)] = make_float2(r == c ? 1.f : 0.f, 0.f);
    }
    return;
  }
  const float4 *xs = Xinv + (size_t)site * N * NRP, *ys = Y + ((size_t)site * 9 + d) * N * NRP;
  for (int e = threadIdx.x; e < N * NRP; e += blockDim.x) {
    const int c = e / NRP, rp = e - c * NRP;
    const float4 a = xs[e], b = ys[e];
    A[(2 * rp) * N + c] = make_float2(a.x, a.y); A[(2 * rp + 1) * N + c] = make_float2(a.z, a.w);
    B[(2 * rp) * N + c] = make_float2(b.x, b.y); B[(2 * rp + 1) * N + c] = make_float2(b.z, b.w);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
    const int c = e / N, r = e - c * N;   // consecutive threads: consecutive rows of one column (contiguous stores)
    cplx<float> acc(0.f, 0.f);
#pragma unroll 4
    for (int k = 0; k < N; k++) {
      const float2 a = A[r * N + k], b = B[k * N + c];
      cmac(acc, cplx<float>(a.x, a.y), cplx<float>(b.x, b.y));
    }
    dst[((size_t)c * NRP + (r >> 1)) * 2 + (r & 1)] = make_float2(acc.re, acc.im);
  }
}

template <int N> static void launch_yhat(float *Yhat, const float *Y, const float *Xinv, long nsites) {
  const size_t sm = (size_t)2 * N * N * sizeof(float2);
  QB_CUDA(cudaFuncSetAttribute(yhat_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  yhat_kernel<N><<<dim3((unsigned)nsites, 9), 256, sm, rt().compute>>>((float4 *)Yhat, (const float4 *)Y, (const float4 *)Xinv);
  QB_CHECK_LAUNCH();
}

void CoarseOperator::compute_yhat() {
  if (!Xinv) compute_xinv();
  if (!Yhat) Yhat = (float *)pool_malloc(link_bytes());
  switch (N) {
    case 4: launch_yhat<4>(Yhat, Y, Xinv, geom.V()); break;
    case 8: launch_yhat<8>(Yhat, Y, Xinv, geom.V()); break;
    case 12: launch_yhat<12>(Yhat, Y, Xinv, geom.V()); break;
    case 16: launch_yhat<16>(Yhat, Y, Xinv, geom.V()); break;
    case 24: launch_yhat<24>(Yhat, Y, Xinv, geom.V()); break;
    case 32: launch_yhat<32>(Yhat, Y, Xinv, geom.V()); break;
    case 40: launch_yhat<40>(Yhat, Y, Xinv, geom.V()); break;
    case 48: launch_yhat<48>(Yhat, Y, Xinv, geom.V()); break;
    case 64: launch_yhat<64>(Yhat, Y, Xinv, geom.V()); break;
    default: QB_ERROR("Yhat for n_vec = %d is not instantiated", nvec);
  }
}

// -----------------------------------------------------------------------------------------------------
// DiracCoarse
// -----------------------------------------------------------------------------------------------------
static void field_offsets(long *poff, const SpinorField &f, int parity_of_single) {
  // float4 offsets of the even / odd blocks inside f.v; a single-parity field serves `parity_of_single`
  const long block = (long)(f.parity_bytes / 16);
  if (f.nparity == 2) { poff[0] = 0; poff[1] = block; }
  else { poff[0] = poff[1] = 0; (void)parity_of_single; }
}

static void check_coarse_field(const SpinorField &f, const CoarseOperator &op) {
  if (f.prec != PREC_SINGLE || f.ncomplex != op.N || f.Vh != op.geom.Vh) QB_ERROR("coarse operator: field does not match the coarse lattice (N=%d)", op.N);
}

void DiracCoarse::Dslash(SpinorField &out, const SpinorField &in, int parity) const {
  check_coarse_field(out, *op); check_coarse_field(in, *op);
  CoarseApplyArgs a{};
  a.op = op.get(); a.out = (float *)out.v; a.in_hop = (const float *)in.v; a.in_diag = nullptr; a.xpay = nullptr;
  field_offsets(a.out_poff, out, parity); field_offsets(a.hop_poff, in, 1 - parity);
  a.parity = parity; a.use_y = true; a.use_x = false; a.use_xinv = false; a.a = 1.f; a.b = 0.f;
  coarse_apply(a);
  flops += 8ll * 8 * op->N * op->N * op->geom.Vh;
}

void DiracCoarse::DslashXpay(SpinorField &out, const SpinorField &in, int parity, const SpinorField &x, double k) const {
  check_coarse_field(out, *op); check_coarse_field(in, *op); check_coarse_field(x, *op);
  CoarseApplyArgs a{};
  a.op = op.get(); a.out = (float *)out.v; a.in_hop = (const float *)in.v; a.xpay = (const float *)x.v;
  field_offsets(a.out_poff, out, parity); field_offsets(a.hop_poff, in, 1 - parity); field_offsets(a.xpay_poff, x, parity);
  a.parity = parity; a.use_y = true; a.a = (float)k; a.b = 1.f;
  coarse_apply(a);
  flops += 8ll * 8 * op->N * op->N * op->geom.Vh;
}

void DiracCoarse::Clover(SpinorField &out, const SpinorField &in, int parity) const {
  CoarseApplyArgs a{};
  a.op = op.get(); a.out = (float *)out.v; a.in_diag = (const float *)in.v;
  field_offsets(a.out_poff, out, parity); field_offsets(a.diag_poff, in, parity);
  a.parity = parity; a.use_x = true; a.a = 1.f;
  coarse_apply(a);
  flops += 8ll * op->N * op->N * op->geom.Vh;
}

void DiracCoarse::CloverInv(SpinorField &out, const SpinorField &in, int parity) const {
  CoarseApplyArgs a{};
  a.op = op.get(); a.out = (float *)out.v; a.in_diag = (const float *)in.v;
  field_offsets(a.out_poff, out, parity); field_offsets(a.diag_poff, in, parity);
  a.parity = parity; a.use_xinv = true; a.a = 1.f;
  coarse_apply(a);
  flops += 8ll * op->N * op->N * op->geom.Vh;
}

static SpinorField &ensure_tmp(std::unique_ptr<SpinorField> &t, const SpinorField &like) {
  if (!t || t->Vh != like.Vh || t->nparity != like.nparity || t->ncomplex != like.ncomplex) t.reset(new SpinorField(like.Vh, like.nparity, PREC_SINGLE, like.nspin, like.ncolor));
  return *t;
}

void DiracCoarse::M(SpinorField &out, const SpinorField &in) const {
  check_coarse_field(out, *op); check_coarse_field(in, *op);
  if (dagger) QB_ERROR("DiracCoarse: the daggered coarse operator is not implemented");
  if (!pc) {
    if (in.nparity != 2 || out.nparity != 2) QB_ERROR("DiracCoarse::M needs full fields");
    CoarseApplyArgs a{};
    a.op = op.get(); a.out = (float *)out.v; a.in_hop = (const float *)in.v; a.in_diag = (const float *)in.v;
    field_offsets(a.out_poff, out, 0); field_offsets(a.hop_poff, in, 0); field_offsets(a.diag_poff, in, 0);
    a.parity = -1; a.use_y = true; a.use_x = true; a.a = 1.f;
    coarse_apply(a);
    flops += 8ll * 9 * op->N * op->N * op->geom.V();
    return;
  }
  // symmetric even-odd Schur complement: out = in - Xinv_p Y_pq Xinv_q Y_qp in
  if (in.nparity != 1 || out.nparity != 1) QB_ERROR("DiracCoarsePC::M needs single-parity fields");
  const int p = p_parity(), q = 1 - p;
  if (op->Yhat) {
    // out = in - Yhat_pq Yhat_qp in: two launches, no separate Xinv pass (DiracCoarsePC::M, lib/dirac_coarse.cpp:245-283)
    SpinorField &t = ensure_tmp(tmp1, in);
    CoarseApplyArgs a{};
    a.op = op.get(); a.out = (float *)t.v; a.in_hop = (const float *)in.v;
    field_offsets(a.out_poff, t, q); field_offsets(a.hop_poff, in, p);
    a.parity = q; a.use_y = true; a.use_yhat = true; a.a = 1.f; a.b = 0.f;
    coarse_apply(a);
    CoarseApplyArgs b{};
    b.op = op.get(); b.out = (float *)out.v; b.in_hop = (const float *)t.v; b.xpay = (const float *)in.v;
    field_offsets(b.out_poff, out, p); field_offsets(b.hop_poff, t, q); field_offsets(b.xpay_poff, in, p);
    b.parity = p; b.use_y = true; b.use_yhat = true; b.a = -1.f; b.b = 1.f;
    coarse_apply(b);
    flops += 2ll * 8 * 8 * op->N * op->N * op->geom.Vh;
    return;
  }
  SpinorField &t1 = ensure_tmp(tmp1, in), &t2 = ensure_tmp(tmp2, in);
  Dslash(t1, in, q);
  CloverInv(t2, t1, q);
  Dslash(t1, t2, p);
  // out = in - Xinv_p t1
  CoarseApplyArgs a{};
  a.op = op.get(); a.out = (float *)out.v; a.in_diag = (const float *)t1.v; a.xpay = (const float *)in.v;
  field_offsets(a.out_poff, out, p); field_offsets(a.diag_poff, t1, p); field_offsets(a.xpay_poff, in, p);
  a.parity = p; a.use_xinv = true; a.a = -1.f; a.b = 1.f;
  coarse_apply(a);
  flops += 8ll * op->N * op->N * op->geom.Vh;
}

// M x = b on the full coarse lattice through the Schur complement (symmetric preconditioning):
//   src = Xinv_p (b_p - Y_pq Xinv_q b_q);   x_q = Xinv_q (b_q - Y_qp x_p)     (dirac_coarse.cpp:285-372)
void DiracCoarse::prepare(SpinorField &src, SpinorField &sol, SpinorField &x, SpinorField &b, SolutionType sol_type) const {
  if (!pc) {
    b.view_parity(src, 0); src.nparity = b.nparity;
    x.view_parity(sol, 0); sol.nparity = x.nparity;
    return;
  }
  if (sol_type == SOL_MATPC || sol_type == SOL_MATPCDAG_MATPC) { b.view_parity(src, 0); x.view_parity(sol, 0); return; }
  const int p = p_parity(), q = 1 - p;
  SpinorField bp, bq;
  b.view_parity(bp, p); b.view_parity(bq, q);
  x.view_parity(src, q);
  x.view_parity(sol, p);
  SpinorField &t1 = ensure_tmp(tmp1, bp), &t2 = ensure_tmp(tmp2, bp);
  CloverInv(t1, bq, q);
  DslashXpay(t2, t1, p, bp, -1.0);  // t2 = b_p - Y_pq Xinv_q b_q
  CloverInv(src, t2, p);
}

void DiracCoarse::reconstruct(SpinorField &x, const SpinorField &b, SolutionType sol_type) const {
  if (!pc) return;
  if (sol_type == SOL_MATPC || sol_type == SOL_MATPCDAG_MATPC) return;
  const int p = p_parity(), q = 1 - p;
  SpinorField xp, xq, bq;
  x.view_parity(xp, p); x.view_parity(xq, q); b.view_parity(bq, q);
  SpinorField &t1 = ensure_tmp(tmp1, xp);
  DslashXpay(t1, xp, q, bq, -1.0);  // b_q - Y_qp x_p
  CloverInv(xq, t1, q);
}

void DiracCoarse::create_coarse_op(CoarseOperator &coarse, const Transfer &T, bool preconditioned) const {
  if (preconditioned && !op->Yhat) op->compute_yhat();
  build_coarse_from_coarse(coarse, T, *op, preconditioned);
}

// X^-1 on every site of a full field
void DiracCoarse::DiagInv(SpinorField &out, const SpinorField &in) const {
  if (in.nparity != 2 || out.nparity != 2) QB_ERROR("DiracCoarse::DiagInv needs full fields");
  CoarseApplyArgs a{};
  a.op = op.get(); a.out = (float *)out.v; a.in_diag = (const float *)in.v;
  field_offsets(a.out_poff, out, 0); field_offsets(a.diag_poff, in, 0);
  a.parity = -1; a.use_xinv = true; a.a = 1.f;
  coarse_apply(a);
}

}  // namespace qb
