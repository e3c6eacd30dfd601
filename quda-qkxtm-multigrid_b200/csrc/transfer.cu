// Transfer operator: geometry maps, V assembly, block Gram-Schmidt, prolongator and restrictor kernels.
#include <algorithm>
#include <vector>
#include "layout.cuh"
#include <cstdlib>
#include <cuda_fp16.h>
#include "transfer.h"
#include "comm.h"

namespace qb {

// ---- geometry ------------------------------------------------------------------------------------
static inline void cb_to_coords(int *x, long cb, int parity, const LevelGeom &g) {
  const long za = cb / g.Xh, zb = za / g.X[1];
  x[1] = (int)(za - zb * g.X[1]);
  x[3] = (int)(zb / g.X[2]);
  x[2] = (int)(zb - (long)x[3] * g.X[2]);
  x[0] = (int)(2 * cb + ((x[1] + x[2] + x[3] + parity) & 1) - za * g.X[0]);
}
static inline long coords_to_full(const int *x, const LevelGeom &g) {
  const long lex = (((long)x[3] * g.X[2] + x[2]) * g.X[1] + x[1]) * g.X[0] + x[0];
  const int parity = (x[0] + x[1] + x[2] + x[3]) & 1;
  return (long)parity * g.Vh + (lex >> 1);
}

// ---- kernels ---------------------------------------------------------------------------------------
// V[(parity*Nf + k) * nvec/2 + j/2][cb] float4, complex j%2 inside
__device__ __forceinline__ size_t v_plane(int parity, int k, int jp, int Nf, int nvh) { return ((size_t)parity * Nf + k) * nvh + jp; }

__global__ void fill_v_kernel(float4 *V, const float4 *B, long Vh, int Nf, int nvec, int j) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nkp = Nf / 2;
  if (t >= 2 * Vh * nkp) return;
  const long cb = t % Vh;
  const int kp = (int)((t / Vh) % nkp), parity = (int)(t / (Vh * nkp));
  const float4 b = B[((size_t)parity * nkp + kp) * Vh + cb];  // field layout [parity][plane][cb]
  float2 *v0 = (float2 *)(V + v_plane(parity, 2 * kp, j / 2, Nf, nvec / 2) * Vh + cb) + (j & 1);
  float2 *v1 = (float2 *)(V + v_plane(parity, 2 * kp + 1, j / 2, Nf, nvec / 2) * Vh + cb) + (j & 1);
  *v0 = make_float2(b.x, b.y);
  *v1 = make_float2(b.z, b.w);
}

__device__ __forceinline__ double2 block_sum(double2 v, double2 *sm) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v.x += __shfl_xor_sync(0xffffffffu, v.x, o);
    v.y += __shfl_xor_sync(0xffffffffu, v.y, o);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  double2 s = make_double2(0.0, 0.0);
  for (int w = 0; w < nw; w++) { s.x += sm[w].x; s.y += sm[w].y; }
  return s;
}

// One CTA per (aggregate, chirality): modified Gram-Schmidt over the nvec block vectors, fp64 sums
// (order of operations of blockGramSchmidt, lib/transfer_util.cu:327-363).
__global__ void __launch_bounds__(256) block_ortho_kernel(float *V, const int *c2f, long Vh_f, int Nf, int nvec, int block_sites) {
  __shared__ double2 sm[8];
  const int X = blockIdx.x, S = blockIdx.y;
  const int cpc = Nf / 2;            // components per chirality
  const int E = block_sites * cpc;   // block-vector length
  const int nvh = nvec / 2;
  auto addr = [&](int e, int j) -> float2 * {
    const int i = e / cpc, k = S * cpc + (e - i * cpc);
    const int fs = c2f[(size_t)X * block_sites + i];
    const int parity = fs >= Vh_f ? 1 : 0;
    const long cb = fs - (long)parity * Vh_f;
    return (float2 *)((float4 *)V + v_plane(parity, k, j >> 1, Nf, nvh) * Vh_f + cb) + (j & 1);
  };
  for (int jc = 0; jc < nvec; jc++) {
    for (int ic = 0; ic < jc; ic++) {
      double2 d = make_double2(0.0, 0.0);
      for (int e = threadIdx.x; e < E; e += blockDim.x) {
        const float2 a = *addr(e, ic), b = *addr(e, jc);
        d.x += (double)a.x * b.x + (double)a.y * b.y;
        d.y += (double)a.x * b.y - (double)a.y * b.x;
      }
      d = block_sum(d, sm);
      const float dr = (float)d.x, di = (float)d.y;
      for (int e = threadIdx.x; e < E; e += blockDim.x) {
        const float2 a = *addr(e, ic);
        float2 *pb = addr(e, jc);
        float2 b = *pb;
        b.x -= dr * a.x - di * a.y;
        b.y -= dr * a.y + di * a.x;
        *pb = b;
      }
      __syncthreads();
    }
    double2 n = make_double2(0.0, 0.0);
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
      const float2 b = *addr(e, jc);
      n.x += (double)b.x * b.x + (double)b.y * b.y;
    }
    n = block_sum(n, sm);
    const float scale = n.x > 0.0 ? (float)(1.0 / sqrt(n.x)) : 0.0f;
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
      float2 *pb = addr(e, jc);
      float2 b = *pb;
      b.x *= scale; b.y *= scale;
      *pb = b;
    }
    __syncthreads();
  }
}

// The same modified Gram-Schmidt, right-looking: step ic normalises v_ic and immediately removes it from ALL later vectors, 8 at a time --
// for every target vector the projections are applied in the same order (ic = 0, 1, ...) to the same data as in the column-by-column loop
// of the reference, so the arithmetic is identical up to the order of the fp64 sums.  What changes is the traffic: every thread owns NE
// elements (site fastest across the lanes: full 32-byte sectors), keeps the pivot in registers, and one pass over the trailing vectors
// gives 8 dot products at once instead of one.  The 295 KB of a 4^4 x 6 x 24 block stay in L2 (two CTAs per SM at most), so the
// ~11 MB of loads and stores per block never reach HBM: 223 ms -> ~30 ms for 32^3x64.
template <int NE>
__global__ void __launch_bounds__(256, 2) block_ortho_rl_kernel(float *V, const int *c2f, long Vh_f, int Nf, int nvec, int block_sites) {
  __shared__ double2 red[8][8];
  __shared__ double2 sm[8];
  const int X = blockIdx.x, S = blockIdx.y;
  const int cpc = Nf / 2, E = block_sites * cpc, nvh = nvec / 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 *base[NE];
  bool ok[NE];
#pragma unroll
  for (int n = 0; n < NE; n++) {
    const int e = threadIdx.x + n * 256;
    ok[n] = e < E;
    const int ee = ok[n] ? e : 0;
    const int k = S * cpc + ee / block_sites, i = ee - (ee / block_sites) * block_sites;
    const int fs = c2f[(size_t)X * block_sites + i];
    const int parity = fs >= Vh_f ? 1 : 0;
    base[n] = (float4 *)V + v_plane(parity, k, 0, Nf, nvh) * Vh_f + (fs - (long)parity * Vh_f);
  }
  for (int ic = 0; ic < nvec; ic++) {
    // pivot: normalise, keep in registers
    float2 pv[NE];
    double2 nn = make_double2(0.0, 0.0);
#pragma unroll
    for (int n = 0; n < NE; n++) {
      pv[n] = make_float2(0.f, 0.f);
      if (ok[n]) {
        pv[n] = *((const float2 *)(base[n] + (size_t)(ic >> 1) * Vh_f) + (ic & 1));
        nn.x += (double)pv[n].x * pv[n].x + (double)pv[n].y * pv[n].y;
      }
    }
    nn = block_sum(nn, sm);
    const float scale = nn.x > 0.0 ? (float)(1.0 / sqrt(nn.x)) : 0.0f;
#pragma unroll
    for (int n = 0; n < NE; n++) {
      pv[n].x *= scale; pv[n].y *= scale;
      if (ok[n]) *((float2 *)(base[n] + (size_t)(ic >> 1) * Vh_f) + (ic & 1)) = pv[n];
    }
    // trailing vectors in groups of 4 pairs (8 vectors); a group may start with the pair that holds the pivot itself (masked below)
    for (int jp0 = (ic + 1) >> 1; jp0 < nvh; jp0 += 4) {
      double2 acc[8];
#pragma unroll
      for (int q = 0; q < 8; q++) acc[q] = make_double2(0.0, 0.0);
#pragma unroll
      for (int n = 0; n < NE; n++) {
        if (!ok[n]) continue;
        const float2 a = pv[n];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          if (jp0 + q >= nvh) break;
          const float4 w = base[n][(size_t)(jp0 + q) * Vh_f];
          acc[2 * q].x += (double)a.x * w.x + (double)a.y * w.y;
          acc[2 * q].y += (double)a.x * w.y - (double)a.y * w.x;
          acc[2 * q + 1].x += (double)a.x * w.z + (double)a.y * w.w;
          acc[2 * q + 1].y += (double)a.x * w.w - (double)a.y * w.z;
        }
      }
#pragma unroll
      for (int q = 0; q < 8; q++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          acc[q].x += __shfl_xor_sync(0xffffffffu, acc[q].x, o);
          acc[q].y += __shfl_xor_sync(0xffffffffu, acc[q].y, o);
        }
      }
      __syncthreads();
      if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 8; q++) red[warp][q] = acc[q];
      }
      __syncthreads();
      float dr[8], di[8];
#pragma unroll
      for (int q = 0; q < 8; q++) {
        double2 t = make_double2(0.0, 0.0);
#pragma unroll
        for (int w = 0; w < 8; w++) { t.x += red[w][q].x; t.y += red[w][q].y; }
        const bool live = 2 * jp0 + q > ic;   // the pivot's own pair: its earlier member and the pivot itself stay as they are
        dr[q] = live ? (float)t.x : 0.f;
        di[q] = live ? (float)t.y : 0.f;
      }
#pragma unroll
      for (int n = 0; n < NE; n++) {
        if (!ok[n]) continue;
        const float2 a = pv[n];
#pragma unroll
        for (int q = 0; q < 4; q++) {
          if (jp0 + q >= nvh) break;
          float4 *pw = base[n] + (size_t)(jp0 + q) * Vh_f;
          float4 w = *pw;
          w.x -= dr[2 * q] * a.x - di[2 * q] * a.y;
          w.y -= dr[2 * q] * a.y + di[2 * q] * a.x;
          w.z -= dr[2 * q + 1] * a.x - di[2 * q + 1] * a.y;
          w.w -= dr[2 * q + 1] * a.y + di[2 * q + 1] * a.x;
          *pw = w;
        }
      }
    }
  }
}

static bool launch_block_ortho_rl(float *V, const int *c2f, long Vh_f, int Nf, int nvec, int block_sites, long Vc, cudaStream_t s) {
  if (getenv("QB_BLOCK_ORTHO_OLD")) return false;
  const int E = block_sites * (Nf / 2);
  const int ne = div_up(E, 256);
  const dim3 grid((unsigned)Vc, 2);
#define QB_RL(N) block_ortho_rl_kernel<N><<<grid, 256, 0, s>>>(V, c2f, Vh_f, Nf, nvec, block_sites)
  if (ne <= 1) QB_RL(1);
  else if (ne == 2) QB_RL(2);
  else if (ne == 3) QB_RL(3);
  else if (ne == 4) QB_RL(4);
  else if (ne <= 6) QB_RL(6);
  else if (ne <= 8) QB_RL(8);
  else if (ne <= 12) QB_RL(12);
  else return false;
#undef QB_RL
  return true;
}

// V element loaders: fp32 planes (float4 = 2 complex) or the fp16 copy (uint2 = 2 complex, scaled by V16_SCALE so that the small
// entries of the block-orthonormal vectors stay normal numbers); arithmetic is fp32 either way
constexpr float V16_SCALE = 64.0f;
__device__ __forceinline__ float4 ldv(const float4 *p) { return ld_stream(p); }
__device__ __forceinline__ float4 ldv(const uint2 *p) {
  const int2 h = ld_stream((const int2 *)p);
  const float2 lo = __half22float2(*reinterpret_cast<const __half2 *>(&h.x)), hi = __half22float2(*reinterpret_cast<const __half2 *>(&h.y));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
template <typename VT> __device__ __forceinline__ float v_unscale() { return sizeof(VT) == sizeof(uint2) ? 1.0f / V16_SCALE : 1.0f; }
__global__ void v_to_half_kernel(uint2 *dst, const float4 *src, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 v = src[i];
  const __half2 lo = __floats2half2_rn(v.x * V16_SCALE, v.y * V16_SCALE), hi = __floats2half2_rn(v.z * V16_SCALE, v.w * V16_SCALE);
  dst[i] = make_uint2(*reinterpret_cast<const unsigned *>(&lo), *reinterpret_cast<const unsigned *>(&hi));
}

// thread = (fine site, component pair): out(x, k) = sum_j V(x,k,j) c(X, chi(k), j)
template <typename VT>
__global__ void __launch_bounds__(128) prolong_kernel(float4 *out, const float4 *cin, const VT *V, const int *f2c, long Vh_f, long Vh_c,
                                                      int Nf, int nvec, int cpc) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nkp = Nf / 2;
  if (t >= 2 * Vh_f * nkp) return;
  const long cb = t % Vh_f;
  const int kp = (int)((t / Vh_f) % nkp), parity = (int)(t / (Vh_f * nkp));
  const int k0 = 2 * kp, S = k0 / cpc, nvh = nvec / 2;
  const int cs = f2c[(size_t)parity * Vh_f + cb];
  const int cpar = cs >= Vh_c ? 1 : 0;
  const long ccb = cs - (long)cpar * Vh_c;
  // coarse field [parity][plane = (S*nvec + j)/2][cb]
  const float4 *c = cin + ((size_t)cpar * nvec + (size_t)S * nvh) * Vh_c + ccb;
  const VT *v0 = V + v_plane(parity, k0, 0, Nf, nvh) * Vh_f + cb;
  const VT *v1 = V + v_plane(parity, k0 + 1, 0, Nf, nvh) * Vh_f + cb;
  cplx<float> a0(0.f, 0.f), a1(0.f, 0.f);
#pragma unroll 4
  for (int jp = 0; jp < nvh; jp++) {
    const float4 cc = __ldg(c + (size_t)jp * Vh_c);
    const float4 w0 = ldv(v0 + (size_t)jp * Vh_f), w1 = ldv(v1 + (size_t)jp * Vh_f);
    cmac(a0, cplx<float>(w0.x, w0.y), cplx<float>(cc.x, cc.y));
    cmac(a0, cplx<float>(w0.z, w0.w), cplx<float>(cc.z, cc.w));
    cmac(a1, cplx<float>(w1.x, w1.y), cplx<float>(cc.x, cc.y));
    cmac(a1, cplx<float>(w1.z, w1.w), cplx<float>(cc.z, cc.w));
  }
  const float us = v_unscale<VT>();
  out[((size_t)parity * nkp + kp) * Vh_f + cb] = make_float4(us * a0.re, us * a0.im, us * a1.re, us * a1.im);
}

// CTA = one aggregate; threads = (32 site lanes) x (nvec/2 vector pairs).  Deterministic: fixed site order
// per lane, shuffle tree across lanes (the reference uses cub::BlockReduce, restrictor.cu:159-237).
template <int NKP>  // NKP = Nf/2 known at compile time (full unrolling => all V loads of a site in flight), 0 = generic
__global__ void restrict_kernel(float4 *out, const float4 *fin, const float4 *V, const int *c2f, long Vh_f, long Vh_c, int Nf, int nvec,
                                int cpc, int block_sites, int par, long in_pstride) {
  const int X = blockIdx.x;  // coarse full index
  const int lane = threadIdx.x, jp = threadIdx.y;
  const int nkp = NKP ? NKP : Nf / 2, nvh = nvec / 2;
  cplx<float> acc[2][2];  // [chirality][j within pair]
#pragma unroll
  for (int s = 0; s < 2; s++) { acc[s][0] = cplx<float>(0.f, 0.f); acc[s][1] = cplx<float>(0.f, 0.f); }
  for (int i = lane; i < block_sites; i += 32) {
    const int fs = c2f[(size_t)X * block_sites + i];
    const int parity = fs >= Vh_f ? 1 : 0;
    if (par >= 0 && parity != par) continue;   // single-parity input: the other parity counts as zero
    const long cb = fs - (long)parity * Vh_f;
    if (NKP > 0) {
      // issue every load of this site before the first use: 3 * NKP independent 128-bit requests in flight per thread
      float4 f[NKP ? NKP : 1], w0[NKP ? NKP : 1], w1[NKP ? NKP : 1];
#pragma unroll
      for (int kp = 0; kp < NKP; kp++) {
        f[kp] = __ldg(fin + (size_t)parity * in_pstride + (size_t)kp * Vh_f + cb);
        w0[kp] = ld_stream(V + v_plane(parity, 2 * kp, jp, Nf, nvh) * Vh_f + cb);
        w1[kp] = ld_stream(V + v_plane(parity, 2 * kp + 1, jp, Nf, nvh) * Vh_f + cb);
      }
#pragma unroll
      for (int kp = 0; kp < NKP; kp++) {
        const int S = (2 * kp) / cpc;
        const cplx<float> f0(f[kp].x, f[kp].y), f1(f[kp].z, f[kp].w);
        cmac_conj(acc[S][0], cplx<float>(w0[kp].x, w0[kp].y), f0); cmac_conj(acc[S][1], cplx<float>(w0[kp].z, w0[kp].w), f0);
        cmac_conj(acc[S][0], cplx<float>(w1[kp].x, w1[kp].y), f1); cmac_conj(acc[S][1], cplx<float>(w1[kp].z, w1[kp].w), f1);
      }
      continue;
    }
    for (int kp = 0; kp < nkp; kp++) {
      const float4 f = __ldg(fin + (size_t)parity * in_pstride + (size_t)kp * Vh_f + cb);
      const float4 w0 = ld_stream(V + v_plane(parity, 2 * kp, jp, Nf, nvh) * Vh_f + cb);
      const float4 w1 = ld_stream(V + v_plane(parity, 2 * kp + 1, jp, Nf, nvh) * Vh_f + cb);
      const int S = (2 * kp) / cpc;
      const cplx<float> f0(f.x, f.y), f1(f.z, f.w);
      if (S == 0) {
        cmac_conj(acc[0][0], cplx<float>(w0.x, w0.y), f0); cmac_conj(acc[0][1], cplx<float>(w0.z, w0.w), f0);
        cmac_conj(acc[0][0], cplx<float>(w1.x, w1.y), f1); cmac_conj(acc[0][1], cplx<float>(w1.z, w1.w), f1);
      } else {
        cmac_conj(acc[1][0], cplx<float>(w0.x, w0.y), f0); cmac_conj(acc[1][1], cplx<float>(w0.z, w0.w), f0);
        cmac_conj(acc[1][0], cplx<float>(w1.x, w1.y), f1); cmac_conj(acc[1][1], cplx<float>(w1.z, w1.w), f1);
      }
    }
  }
#pragma unroll
  for (int s = 0; s < 2; s++)
#pragma unroll
    for (int q = 0; q < 2; q++)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        acc[s][q].re += __shfl_xor_sync(0xffffffffu, acc[s][q].re, o);
        acc[s][q].im += __shfl_xor_sync(0xffffffffu, acc[s][q].im, o);
      }
  if (lane == 0) {
    const int cpar = X >= Vh_c ? 1 : 0;
    const long ccb = X - (long)cpar * Vh_c;
#pragma unroll
    for (int s = 0; s < 2; s++)
      out[((size_t)cpar * nvec + (size_t)s * nvh + jp) * Vh_c + ccb] = make_float4(acc[s][0].re, acc[s][0].im, acc[s][1].re, acc[s][1].im);
  }
}

// Row-major variant of the restrictor for fine lattices with Xh % 8 == 0 (every fine level of practical size).
// The aggregate-major kernel above gathers V in runs of bs_x / 2 sites (32 bytes for 4^4 blocks) scattered over the 144
// planes of V: every access opens another DRAM page, ~52 % of the HBM roofline.  Here a CTA owns the aggregates of one
// (by, bz, bt) block row over XL = min(Xh, 32) consecutive checkerboard x positions, and a warp reads RS = 32 / XL
// consecutive y rows at a time: when XL = Xh these are adjacent in memory, i.e. one contiguous 512-byte run per V plane,
// the access pattern of the prolongator.  lane <-> x position <-> aggregate, so each lane accumulates for one aggregate
// over all rows of the block; the final reduction is a fixed shuffle tree over the row lanes and the bs_x / 2 lanes of
// the aggregate (deterministic).
template <int NKP, typename VT>
__global__ void restrict_rows_kernel(float4 *out, const float4 *fin, const VT *V, const int *f2c, LevelGeom fg, int b0, int b1, int b2, int b3,
                                     long Vh_c, int nvec, int cpc, int XL, int par, long in_pstride) {
  const int lane = threadIdx.x, jp = threadIdx.y, nvh = nvec / 2;
  const int RS = 32 / XL;
  const int xl = lane % XL, rs = lane / XL;
  const long Vh_f = fg.Vh;
  const int ng = fg.Xh / XL, nby = fg.X[1] / b1, nbz = fg.X[2] / b2;
  int bid = blockIdx.x;
  const int g = bid % ng; bid /= ng;
  const int by = bid % nby; bid /= nby;
  const int bz = bid % nbz;
  const int bt = bid / nbz;
  const int nyg = b1 / RS;                 // groups of RS consecutive y rows inside the block
  const int npar = par < 0 ? 2 : 1;        // par >= 0: single-parity input, the other parity counts as zero and is never read
  const int niter = npar * nyg * b2 * b3;  // (parity, y group, z, t)
  cplx<float> acc[2][2];  // [chirality][j within pair]
#pragma unroll
  for (int s = 0; s < 2; s++) { acc[s][0] = cplx<float>(0.f, 0.f); acc[s][1] = cplx<float>(0.f, 0.f); }
  long first_fs = -1;
  for (int it = 0; it < niter; it++) {
    const int parity = par < 0 ? (it & 1) : par;
    int r = par < 0 ? (it >> 1) : it;
    const int yg = r % nyg; r /= nyg;
    const int y = by * b1 + yg * RS + rs, z = bz * b2 + r % b2, t = bt * b3 + r / b2;
    const long cb = (((long)t * fg.X[2] + z) * fg.X[1] + y) * fg.Xh + XL * g + xl;
    if (first_fs < 0) first_fs = (long)parity * Vh_f + cb;
    float4 f[NKP], w0[NKP], w1[NKP];
#pragma unroll
    for (int kp = 0; kp < NKP; kp++) {
      f[kp] = __ldg(fin + (size_t)parity * in_pstride + (size_t)kp * Vh_f + cb);
      w0[kp] = ldv(V + v_plane(parity, 2 * kp, jp, 2 * NKP, nvh) * Vh_f + cb);
      w1[kp] = ldv(V + v_plane(parity, 2 * kp + 1, jp, 2 * NKP, nvh) * Vh_f + cb);
    }
#pragma unroll
    for (int kp = 0; kp < NKP; kp++) {
      const int S = (2 * kp) / cpc;
      const cplx<float> f0(f[kp].x, f[kp].y), f1(f[kp].z, f[kp].w);
      cmac_conj(acc[S][0], cplx<float>(w0[kp].x, w0[kp].y), f0); cmac_conj(acc[S][1], cplx<float>(w0[kp].z, w0[kp].w), f0);
      cmac_conj(acc[S][0], cplx<float>(w1[kp].x, w1[kp].y), f1); cmac_conj(acc[S][1], cplx<float>(w1[kp].z, w1[kp].w), f1);
    }
  }
  {
    const float us = v_unscale<VT>();
#pragma unroll
    for (int s = 0; s < 2; s++)
#pragma unroll
      for (int qq = 0; qq < 2; qq++) { acc[s][qq].re *= us; acc[s][qq].im *= us; }
  }
  const int spa = b0 / 2;  // lanes (cb x positions) per aggregate
#pragma unroll
  for (int s = 0; s < 2; s++)
#pragma unroll
    for (int qq = 0; qq < 2; qq++) {
      float re = acc[s][qq].re, im = acc[s][qq].im;
      for (int o = 16; o >= XL; o >>= 1) { re += __shfl_xor_sync(0xffffffffu, re, o); im += __shfl_xor_sync(0xffffffffu, im, o); }
      for (int o = 1; o < spa; o <<= 1) { re += __shfl_xor_sync(0xffffffffu, re, o); im += __shfl_xor_sync(0xffffffffu, im, o); }
      acc[s][qq] = cplx<float>(re, im);
    }
  if (rs == 0 && (xl % spa) == 0) {
    const int X = f2c[first_fs];
    const int cpar = X >= Vh_c ? 1 : 0;
    const long ccb = X - (long)cpar * Vh_c;
#pragma unroll
    for (int s = 0; s < 2; s++)
      out[((size_t)cpar * nvec + (size_t)s * nvh + jp) * Vh_c + ccb] = make_float4(acc[s][0].re, acc[s][0].im, acc[s][1].re, acc[s][1].im);
  }
}

// ---- multi-right-hand-side variants (block multigrid, block_solver.cu): V is streamed once for NR vectors ------------
constexpr int TRANSFER_MAX_NR = 6;
struct MultiPtrs { const float4 *in[TRANSFER_MAX_NR]; float4 *out[TRANSFER_MAX_NR]; };

// prolong_kernel for NR coarse vectors at once; ACC: out += P c (the coarse-grid correction is added in place)
template <int NR, bool ACC, typename VT>
__global__ void __launch_bounds__(128) prolong_multi_kernel(MultiPtrs p, const VT *V, const int *f2c, long Vh_f, long Vh_c, int Nf, int nvec, int cpc,
                                                            int par, long out_pstride) {
  // par < 0: both parities of full fine fields; par = 0 / 1: only that parity (preconditioned coarsening: the coarse-grid correction
  // of the even-odd system lives on one parity, half of V is never touched), out_pstride = float4 distance of the parity blocks of out
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nkp = Nf / 2;
  if (t >= (par < 0 ? 2 : 1) * Vh_f * nkp) return;
  const long cb = t % Vh_f;
  const int kp = (int)((t / Vh_f) % nkp), parity = par < 0 ? (int)(t / (Vh_f * nkp)) : par;
  const int k0 = 2 * kp, S = k0 / cpc, nvh = nvec / 2;
  const int cs = f2c[(size_t)parity * Vh_f + cb];
  const int cpar = cs >= Vh_c ? 1 : 0;
  const long ccb = cs - (long)cpar * Vh_c;
  const size_t coff = ((size_t)cpar * nvec + (size_t)S * nvh) * Vh_c + ccb;
  const VT *v0 = V + v_plane(parity, k0, 0, Nf, nvh) * Vh_f + cb;
  const VT *v1 = V + v_plane(parity, k0 + 1, 0, Nf, nvh) * Vh_f + cb;
  cplx<float> a0[NR], a1[NR];
#pragma unroll
  for (int r = 0; r < NR; r++) { a0[r] = cplx<float>(0.f, 0.f); a1[r] = cplx<float>(0.f, 0.f); }
#pragma unroll 4
  for (int jp = 0; jp < nvh; jp++) {
    const float4 w0 = ldv(v0 + (size_t)jp * Vh_f), w1 = ldv(v1 + (size_t)jp * Vh_f);
#pragma unroll
    for (int r = 0; r < NR; r++) {
      const float4 cc = __ldg(p.in[r] + coff + (size_t)jp * Vh_c);
      cmac(a0[r], cplx<float>(w0.x, w0.y), cplx<float>(cc.x, cc.y));
      cmac(a0[r], cplx<float>(w0.z, w0.w), cplx<float>(cc.z, cc.w));
      cmac(a1[r], cplx<float>(w1.x, w1.y), cplx<float>(cc.x, cc.y));
      cmac(a1[r], cplx<float>(w1.z, w1.w), cplx<float>(cc.z, cc.w));
    }
  }
  const size_t o = (size_t)parity * out_pstride + (size_t)kp * Vh_f + cb;
#pragma unroll
  for (int r = 0; r < NR; r++) {
    const float us = v_unscale<VT>();
    float4 v = make_float4(us * a0[r].re, us * a0[r].im, us * a1[r].re, us * a1[r].im);
    if (ACC) { const float4 old = p.out[r][o]; v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w; }
    p.out[r][o] = v;
  }
}

// restrict_rows_kernel for NR fine vectors at once (same thread mapping and reduction tree: deterministic)
template <int NKP, int NR, typename VT>
__global__ void __launch_bounds__(384, 1) restrict_rows_multi_kernel(MultiPtrs p, const VT *V, const int *f2c, LevelGeom fg, int b0, int b1, int b2, int b3,
                                           long Vh_c, int nvec, int cpc, int XL, int par, long in_pstride) {
  const int lane = threadIdx.x, jp = threadIdx.y, nvh = nvec / 2;
  const int RS = 32 / XL;
  const int xl = lane % XL, rs = lane / XL;
  const long Vh_f = fg.Vh;
  const int ng = fg.Xh / XL, nby = fg.X[1] / b1, nbz = fg.X[2] / b2;
  int bid = blockIdx.x;
  const int g = bid % ng; bid /= ng;
  const int by = bid % nby; bid /= nby;
  const int bz = bid % nbz;
  const int bt = bid / nbz;
  const int nyg = b1 / RS;
  const int niter = (par < 0 ? 2 : 1) * nyg * b2 * b3;
  cplx<float> acc[NR][2][2];
#pragma unroll
  for (int r = 0; r < NR; r++)
#pragma unroll
    for (int s = 0; s < 2; s++) { acc[r][s][0] = cplx<float>(0.f, 0.f); acc[r][s][1] = cplx<float>(0.f, 0.f); }
  long first_fs = -1;
  for (int it = 0; it < niter; it++) {
    const int parity = par < 0 ? (it & 1) : par;
    int rr = par < 0 ? (it >> 1) : it;
    const int yg = rr % nyg; rr /= nyg;
    const int y = by * b1 + yg * RS + rs, z = bz * b2 + rr % b2, t = bt * b3 + rr / b2;
    const long cb = (((long)t * fg.X[2] + z) * fg.X[1] + y) * fg.Xh + XL * g + xl;
    if (first_fs < 0) first_fs = (long)parity * Vh_f + cb;
    float4 w0[NKP], w1[NKP];
#pragma unroll
    for (int kp = 0; kp < NKP; kp++) {
      w0[kp] = ldv(V + v_plane(parity, 2 * kp, jp, 2 * NKP, nvh) * Vh_f + cb);
      w1[kp] = ldv(V + v_plane(parity, 2 * kp + 1, jp, 2 * NKP, nvh) * Vh_f + cb);
    }
#pragma unroll
    for (int r = 0; r < NR; r++) {
#pragma unroll
      for (int kp = 0; kp < NKP; kp++) {
        const float4 f = __ldg(p.in[r] + (size_t)parity * in_pstride + (size_t)kp * Vh_f + cb);
        constexpr int CPC = NKP;            // components per chirality = Nf / 2 (two coarse spins): compile-time, keeps acc in registers
        const int S = (2 * kp) / CPC;
        const cplx<float> f0(f.x, f.y), f1(f.z, f.w);
        cmac_conj(acc[r][S][0], cplx<float>(w0[kp].x, w0[kp].y), f0); cmac_conj(acc[r][S][1], cplx<float>(w0[kp].z, w0[kp].w), f0);
        cmac_conj(acc[r][S][0], cplx<float>(w1[kp].x, w1[kp].y), f1); cmac_conj(acc[r][S][1], cplx<float>(w1[kp].z, w1[kp].w), f1);
      }
    }
  }
  const int spa = b0 / 2;
  const bool writer = rs == 0 && (xl % spa) == 0;
  int X = 0;
  if (writer) X = f2c[first_fs];
  const int cpar = X >= Vh_c ? 1 : 0;
  const long ccb = X - (long)cpar * Vh_c;
#pragma unroll
  for (int r = 0; r < NR; r++)
#pragma unroll
    for (int s = 0; s < 2; s++) {
      const float us = v_unscale<VT>();
      float v[4] = {us * acc[r][s][0].re, us * acc[r][s][0].im, us * acc[r][s][1].re, us * acc[r][s][1].im};
#pragma unroll
      for (int e = 0; e < 4; e++) {
        for (int o = 16; o >= XL; o >>= 1) v[e] += __shfl_xor_sync(0xffffffffu, v[e], o);
        for (int o = 1; o < spa; o <<= 1) v[e] += __shfl_xor_sync(0xffffffffu, v[e], o);
      }
      if (writer) p.out[r][((size_t)cpar * nvec + (size_t)s * nvh + jp) * Vh_c + ccb] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// face site (3-d lexicographic >> 1 of the remaining coordinates, as in the fine Dslash) -> cb index on slice x_mu = slice
__device__ __forceinline__ long level_face_to_cb(int mu, int fidx, int slice, int parity, const int *X) {
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

struct LevelDims { int X[4]; };

// gathers nplanes float4 planes per parity of the slice x_mu = slice into [parity][plane][faceVh]
__global__ void gather_slice_kernel(float4 *dst, const float4 *src, LevelDims dims, long Vh, int faceVh, int nplanes, int mu, int slice) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2L * nplanes * faceVh) return;
  const int fidx = (int)(t % faceVh);
  const int pl = (int)((t / faceVh) % nplanes), parity = (int)(t / ((long)faceVh * nplanes));
  const long cb = level_face_to_cb(mu, fidx, slice, parity, dims.X);
  dst[t] = src[((size_t)parity * nplanes + pl) * Vh + cb];
}

void gather_slice(float4 *dst, const float4 *src, const LevelGeom &g, int nplanes, int mu, int slice, cudaStream_t s) {
  LevelDims d;
  for (int k = 0; k < 4; k++) d.X[k] = g.X[k];
  const long n = 2L * nplanes * g.faceVh[mu];
  gather_slice_kernel<<<div_up(n, 256), 256, 0, s>>>(dst, src, d, g.Vh, g.faceVh[mu], nplanes, mu, slice);
  QB_CHECK_LAUNCH();
}

// ---- host ------------------------------------------------------------------------------------------
void Transfer::exchange_v_ghost() {
  if (!fine.partitioned()) return;
  Runtime &r = rt();
  const int nplanes = Nf * nvec / 2;
  for (int d = 0; d < 4; d++) {
    if (!fine.part[d]) continue;
    const size_t bytes = (size_t)2 * nplanes * fine.faceVh[d] * sizeof(float4);
    float4 *lo, *hi;  // my x_d = 0 and x_d = X_d - 1 slices
    QB_CUDA(cudaMalloc((void **)&lo, bytes));
    QB_CUDA(cudaMalloc((void **)&hi, bytes));
    gather_slice(lo, (const float4 *)V, fine, nplanes, d, 0, r.compute);
    gather_slice(hi, (const float4 *)V, fine, nplanes, d, fine.X[d] - 1, r.compute);
    if (comm_self_exchange()) {
      Vghost[d][0] = (float *)hi;  // backward neighbour (= myself) last slice
      Vghost[d][1] = (float *)lo;  // forward neighbour first slice
    } else {
      QB_CUDA(cudaMalloc((void **)&Vghost[d][0], bytes));
      QB_CUDA(cudaMalloc((void **)&Vghost[d][1], bytes));
      // my last slice -> forward neighbour's "from backward" ghost; my first slice -> backward neighbour's "from forward" ghost
      comm_sendrecv(hi, comm_neighbor_rank(d, 1), Vghost[d][0], comm_neighbor_rank(d, 0), bytes, r.compute);
      comm_sendrecv(lo, comm_neighbor_rank(d, 0), Vghost[d][1], comm_neighbor_rank(d, 1), bytes, r.compute);
      QB_CUDA(cudaStreamSynchronize(r.compute));
      QB_CUDA(cudaFree(lo));
      QB_CUDA(cudaFree(hi));
    }
  }
  QB_CUDA(cudaStreamSynchronize(r.compute));
}


Transfer::Transfer(const std::vector<SpinorField *> &B, int nvec_, int *bs, int spin_bs_, const int *fine_X) : nvec(nvec_), spin_bs(spin_bs_) {
  if ((int)B.size() < nvec) QB_ERROR("Transfer: %d null vectors supplied, %d needed", (int)B.size(), nvec);
  if (nvec < 2 || (nvec & 1)) QB_ERROR("Transfer: n_vec = %d must be even and >= 2", nvec);
  fine.set(fine_X);
  fine_nspin = B[0]->nspin; fine_ncolor = B[0]->ncolor; Nf = B[0]->ncomplex;
  if (fine_nspin % spin_bs || fine_nspin / spin_bs != 2) QB_ERROR("Transfer: spin block size %d does not give two chiralities for nSpin = %d", spin_bs, fine_nspin);
  if ((Nf / 2) & 1) QB_ERROR("Transfer: odd number of components per chirality (%d) is not supported", Nf / 2);
  for (auto *b : B)
    if (b->prec != PREC_SINGLE || b->nparity != 2 || b->Vh != fine.Vh) QB_ERROR("Transfer: null vectors must be full single-precision fields of the fine lattice");
  // block-size fix-up of the reference (transfer.cpp:31-44)
  int cX[4];
  block_sites = 1;
  for (int d = 0; d < 4; d++) {
    while (bs[d] > 0) {
      if (d == 0 && fine.X[0] == bs[0]) log_msg(1, "WARNING: X-dimension length %d cannot block length %d\n", fine.X[0], bs[0]);
      else if ((fine.X[d] / bs[d] + 1) % 2 == 0) log_msg(1, "WARNING: Indexing does not (yet) support odd coarse dimensions: X(%d) = %d\n", d, fine.X[d] / bs[d]);
      else if ((fine.X[d] / bs[d]) * bs[d] != fine.X[d]) log_msg(1, "WARNING: cannot block dim[%d]=%d with block size = %d\n", d, fine.X[d], bs[d]);
      else break;
      bs[d] /= 2;
    }
    if (bs[d] == 0) QB_ERROR("Unable to block dimension %d", d);
    geo_bs[d] = bs[d];
    cX[d] = fine.X[d] / bs[d];
    block_sites *= bs[d];
  }
  if (block_sites == 1) QB_ERROR("Total geometric block size is 1");
  coarse.set(cX);
  log_msg(2, "Transfer: using block size %d x %d x %d x %d, coarse lattice %d x %d x %d x %d, %d vectors\n", bs[0], bs[1], bs[2], bs[3], cX[0], cX[1], cX[2], cX[3], nvec);

  // geometry maps (createGeoMap, transfer.cpp:220-258): fine -> coarse by integer division of coordinates;
  // coarse -> fine = fine sites sorted by (coarse index, fine index)
  const long Vf = fine.V(), Vc = coarse.V();
  std::vector<int> h_f2c(Vf), h_c2f(Vf), fill(Vc, 0);
  for (int p = 0; p < 2; p++)
    for (long cb = 0; cb < fine.Vh; cb++) {
      int x[4], xc[4];
      cb_to_coords(x, cb, p, fine);
      for (int d = 0; d < 4; d++) xc[d] = x[d] / geo_bs[d];
      h_f2c[p * fine.Vh + cb] = (int)coords_to_full(xc, coarse);
    }
  for (long f = 0; f < Vf; f++) {  // ascending fine index => parity-0 sites first within each aggregate
    const int c = h_f2c[f];
    h_c2f[(size_t)c * block_sites + fill[c]++] = (int)f;
  }
  QB_CUDA(cudaMalloc((void **)&f2c, sizeof(int) * Vf));
  QB_CUDA(cudaMalloc((void **)&c2f, sizeof(int) * Vf));
  QB_CUDA(cudaMemcpy(f2c, h_f2c.data(), sizeof(int) * Vf, cudaMemcpyHostToDevice));
  QB_CUDA(cudaMemcpy(c2f, h_c2f.data(), sizeof(int) * Vf, cudaMemcpyHostToDevice));

  // V <- null vectors, then block orthonormalisation
  cudaStream_t s = rt().compute;
  V = (float *)pool_malloc(v_bytes());   // multi-GB: through the caching allocator (a second setup of the same shape, e.g. the DN flavour, reuses the block)
  const long nt = 2 * fine.Vh * (Nf / 2);
  for (int j = 0; j < nvec; j++) {
    fill_v_kernel<<<div_up(nt, 256), 256, 0, s>>>((float4 *)V, (const float4 *)B[j]->v, fine.Vh, Nf, nvec, j);
    QB_CHECK_LAUNCH();
  }
  if (!launch_block_ortho_rl(V, c2f, fine.Vh, Nf, nvec, block_sites, Vc, s))
    block_ortho_kernel<<<dim3((unsigned)Vc, 2), 256, 0, s>>>(V, c2f, fine.Vh, Nf, nvec, block_sites);
  QB_CHECK_LAUNCH();
  QB_CUDA(cudaStreamSynchronize(s));
  exchange_v_ghost();
}

Transfer::~Transfer() {
  if (V) pool_free(V);
  if (V16) pool_free(V16);
  if (f2c) cudaFree(f2c);
  if (c2f) cudaFree(c2f);
  for (int d = 0; d < 4; d++)
    for (int k = 0; k < 2; k++)
      if (Vghost[d][k]) cudaFree(Vghost[d][k]);
}

void Transfer::P(SpinorField &fo, const SpinorField &ci) const {
  if (fo.prec != PREC_SINGLE || ci.prec != PREC_SINGLE) QB_ERROR("Transfer::P works in single precision");
  if (fo.nparity != 2 || ci.nparity != 2 || fo.Vh != fine.Vh || ci.Vh != coarse.Vh || fo.ncomplex != Nf || ci.ncomplex != 2 * nvec)
    QB_ERROR("Transfer::P: field geometry mismatch");
  const long nt = 2 * fine.Vh * (Nf / 2);
  if (V16) prolong_kernel<uint2><<<div_up(nt, 128), 128, 0, rt().compute>>>((float4 *)fo.v, (const float4 *)ci.v, (const uint2 *)V16, f2c, fine.Vh, coarse.Vh, Nf, nvec, Nf / 2);
  else prolong_kernel<float4><<<div_up(nt, 128), 128, 0, rt().compute>>>((float4 *)fo.v, (const float4 *)ci.v, (const float4 *)V, f2c, fine.Vh, coarse.Vh, Nf, nvec, Nf / 2);
  QB_CHECK_LAUNCH();
  flops += 8ll * Nf * nvec * fine.V();
}

// fine fields of the parity-restricted variants: a single-parity field, or a full field of which only block `parity` is used
static long fine_pstride(const SpinorField &f, int Nf) { return f.nparity == 2 ? (long)(Nf / 2) * f.Vh : 0; }

void Transfer::R(SpinorField &co, const SpinorField &fi, int parity) const {
  if (co.prec != PREC_SINGLE || fi.prec != PREC_SINGLE) QB_ERROR("Transfer::R works in single precision");
  if ((parity < 0 && fi.nparity != 2) || co.nparity != 2 || fi.Vh != fine.Vh || co.Vh != coarse.Vh || fi.ncomplex != Nf || co.ncomplex != 2 * nvec)
    QB_ERROR("Transfer::R: field geometry mismatch");
  const long ips = fine_pstride(fi, Nf);
#define RK(NKP) restrict_kernel<NKP><<<(unsigned)coarse.V(), dim3(32, nvec / 2), 0, rt().compute>>>((float4 *)co.v, (const float4 *)fi.v, (const float4 *)V, c2f, \
                                                                                      fine.Vh, coarse.Vh, Nf, nvec, Nf / 2, block_sites, parity, ips)
  const int b0 = geo_bs[0];
  // lanes per row: the whole row if it fits a warp (contiguous runs over consecutive y rows), else pieces of 8
  int XL = 0;
  if (fine.Xh == 8 || fine.Xh == 16 || fine.Xh % 32 == 0) XL = fine.Xh < 32 ? fine.Xh : 32;
  else if (fine.Xh % 8 == 0) XL = 8;
  if (Nf == 12 && XL && (b0 == 2 || b0 == 4 || b0 == 8) && geo_bs[1] % (32 / XL) == 0 && !getenv("QB_RESTRICT_OLD")) {
    const unsigned nblk = (unsigned)((fine.Xh / XL) * (fine.X[1] / geo_bs[1]) * (fine.X[2] / geo_bs[2]) * (fine.X[3] / geo_bs[3]));
    if (V16) restrict_rows_kernel<6, uint2><<<nblk, dim3(32, nvec / 2), 0, rt().compute>>>((float4 *)co.v, (const float4 *)fi.v, (const uint2 *)V16, f2c, fine, b0,
                                                                                          geo_bs[1], geo_bs[2], geo_bs[3], coarse.Vh, nvec, Nf / 2, XL, parity, ips);
    else restrict_rows_kernel<6, float4><<<nblk, dim3(32, nvec / 2), 0, rt().compute>>>((float4 *)co.v, (const float4 *)fi.v, (const float4 *)V, f2c, fine, b0,
                                                                                       geo_bs[1], geo_bs[2], geo_bs[3], coarse.Vh, nvec, Nf / 2, XL, parity, ips);
  } else if (Nf == 12) RK(6);
  else RK(0);
#undef RK
  QB_CHECK_LAUNCH();
  flops += 8ll * Nf * nvec * fine.Vh * (parity < 0 ? 2 : 1);
}

// n fine fields fo[i] (+)= P ci[i]: V is streamed once per group of up to TRANSFER_MAX_NR vectors
void Transfer::P_multi(SpinorField *const *fo, const SpinorField *const *ci, int n, bool accumulate, int parity) const {
  const long nt = (parity < 0 ? 2 : 1) * fine.Vh * (Nf / 2);
  const long ops = fine_pstride(*fo[0], Nf);
  for (int first = 0; first < n; first += TRANSFER_MAX_NR) {
    const int nr = std::min(TRANSFER_MAX_NR, n - first);
    MultiPtrs p{};
    for (int r = 0; r < nr; r++) {
      SpinorField &f = *fo[first + r];
      const SpinorField &c = *ci[first + r];
      if (f.prec != PREC_SINGLE || c.prec != PREC_SINGLE || (parity < 0 && f.nparity != 2) || f.nparity != fo[0]->nparity || c.nparity != 2 || f.Vh != fine.Vh ||
          c.Vh != coarse.Vh || f.ncomplex != Nf || c.ncomplex != 2 * nvec)
        QB_ERROR("Transfer::P_multi: field geometry mismatch");
      p.out[r] = (float4 *)f.v; p.in[r] = (const float4 *)c.v;
    }
#define PMV(NR_, ACC_, VT_, VP_) prolong_multi_kernel<NR_, ACC_, VT_><<<div_up(nt, 128), 128, 0, rt().compute>>>(p, (const VT_ *)VP_, f2c, fine.Vh, coarse.Vh, Nf, nvec, Nf / 2, parity, ops)
#define PM(NR_) \
    if (V16) { if (accumulate) PMV(NR_, true, uint2, V16); else PMV(NR_, false, uint2, V16); } \
    else { if (accumulate) PMV(NR_, true, float4, V); else PMV(NR_, false, float4, V); }
    switch (nr) {
      case 1: PM(1); break;
      case 2: PM(2); break;
      case 3: PM(3); break;
      case 4: PM(4); break;
      case 5: PM(5); break;
      default: PM(6); break;
    }
#undef PM
#undef PMV
    QB_CHECK_LAUNCH();
    flops += 8ll * Nf * nvec * fine.Vh * (parity < 0 ? 2 : 1) * nr;
  }
}

// n coarse fields co[i] = R fi[i]
void Transfer::R_multi(SpinorField *const *co, const SpinorField *const *fi, int n, int parity) const {
  const int b0 = geo_bs[0];
  int XL = 0;
  if (fine.Xh == 8 || fine.Xh == 16 || fine.Xh % 32 == 0) XL = fine.Xh < 32 ? fine.Xh : 32;
  else if (fine.Xh % 8 == 0) XL = 8;
  const bool rows = Nf == 12 && XL && (b0 == 2 || b0 == 4 || b0 == 8) && geo_bs[1] % (32 / XL) == 0 && 32 * (nvec / 2) <= 384;
  if (!rows) {  // geometries the row-major kernel does not take: one vector at a time
    for (int i = 0; i < n; i++) R(*co[i], *fi[i], parity);
    return;
  }
  const long ips = fine_pstride(*fi[0], Nf);
  const unsigned nblk = (unsigned)((fine.Xh / XL) * (fine.X[1] / geo_bs[1]) * (fine.X[2] / geo_bs[2]) * (fine.X[3] / geo_bs[3]));
  for (int first = 0; first < n; first += TRANSFER_MAX_NR) {
    const int nr = std::min(TRANSFER_MAX_NR, n - first);
    MultiPtrs p{};
    for (int r = 0; r < nr; r++) {
      SpinorField &c = *co[first + r];
      const SpinorField &f = *fi[first + r];
      if (f.prec != PREC_SINGLE || c.prec != PREC_SINGLE || (parity < 0 && f.nparity != 2) || f.nparity != fi[0]->nparity || c.nparity != 2 || f.Vh != fine.Vh ||
          c.Vh != coarse.Vh || f.ncomplex != Nf || c.ncomplex != 2 * nvec)
        QB_ERROR("Transfer::R_multi: field geometry mismatch");
      p.out[r] = (float4 *)c.v; p.in[r] = (const float4 *)f.v;
    }
#define RM(NR_) \
    if (V16) restrict_rows_multi_kernel<6, NR_, uint2><<<nblk, dim3(32, nvec / 2), 0, rt().compute>>>(p, (const uint2 *)V16, f2c, fine, b0, geo_bs[1], geo_bs[2], geo_bs[3], coarse.Vh, nvec, Nf / 2, XL, parity, ips); \
    else restrict_rows_multi_kernel<6, NR_, float4><<<nblk, dim3(32, nvec / 2), 0, rt().compute>>>(p, (const float4 *)V, f2c, fine, b0, geo_bs[1], geo_bs[2], geo_bs[3], coarse.Vh, nvec, Nf / 2, XL, parity, ips)
    switch (nr) {
      case 1: RM(1); break;
      case 2: RM(2); break;
      case 3: RM(3); break;
      case 4: RM(4); break;
      case 5: RM(5); break;
      default: RM(6); break;
    }
#undef RM
    QB_CHECK_LAUNCH();
    flops += 8ll * Nf * nvec * fine.Vh * (parity < 0 ? 2 : 1) * nr;
  }
}

// fp16 copy of V for the prolongator / restrictor (both are bound by streaming V); the coarse-link build, the block orthogonalisation
// and the fallback aggregate-major restrictor keep the fp32 V
void Transfer::enable_half_v() {
  const size_t n = v_bytes() / 16;
  if (!V16) V16 = pool_malloc(n * sizeof(uint2));
  v_to_half_kernel<<<(unsigned)div_up((long)n, 256), 256, 0, rt().compute>>>((uint2 *)V16, (const float4 *)V, n);
  QB_CHECK_LAUNCH();
}

}  // namespace qb
