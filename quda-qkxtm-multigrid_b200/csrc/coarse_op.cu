// Galerkin construction of the coarse links on the GPU:  L^c_d(X) = sum_{x in X} V(x)^dag L_d(x) V(x + e_d).
//
// Replaces the reference's CPU-only path (/root/reference/lib/coarse_op.cu:152-213, lib/coarsecoarse_op.cu:148-184,
// lib/coarse_op.cuh: computeUV :59-125, computeVUV :487-599, computeCoarseLocal :670-711, AddCoarseDiagonal /
// AddCoarseTmDiagonal :813-869; `errorQuda("GPU variant not yet implemented")` :1117-1119).  The reference
// materialises UV = U V for the whole lattice and accumulates V^dag UV site by site on one CPU core.
// Here one CTA owns one aggregate X and walks its fine sites; for every site and direction it
//   1. stages V(x), V(x + e_d) (and the link) in shared memory,
//   2. forms W = L_d(x) V(x+e_d) split by source chirality (the reference's UV, never written to HBM),
//   3. accumulates V(x)^dag W into register tiles of the N x N coarse matrix: into Y_d(X) if x + e_d leaves
//      the aggregate, into the site-diagonal block L_8(X) otherwise (computeCoarseLocal).
// The mass / twist term enters as a ninth "direction" with e_8 = 0 (AddCoarseDiagonal / AddCoarseTmDiagonal).
// All sums run in a fixed order: the result is deterministic.
#include "coarse.h"
#include "dslash.cuh"

namespace qb {

// ---- fp32 AoS copy of the fine links, any resident precision / reconstruction -----------------------
template <typename Store, int RECON>
__global__ void decompress_gauge_kernel(float *out, const void *src, Geom g, long Vh) {
  typedef typename Store::real real;
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 8 * Vh) return;
  const int pm = (int)(t / Vh);
  const long cb = t - (long)pm * Vh;
  const int mu = pm & 3;
  real raw[RECON];
  LinkRaw<Store, RECON>::load(raw, (const char *)src + (size_t)pm * RECON * StoreTraits<Store>::real_bytes * Vh, Vh, cb);
  real u0;
  if (mu < 3) u0 = RECON == 8 ? (real)1 / (real)g.aniso : (real)g.aniso;
  else {
    const long za = cb / g.Xh, zb = za / g.X[1];
    const int tt = (int)(zb / g.X[2]);
    u0 = (tt == g.X[3] - 1) ? (real)g.tb_fwd : (real)1;
  }
  cplx<real> U[9];
  reconstruct_link<real, RECON>(U, raw, link_u0<Store, RECON>(u0));
  const real ls = link_scale<Store, RECON>();
  float *dst = out + (size_t)t * 18;  // [parity][mu][cb][row][col][re,im]
#pragma unroll
  for (int k = 0; k < 9; k++) { dst[2 * k] = (float)(U[k].re * ls); dst[2 * k + 1] = (float)(U[k].im * ls); }
}

float *decompress_gauge(const GaugeField &gf, const Geom &g) {
  float *out;
  out = (float *)pool_malloc(sizeof(float) * 18 * 8 * gf.Vh);
  const int bs = 256, nb = div_up(8 * gf.Vh, bs);
  cudaStream_t s = rt().compute;
#define DC(ST, RC) decompress_gauge_kernel<ST, RC><<<nb, bs, 0, s>>>(out, gf.data, g, gf.Vh)
#define BY_RECON(ST)                  \
  if (gf.recon == 18) DC(ST, 18);     \
  else if (gf.recon == 12) DC(ST, 12);\
  else DC(ST, 8)
  if (gf.prec == PREC_DOUBLE) { BY_RECON(StoreD); }
  else if (gf.prec == PREC_SINGLE) { BY_RECON(StoreS); }
  else { BY_RECON(StoreH); }
#undef BY_RECON
#undef DC
  QB_CHECK_LAUNCH();
  return out;
}

// ghost links (compressed [parity][plane][faceVh]) -> [parity][faceVh][18]; they sit on the backward neighbour's last slice
template <typename Store, int RECON>
__global__ void decompress_ghost_kernel(float *out, const void *src, Geom g, int mu, int faceVh) {
  typedef typename Store::real real;
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2L * faceVh) return;
  const int parity = (int)(t / faceVh);
  const long f = t - (long)parity * faceVh;
  real raw[RECON];
  LinkRaw<Store, RECON>::load(raw, (const char *)src + (size_t)parity * RECON * StoreTraits<Store>::real_bytes * faceVh, faceVh, f);
  real u0;
  if (mu < 3) u0 = RECON == 8 ? (real)1 / (real)g.aniso : (real)g.aniso;
  else u0 = (real)g.tb_bwd;  // links used by the backward hop of sites at t = 0
  cplx<real> U[9];
  reconstruct_link<real, RECON>(U, raw, link_u0<Store, RECON>(u0));
  const real ls = link_scale<Store, RECON>();
  float *dst = out + (size_t)t * 18;
#pragma unroll
  for (int k = 0; k < 9; k++) { dst[2 * k] = (float)(U[k].re * ls); dst[2 * k + 1] = (float)(U[k].im * ls); }
}

static float *decompress_ghost(const GaugeField &gf, const Geom &g, int mu) {
  float *out;
  const int fv = g.faceVh[mu];
  QB_CUDA(cudaMalloc((void **)&out, sizeof(float) * 18 * 2 * fv));
  const int bs = 256, nb = div_up(2L * fv, bs);
  cudaStream_t s = rt().compute;
#define DG(ST, RC) decompress_ghost_kernel<ST, RC><<<nb, bs, 0, s>>>(out, gf.ghost[mu], g, mu, fv)
#define BY_RECON(ST)                  \
  if (gf.recon == 18) DG(ST, 18);     \
  else if (gf.recon == 12) DG(ST, 12);\
  else DG(ST, 8)
  if (gf.prec == PREC_DOUBLE) { BY_RECON(StoreD); }
  else if (gf.prec == PREC_SINGLE) { BY_RECON(StoreS); }
  else { BY_RECON(StoreH); }
#undef BY_RECON
#undef DG
  QB_CHECK_LAUNCH();
  return out;
}

// ---- shared pieces -------------------------------------------------------------------------------------
struct GalerkinArgs {
  // transfer
  const float4 *V;
  const float4 *VL;   // left vectors of the Galerkin product (same layout as V); nullptr: V itself.  Preconditioned coarsening with a clover
                      // term: VL = A^-dag V, so that VL^dag L V = V^dag A^-1 L V (the reference's AV, lib/coarse_op.cuh:384-456, on the other side)
  const int *f2c, *c2f;
  int Nf, nvec, block_sites;
  long Vh_f;
  int Xf[4], Xfh;
  // output
  float4 *Yc;  // [Vc][9][N][N/2]
  int N;
  // fine-level links
  const float *U;  // decompressed [parity][mu][cb][18]
  float kappa, twist_a;
  const float *clover;  // site-major packed clover term [Vf][72], or nullptr
  // coarse-level links (the finer coarse operator)
  const float4 *Yf;  // [Vf][9][Nf][Nf/2]
  // partitioned dimensions: V (and, on the fine level, backward links) of the neighbours' boundary slices
  int part[4], faceVh[4];
  const float4 *Vghost[4][2];  // [d][0] from the backward, [d][1] from the forward neighbour: [parity][k][nvec/2][faceVh]
  const float *Ughost[4];      // decompressed U_d at x_d = X_d - 1 of the backward neighbour: [parity][faceVh][18]
};

__device__ __forceinline__ void fine_coords(int *x, long cb, int parity, const GalerkinArgs &a) {
  const long za = cb / a.Xfh, zb = za / a.Xf[1];
  x[1] = (int)(za - zb * a.Xf[1]);
  x[3] = (int)(zb / a.Xf[2]);
  x[2] = (int)(zb - (long)x[3] * a.Xf[2]);
  x[0] = (int)(2 * cb + ((x[1] + x[2] + x[3] + parity) & 1) - za * a.Xf[0]);
}

// V(x) -> smem as [k][j] complex; base/stride select the local field or a ghost slice
__device__ __forceinline__ void stage_V(float2 *dst, const GalerkinArgs &a, const float4 *base, long stride, int parity, long idx) {
  const int nvh = a.nvec / 2;
  for (int e = threadIdx.x; e < a.Nf * nvh; e += blockDim.x) {
    const int k = e / nvh, jp = e - k * nvh;
    const float4 v = __ldg(base + (((size_t)parity * a.Nf + k) * nvh + jp) * stride + idx);
    dst[k * a.nvec + 2 * jp] = make_float2(v.x, v.y);
    dst[k * a.nvec + 2 * jp + 1] = make_float2(v.z, v.w);
  }
}

// acc[t][u] += sum_{k in chirality S of the tile rows} conj(Vx[k][j0+t]) * W[k][c0+u]
template <int T>
__device__ __forceinline__ void accumulate_tile(cplx<float> (*acc)[T], const float2 *Vx, const float2 *W, int nvec, int N, int cpc, int r0, int c0) {
  const int S = r0 / nvec, j0 = r0 - S * nvec;
  for (int kk = 0; kk < cpc; kk++) {
    const int k = S * cpc + kk;
    cplx<float> v[T], w[T];
#pragma unroll
    for (int t = 0; t < T; t++) { const float2 q = Vx[k * nvec + j0 + t]; v[t] = cplx<float>(q.x, q.y); }
#pragma unroll
    for (int u = 0; u < T; u++) { const float2 q = W[k * N + c0 + u]; w[u] = cplx<float>(q.x, q.y); }
#pragma unroll
    for (int t = 0; t < T; t++)
#pragma unroll
      for (int u = 0; u < T; u++) cmac_conj(acc[t][u], v[t], w[u]);
  }
}

template <int T>
__device__ __forceinline__ void store_tile(float4 *Yc, long X, int d, int N, int r0, int c0, cplx<float> (*acc)[T]) {
  float2 *base = (float2 *)(Yc + ((size_t)X * 9 + d) * N * (N / 2));
#pragma unroll
  for (int t = 0; t < T; t++)
#pragma unroll
    for (int u = 0; u < T; u++) {
      const int r = r0 + t, c = c0 + u;
      base[((size_t)c * (N / 2) + (r >> 1)) * 2 + (r & 1)] = make_float2(acc[t][u].re, acc[t][u].im);
    }
}

// LEVEL 0: fine links are -kappa P_d (x) U; LEVEL 1: dense links of a coarse operator
template <int T, int LEVEL>
__global__ void __launch_bounds__(256) galerkin_kernel(const GalerkinArgs a) {
  extern __shared__ float2 sm[];
  const int Nf = a.Nf, nvec = a.nvec, N = a.N, cpc = Nf / 2;
  float2 *Vx = sm;                    // [Nf][nvec]
  float2 *Vn = Vx + Nf * nvec;        // [Nf][nvec]
  float2 *W = Vn + Nf * nvec;         // [Nf][N]
  float2 *Ld = W + Nf * N;            // LEVEL 0: 9 complex (the link); LEVEL 1: [Nf][Nf]
  float2 *VxL = a.VL ? Ld + (LEVEL == 0 ? 16 : Nf * Nf) : Vx;   // [Nf][nvec] left vectors at x
  __shared__ int s_info[4];           // neighbour parity, leaves-block flag, ghost flag
  __shared__ long s_cb[2];

  const long X = blockIdx.x;
  const int ntc = N / T;
  const int tr = threadIdx.x / ntc, tc = threadIdx.x - tr * ntc;
  const bool owner = threadIdx.x < ntc * ntc;
  const int r0 = tr * T, c0 = tc * T;

  cplx<float> diag[T][T];
#pragma unroll
  for (int t = 0; t < T; t++)
#pragma unroll
    for (int u = 0; u < T; u++) diag[t][u] = cplx<float>(0.f, 0.f);

  for (int d = 0; d < 9; d++) {
    cplx<float> hop[T][T];
#pragma unroll
    for (int t = 0; t < T; t++)
#pragma unroll
      for (int u = 0; u < T; u++) hop[t][u] = cplx<float>(0.f, 0.f);

    for (int i = 0; i < a.block_sites; i++) {
      const int fs = a.c2f[(size_t)X * a.block_sites + i];
      const int parity = fs >= a.Vh_f ? 1 : 0;
      const long cb = fs - (long)parity * a.Vh_f;
      __syncthreads();  // previous iteration's readers are done with the staging buffers
      if (threadIdx.x == 0) {
        int x[4];
        fine_coords(x, cb, parity, a);
        int npar = parity, ghost = 0;
        long ncb = cb;
        if (d < 8) {
          const int mu = d >> 1;
          const bool edge = (d & 1) ? (x[mu] == 0) : (x[mu] == a.Xf[mu] - 1);
          npar = 1 - parity;
          if (edge && a.part[mu]) {
            // neighbour lives on another rank: index of the site inside the face (3-d lexicographic >> 1)
            ghost = 1;
            const int d0 = mu == 0 ? 1 : 0, d1 = mu <= 1 ? 2 : 1, d2 = mu <= 2 ? 3 : 2;
            ncb = (x[d0] + a.Xf[d0] * (x[d1] + (long)a.Xf[d1] * x[d2])) >> 1;
          } else {
            x[mu] = (x[mu] + ((d & 1) ? a.Xf[mu] - 1 : 1)) % a.Xf[mu];
            ncb = ((((long)x[3] * a.Xf[2] + x[2]) * a.Xf[1] + x[1]) * a.Xf[0] + x[0]) >> 1;
          }
        }
        s_info[0] = npar;
        s_cb[0] = ncb;
        s_info[2] = ghost;
        s_info[1] = (d < 8 && (ghost || a.f2c[(size_t)npar * a.Vh_f + ncb] != X)) ? 1 : 0;
      }
      __syncthreads();
      const int npar = s_info[0];
      const long ncb = s_cb[0];
      const bool ghost = s_info[2] != 0;
      stage_V(Vx, a, a.V, a.Vh_f, parity, cb);
      if (a.VL) stage_V(VxL, a, a.VL, a.Vh_f, parity, cb);
      if (d < 8) {
        if (ghost) stage_V(Vn, a, a.Vghost[d >> 1][(d & 1) ? 0 : 1], a.faceVh[d >> 1], npar, ncb);
        else stage_V(Vn, a, a.V, a.Vh_f, npar, ncb);
      }
      if (LEVEL == 0) {
        if (d < 8 && threadIdx.x < 9) {
          // forward: U_mu(x); backward: U_mu(x - mu)^dagger
          const int mu = d >> 1;
          const float *u = (d & 1) ? (ghost ? a.Ughost[mu] + ((size_t)npar * a.faceVh[mu] + ncb) * 18 : a.U + (((size_t)npar * 4 + mu) * a.Vh_f + ncb) * 18)
                                   : a.U + (((size_t)parity * 4 + mu) * a.Vh_f + cb) * 18;
          const int r = threadIdx.x / 3, c = threadIdx.x - 3 * r;
          Ld[threadIdx.x] = (d & 1) ? make_float2(u[(c * 3 + r) * 2], -u[(c * 3 + r) * 2 + 1]) : make_float2(u[threadIdx.x * 2], u[threadIdx.x * 2 + 1]);
        }
      } else {
        const float4 *src = a.Yf + ((size_t)fs * 9 + d) * Nf * (Nf / 2);
        for (int e = threadIdx.x; e < Nf * (Nf / 2); e += blockDim.x) {
          const int c = e / (Nf / 2), rp = e - c * (Nf / 2);
          const float4 v = __ldg(src + e);
          Ld[(2 * rp) * Nf + c] = make_float2(v.x, v.y);
          Ld[(2 * rp + 1) * Nf + c] = make_float2(v.z, v.w);
        }
      }
      __syncthreads();
      // W[k][(S', j')] = sum_{k' in S'} L_d(x)[k][k'] V(x + e_d)[k'][j']
      const float2 *Vsrc = d < 8 ? Vn : Vx;
      for (int e = threadIdx.x; e < Nf * N; e += blockDim.x) {
        const int k = e / N, col = e - k * N;
        const int Sp = col / nvec, jp = col - Sp * nvec;
        cplx<float> w(0.f, 0.f);
        if (LEVEL == 0) {
          const int s = k / 3, c = k - 3 * s;
          const int chi = s >> 1;
          if (d == 8) {
            if (Sp == chi) {
              const float2 v = Vx[k * nvec + jp];
              const float tw = chi == 0 ? a.twist_a : -a.twist_a;  // (1 + i a gamma5)  or  (C + i a gamma5) with a clover term
              if (!a.clover) w = cplx<float>(v.x - tw * v.y, v.y + tw * v.x);
              else {
                const float *cb = a.clover + ((size_t)fs * 2 + chi) * 36;  // Hermitian 6 x 6 block: 6 diagonal reals, 15 complex L(row > col) by columns
                const int r = k - 6 * chi;
                w = cplx<float>(cb[r] * v.x - tw * v.y, cb[r] * v.y + tw * v.x);
                for (int c2 = 0; c2 < 6; c2++) {
                  if (c2 == r) continue;
                  const float2 u = Vx[(6 * chi + c2) * nvec + jp];
                  const int lo = c2 < r ? c2 : r, hi = c2 < r ? r : c2;
                  const int kk = 15 - (6 - lo) * (5 - lo) / 2 + hi - lo - 1;
                  const cplx<float> l(cb[6 + 2 * kk], cb[6 + 2 * kk + 1]);
                  if (r > c2) cmac(w, l, cplx<float>(u.x, u.y));       // H[r][c2] = L
                  else cmac_conj(w, l, cplx<float>(u.x, u.y));          // H[r][c2] = conj(L(c2, r))
                }
              }
            }
          } else {
            const int mu = d >> 1;
            const float sigma = (d & 1) ? 1.f : -1.f;  // forward: 1 - gamma_mu, backward: 1 + gamma_mu
            int ss;
            cplx<float> coef;
            if (Sp == chi) { ss = s; coef = cplx<float>(1.f, 0.f); }
            else {
              ss = mu < 2 ? 3 - s : (s + 2) & 3;
              const int gr = mu == 1 ? ((s == 0 || s == 3) ? -1 : 1) : (mu == 3 ? 1 : 0);
              const int gi = mu == 0 ? (s < 2 ? 1 : -1) : (mu == 2 ? ((s == 0 || s == 3) ? 1 : -1) : 0);
              coef = cplx<float>(sigma * gr, sigma * gi);
            }
            cplx<float> acc(0.f, 0.f);
#pragma unroll
            for (int cc = 0; cc < 3; cc++) {
              const float2 u = Ld[c * 3 + cc], v = Vsrc[(ss * 3 + cc) * nvec + jp];
              cmac(acc, cplx<float>(u.x, u.y), cplx<float>(v.x, v.y));
            }
            w = coef * acc;
            w.re *= -a.kappa; w.im *= -a.kappa;
          }
        } else {
          const int cpcf = Nf / 2;
          for (int kk = 0; kk < cpcf; kk++) {
            const int kp = Sp * cpcf + kk;
            const float2 l = Ld[k * Nf + kp], v = Vsrc[kp * nvec + jp];
            cmac(w, cplx<float>(l.x, l.y), cplx<float>(v.x, v.y));
          }
        }
        W[e] = make_float2(w.re, w.im);
      }
      __syncthreads();
      if (owner) {
        if (s_info[1]) accumulate_tile<T>(hop, VxL, W, nvec, N, cpc, r0, c0);
        else accumulate_tile<T>(diag, VxL, W, nvec, N, cpc, r0, c0);
      }
    }
    if (owner && d < 8) store_tile<T>(a.Yc, X, d, N, r0, c0, hop);
  }
  if (owner) store_tile<T>(a.Yc, X, 8, N, r0, c0, diag);
}

static int tile_for(int N, int nvec) {
  // largest T in {4,3,2,1} dividing nvec with (N/T)^2 <= 256 threads
  for (int T = 1; T <= 4; T++)
    if (nvec % T == 0 && (N / T) * (N / T) <= 256) return T;
  QB_ERROR("no register tiling for the coarse-link build with n_vec = %d", nvec);
}

template <int LEVEL> static void launch_galerkin(const GalerkinArgs &a, long Vc) {
  const int T = tile_for(a.N, a.nvec);
  const size_t sm = sizeof(float2) * ((size_t)(a.VL ? 3 : 2) * a.Nf * a.nvec + (size_t)a.Nf * a.N + (LEVEL == 0 ? 16 : (size_t)a.Nf * a.Nf));
  cudaStream_t s = rt().compute;
  const int threads = std::max(32, (((a.N / T) * (a.N / T) + 31) / 32) * 32);
#define GO(TT)                                                                                               \
  QB_CUDA(cudaFuncSetAttribute(galerkin_kernel<TT, LEVEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
  galerkin_kernel<TT, LEVEL><<<(unsigned)Vc, threads, sm, s>>>(a)
  if (sm > 227 * 1024) QB_ERROR("coarse-link build needs %zu bytes of shared memory (Nf=%d, n_vec=%d): too large", sm, a.Nf, a.nvec);
  if (T == 4) { GO(4); }
  else if (T == 3) { GO(3); }
  else if (T == 2) { GO(2); }
  else { GO(1); }
#undef GO
  QB_CHECK_LAUNCH();
}

static void fill_transfer_args(GalerkinArgs &a, const Transfer &T) {
  for (int d = 0; d < 4; d++) {
    a.part[d] = T.fine.part[d]; a.faceVh[d] = T.fine.faceVh[d];
    a.Vghost[d][0] = (const float4 *)T.Vghost[d][0]; a.Vghost[d][1] = (const float4 *)T.Vghost[d][1];
    a.Ughost[d] = nullptr;
    if (a.part[d] && (!T.Vghost[d][0] || !T.Vghost[d][1])) QB_ERROR("transfer operator has no ghost V for partitioned dimension %d", d);
  }
  a.V = (const float4 *)T.V; a.f2c = T.f2c; a.c2f = T.c2f;
  a.Nf = T.Nf; a.nvec = T.nvec; a.block_sites = T.block_sites; a.Vh_f = T.fine.Vh;
  for (int d = 0; d < 4; d++) a.Xf[d] = T.fine.X[d];
  a.Xfh = T.fine.Xh;
  a.N = 2 * T.nvec;
}

float *decompress_ghost_links(const GaugeField &gf, const Geom &g, int mu) { return decompress_ghost(gf, g, mu); }

void build_coarse_from_fine(CoarseOperator &out, const Transfer &T, const GaugeField &gauge, const Geom &fine_geom, double kappa, double twist_a,
                            const float *clover_site, const float *VL) {
  if (T.Nf != 12) QB_ERROR("build_coarse_from_fine: transfer is not defined on a Wilson-type fine field");
  out.allocate(T.coarse, T.nvec);
  if (!VL && galerkin_mma_supported(T)) {  // tensor-core build (coarse_op_mma.cu); separate left vectors take the CUDA-core kernel
    build_coarse_from_fine_mma(out, T, gauge, fine_geom, kappa, twist_a, clover_site);
    return;
  }
  float *U = decompress_gauge(gauge, fine_geom);
  GalerkinArgs a{};
  fill_transfer_args(a, T);
  a.Yc = (float4 *)out.Y; a.U = U; a.kappa = (float)kappa; a.twist_a = (float)twist_a; a.Yf = nullptr; a.clover = clover_site;
  a.VL = (const float4 *)VL;
  float *ug[4] = {nullptr, nullptr, nullptr, nullptr};
  for (int d = 0; d < 4; d++)
    if (fine_geom.part[d]) {
      if (!gauge.ghost[d]) QB_ERROR("gauge field has no ghost links for partitioned dimension %d", d);
      ug[d] = decompress_ghost(gauge, fine_geom, d);
      a.Ughost[d] = ug[d];
    }
  launch_galerkin<0>(a, T.coarse.V());
  QB_CUDA(cudaStreamSynchronize(rt().compute));
  pool_free(U);
  for (int d = 0; d < 4; d++)
    if (ug[d]) QB_CUDA(cudaFree(ug[d]));
}

void build_coarse_from_coarse(CoarseOperator &out, const Transfer &T, const CoarseOperator &fine, bool preconditioned) {
  if (T.Nf != fine.N || T.fine.Vh != fine.geom.Vh) QB_ERROR("build_coarse_from_coarse: transfer does not match the fine coarse-operator");
  if (preconditioned && !fine.Yhat) QB_ERROR("build_coarse_from_coarse: preconditioned coarsening needs the Yhat links of the fine level");
  out.allocate(T.coarse, T.nvec);
  GalerkinArgs a{};
  fill_transfer_args(a, T);
  // preconditioned: the stencil 1 + sum_d Yhat_d of Xinv M (slot 8 of Yhat holds the identity), same kernel
  a.Yc = (float4 *)out.Y; a.U = nullptr; a.Yf = (const float4 *)(preconditioned ? fine.Yhat : fine.Y);
  launch_galerkin<1>(a, T.coarse.V());
  QB_CUDA(cudaStreamSynchronize(rt().compute));
}

// rows of chirality S (r / nvec) of all nine matrices of a site times c[S]
__global__ void scale_rows_kernel(float2 *Y, size_t n, int N, int nvec, float2 c0, float2 c1) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = (int)(((i >> 1) % (N / 2)) * 2 + (i & 1));   // element index = ((.. * N + col) * N/2 + rp) * 2 + (r & 1)
  const float2 c = r < nvec ? c0 : c1, v = Y[i];
  Y[i] = make_float2(c.x * v.x - c.y * v.y, c.x * v.y + c.y * v.x);
}

void scale_coarse_rows(CoarseOperator &op, std::complex<double> c0, std::complex<double> c1) {
  const size_t n = (size_t)op.geom.V() * 9 * op.N * op.N;
  scale_rows_kernel<<<(unsigned)div_up((long)n, 256), 256, 0, rt().compute>>>((float2 *)op.Y, n, op.N, op.nvec, make_float2((float)c0.real(), (float)c0.imag()),
                                                                            make_float2((float)c1.real(), (float)c1.imag()));
  QB_CHECK_LAUNCH();
}

// `this` is the full operator of the level.  preconditioned: the coarse links of A^-1 M = 1 - kappa A^-1 D, the operator whose even-odd
// Schur complement is the symmetric preconditioned one (the reference gets there through AV = A^-1 V and bidirectional links,
// lib/coarse_op.cuh:202-224, :1349-1440, called from DiracTwistedMassPC::createCoarseOp, lib/dirac_twisted_mass.cpp:580).  For twisted
// mass A = 1 + i a gamma5 is a constant per chirality and V is chirality-blocked, so  V^dag A^-1 M V = A_c^-1 (V^dag M V):  the rows of
// chirality +- of the ordinary Galerkin links get the factor 1 / (1 +- i a) -- exact, and the tensor-core build is reused as it is.
// column j of V as a fine field and back: f(x, k) = V(x, k, j)   (fill_v_kernel of transfer.cu is the other direction)
__global__ void v_column_kernel(float4 *field, const float4 *V, long Vh, int Nf, int nvec, int j, int to_v, float4 *Vout) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nkp = Nf / 2, nvh = nvec / 2;
  if (t >= 2 * Vh * nkp) return;
  const long cb = t % Vh;
  const int kp = (int)((t / Vh) % nkp), parity = (int)(t / (Vh * nkp));
  const size_t i0 = (((size_t)parity * Nf + 2 * kp) * nvh + j / 2) * Vh + cb, i1 = (((size_t)parity * Nf + 2 * kp + 1) * nvh + j / 2) * Vh + cb;
  float4 *f = field + ((size_t)parity * nkp + kp) * Vh + cb;
  if (!to_v) {
    const float2 a = ((const float2 *)(V + i0))[j & 1], b = ((const float2 *)(V + i1))[j & 1];
    *f = make_float4(a.x, a.y, b.x, b.y);
  } else {
    const float4 v = *f;
    ((float2 *)(Vout + i0))[j & 1] = make_float2(v.x, v.y);
    ((float2 *)(Vout + i1))[j & 1] = make_float2(v.z, v.w);
  }
}

// `this` is the full operator of the level.  preconditioned: the coarse links of A^-1 M = 1 - kappa A^-1 D, the operator whose even-odd
// Schur complement is the symmetric preconditioned one (the reference gets there through AV = A^-1 V and bidirectional links,
// lib/coarse_op.cuh:202-224, :1349-1440, called from DiracTwistedMassPC::createCoarseOp, lib/dirac_twisted_mass.cpp:580).  For twisted
// mass A = 1 + i a gamma5 is a constant per chirality and V is chirality-blocked, so  V^dag A^-1 M V = A_c^-1 (V^dag M V):  the rows of
// chirality +- of the ordinary Galerkin links get the factor 1 / (1 +- i a) -- exact, and the tensor-core build is reused as it is.
// With a clover term (DiracTwistedCloverPC::createCoarseOp, computeTMCAV lib/coarse_op.cuh:384-456) A = C + i a gamma5 varies from site
// to site: the product is taken with the left vectors  VL = A^-dag V  (column by column through the clover kernel), VL^dag L V = V^dag A^-1 L V.
void DiracTM::create_coarse_op(CoarseOperator &coarse, const Transfer &T, bool preconditioned) const {
  if (pc) QB_ERROR("create_coarse_op is called on the full operator of the level (preconditioned = true selects the coarsening of A^-1 M)");
  if (dagger) QB_ERROR("create_coarse_op: operator must not be daggered");
  float *cs = clover ? clover->site_major_f32() : nullptr;
  float *VL = nullptr;
  if (preconditioned && clover) {
    if (gauge->prec != PREC_SINGLE) QB_ERROR("preconditioned coarsening with a clover term needs the fp32 operator");
    VL = (float *)pool_malloc(T.v_bytes());
    SpinorField f(lat->geom.Vh, 2, PREC_SINGLE), gfield(lat->geom.Vh, 2, PREC_SINGLE);
    const long nt = 2 * (long)lat->geom.Vh * 6;
    const CloverField &cl = clover->get(PREC_SINGLE, twist_a());
    for (int j = 0; j < T.nvec; j++) {
      v_column_kernel<<<div_up(nt, 256), 256, 0, rt().compute>>>((float4 *)f.v, (const float4 *)T.V, lat->geom.Vh, 12, T.nvec, j, 0, nullptr);
      QB_CHECK_LAUNCH();
      for (int p = 0; p < 2; p++) {
        SpinorField o, i;
        gfield.view_parity(o, p); f.view_parity(i, p);
        clover_apply(o, i, cl, p, CLOVER_INVERSE_ADJ, twist_a(), nullptr, 1.0);
      }
      v_column_kernel<<<div_up(nt, 256), 256, 0, rt().compute>>>((float4 *)gfield.v, nullptr, lat->geom.Vh, 12, T.nvec, j, 1, (float4 *)VL);
      QB_CHECK_LAUNCH();
    }
  }
  build_coarse_from_fine(coarse, T, *gauge, lat->geom, kappa, flavor ? twist_a() : 0.0, cs, VL);
  if (cs) pool_free(cs);
  if (VL) pool_free(VL);
  if (preconditioned && flavor && !clover) {
    const std::complex<double> one(1.0, 0.0), ia(0.0, twist_a());
    scale_coarse_rows(coarse, one / (one + ia), one / (one - ia));
  }
}

}  // namespace qb
