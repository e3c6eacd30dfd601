// Fine-grid Wilson / twisted-mass hopping kernels for sm_100a.
//
// One kernel covers every operator variant of the reference's twisted-mass Dslash
// (/root/reference/lib/dslash_twisted_mass.cu:167-242, lib/tm_dslash_def.h:389-456, generated body
// lib/dslash_core/tm_dslash_gt200_core.h) and the plain Wilson Dslash(+xpay)
// (lib/dslash_wilson.cu:107) through three complex coefficient pairs:
//
//     out(x) = Cx * X(x)  +  Co * sum_mu [ P(-s)_mu U_mu(x) Cin*in(x+mu) + P(+s)_mu U_mu(x-mu)^dag Cin*in(x-mu) ]
//
// where a coefficient C = (p, q) acts as p + i q gamma5, i.e. (p + iq) on the upper two spins and
// (p - iq) on the lower two in the internal DeGrand-Rossi basis, and s = -1 (no dagger) / +1 (dagger).
//   QUDA_DEG_DSLASH_TWIST_INV   : Co = b(1, a)           (A^-1 D)
//   QUDA_DEG_TWIST_INV_DSLASH   : Cin = b(1, a)          (D A^-1; the pack kernel applies Cin too)
//   ..._XPAY of the two above   : Cx = (1,0), Co or Cin scaled by k
//   QUDA_DEG_DSLASH_TWIST_XPAY  : Cx = (1, a), Co = (b, 0)
// (coefficient table: SURVEY.md Appendix A.5, /root/reference/lib/dirac_twisted_mass.cpp:246-403)
//
// Thread mapping: one thread per output checkerboard site, consecutive threads = consecutive cb
// sites, so every plane load of a warp is one contiguous 512-byte (fp32) request.  Links use the
// L1-bypassing streaming path (read once per hop); neighbour spinors go through L1/L2 where the
// 8-fold reuse is caught.  Partitioned dimensions read projected half spinors from the ghost zone.
#pragma once
#include "layout.cuh"

namespace qb {

struct DslashParam {
  Geom g;
  int parity;  // parity of the output sites
  void *out; float *out_norm;
  const void *in; const float *in_norm;   // parity 1 - parity
  const void *x; const float *x_norm;     // same parity as out (may alias out)
  const void *gauge_fwd;                  // [mu][plane][Vh], parity of out
  const void *gauge_bwd;                  // [mu][plane][Vh], other parity
  const void *gauge_ghost[4];             // [plane][faceVh] links U_d(x - d) for sites at x_d = 0
  const void *ghost[4][2];                // received half spinors, [d][0]: from backward nbr, [d][1]: from forward nbr
  const float *ghost_norm[4][2];
  long stride;                            // plane stride of spinors and links (= Vh)
  double cin[2], co[2], cx[2];            // (p, q) pairs
  double sgn_fwd;                         // -1: (1 - gamma) on forward hops (no dagger), +1: dagger
  int nbatch;                             // batched fields: member = threadIdx.y, byte strides below (0 = not batched)
  long batch_in, batch_out, batch_x;
  int site_begin, site_count;             // contiguous range ...
  const int *site_list;                   // ... or explicit list of cb sites (interior / boundary split)
  // twisted clover: (C + i a gamma5)^-1 of the output parity as two full complex 6 x 6 blocks per site, [parity][36 planes][Vh] in the
  // arithmetic type (clover.h); applied to Co * (hop sum) before the x term.  nullptr: no clover term.
  const void *clover_inv;
  const float *clover_norm;               // int16 storage: one float per site, [parity][Vh]; clover_inv then holds 36 planes of 4 x int16
};

// ---- gamma matrices, DeGrand-Rossi: gamma_mu[s][gcol(mu,s)] = gre + i gim -----------------------
__host__ __device__ constexpr int gcol(int mu, int s) { return mu < 2 ? 3 - s : (s + 2) & 3; }
__host__ __device__ constexpr int gre(int mu, int s) {
  return mu == 1 ? ((s == 0 || s == 3) ? -1 : 1) : (mu == 3 ? 1 : 0);
}
__host__ __device__ constexpr int gim(int mu, int s) {
  return mu == 0 ? (s < 2 ? 1 : -1) : (mu == 2 ? ((s == 0 || s == 3) ? 1 : -1) : 0);
}

// z * (re + i im) for re, im in {0, +-1}, exactly one non-zero (folds at compile time)
template <typename T> __device__ __forceinline__ cplx<T> mul_unit(int re, int im, cplx<T> z) {
  if (re == 1) return z;
  if (re == -1) return cplx<T>(-z.re, -z.im);
  if (im == 1) return cplx<T>(-z.im, z.re);
  return cplx<T>(z.im, -z.re);
}

// (p + i q g5) psi in place on a full spinor
template <typename T> __device__ __forceinline__ void apply_twist(cplx<T> *psi, T p, T q) {
  const cplx<T> cu(p, q), cl(p, -q);
#pragma unroll
  for (int k = 0; k < 6; k++) psi[k] = cu * psi[k];
#pragma unroll
  for (int k = 6; k < 12; k++) psi[k] = cl * psi[k];
}

// upper two spin rows of (1 + sigma gamma_mu) psi
template <int MU, typename T> __device__ __forceinline__ void project(cplx<T> *h, const cplx<T> *psi, T sigma) {
#pragma unroll
  for (int s = 0; s < 2; s++)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const cplx<T> g = mul_unit(gre(MU, s), gim(MU, s), psi[gcol(MU, s) * 3 + c]);
      h[s * 3 + c] = cplx<T>(psi[s * 3 + c].re + sigma * g.re, psi[s * 3 + c].im + sigma * g.im);
    }
}

// acc += (1 + sigma gamma_mu)-reconstruction of the link-multiplied half spinor chi
template <bool SCALED, int MU, typename T> __device__ __forceinline__ void reconstruct_acc(cplx<T> *acc, const cplx<T> *chi, T sigma, T scale) {
  if (SCALED) sigma *= scale;
#pragma unroll
  for (int k = 0; k < 6; k++) {
    if (SCALED) { acc[k].re += scale * chi[k].re; acc[k].im += scale * chi[k].im; }
    else acc[k] = acc[k] + chi[k];
  }
#pragma unroll
  for (int s = 2; s < 4; s++)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const cplx<T> g = mul_unit(gre(MU, s), gim(MU, s), chi[gcol(MU, s) * 3 + c]);
      acc[s * 3 + c].re += sigma * g.re;
      acc[s * 3 + c].im += sigma * g.im;
    }
}

template <bool DAG, typename T> __device__ __forceinline__ void su3_mul(cplx<T> *chi, const cplx<T> *U, const cplx<T> *h) {
#pragma unroll
  for (int s = 0; s < 2; s++)
#pragma unroll
    for (int r = 0; r < 3; r++) {
      cplx<T> a((T)0, (T)0);
#pragma unroll
      for (int c = 0; c < 3; c++) {
        if (!DAG) cmac(a, U[r * 3 + c], h[s * 3 + c]);
        else cmac_conj(a, U[c * 3 + r], h[s * 3 + c]);
      }
      chi[s * 3 + r] = a;
    }
}

#ifndef QB_NO_F32X2
// ---- fp32 math on packed pairs (f32x2.cuh): same formulas as the generic templates above, one complex number per
// register pair, a complex multiply-accumulate in two FFMA2.  Used by the storage types with Store::packed_math
// (B200 sweep, profiles/README_r01.md: int16 storage 111 -> 101 us per hop; fp32 storage is latency-bound and does not gain).
__device__ __forceinline__ f2 cpk(const cplx<float> &z) { return pk2(z.re, z.im); }
__device__ __forceinline__ f2 cpk_swap(const cplx<float> &z) { return pk2(z.im, z.re); }
__device__ __forceinline__ cplx<float> cunpk(f2 v) { cplx<float> z; unpk2(v, z.re, z.im); return z; }

// a + sigma * (re + i im) * b   for the unit (re, im)
__device__ __forceinline__ f2 axpy_unit(int re, int im, float sigma, f2 a, const cplx<float> &b) {
  if (re == 1) return fma2(bc2(sigma), cpk(b), a);
  if (re == -1) return fma2(bc2(-sigma), cpk(b), a);
  if (im == 1) return fma2(pk2(-sigma, sigma), cpk_swap(b), a);  // i b = (-b.im, b.re)
  return fma2(pk2(sigma, -sigma), cpk_swap(b), a);
}

__device__ __forceinline__ void apply_twist_pk(cplx<float> *psi, float p, float q) {
  const f2 pp = bc2(p), qu = pk2(-q, q), ql = pk2(q, -q);
#pragma unroll
  for (int k = 0; k < 12; k++) psi[k] = cunpk(fma2(k < 6 ? qu : ql, cpk_swap(psi[k]), mul2(pp, cpk(psi[k]))));
}

template <int MU> __device__ __forceinline__ void project_pk(cplx<float> *h, const cplx<float> *psi, float sigma) {
#pragma unroll
  for (int s = 0; s < 2; s++)
#pragma unroll
    for (int c = 0; c < 3; c++)
      h[s * 3 + c] = cunpk(axpy_unit(gre(MU, s), gim(MU, s), sigma, cpk(psi[s * 3 + c]), psi[gcol(MU, s) * 3 + c]));
}

template <bool SCALED, int MU> __device__ __forceinline__ void reconstruct_acc_pk(cplx<float> *acc, const cplx<float> *chi, float sigma, float scale) {
  if (SCALED) sigma *= scale;
#pragma unroll
  for (int k = 0; k < 6; k++) acc[k] = cunpk(SCALED ? fma2(bc2(scale), cpk(chi[k]), cpk(acc[k])) : add2(cpk(acc[k]), cpk(chi[k])));
#pragma unroll
  for (int s = 2; s < 4; s++)
#pragma unroll
    for (int c = 0; c < 3; c++)
      acc[s * 3 + c] = cunpk(axpy_unit(gre(MU, s), gim(MU, s), sigma, cpk(acc[s * 3 + c]), chi[gcol(MU, s) * 3 + c]));
}

template <bool DAG> __device__ __forceinline__ void su3_mul_pk(cplx<float> *chi, const cplx<float> *U, const cplx<float> *h) {
#pragma unroll
  for (int s = 0; s < 2; s++) {
    f2 hp[3], hs[3];  // h and i*h (no dagger) / -i*h (dagger: conj(u) h = u.re h + u.im (-i h))
#pragma unroll
    for (int c = 0; c < 3; c++) {
      hp[c] = cpk(h[s * 3 + c]);
      hs[c] = mul2(cpk_swap(h[s * 3 + c]), DAG ? pk2(1.0f, -1.0f) : pk2(-1.0f, 1.0f));  // one packed op, no scalar negation + pair rebuild
    }
#pragma unroll
    for (int r = 0; r < 3; r++) {
      f2 a = mul2(bc2(U[DAG ? r : r * 3].re), hp[0]);
#pragma unroll
      for (int c = 1; c < 3; c++) a = fma2(bc2(U[DAG ? c * 3 + r : r * 3 + c].re), hp[c], a);
#pragma unroll
      for (int c = 0; c < 3; c++) a = fma2(bc2(U[DAG ? c * 3 + r : r * 3 + c].im), hs[c], a);
      chi[s * 3 + r] = cunpk(a);
    }
  }
}

// a += b * c
__device__ __forceinline__ void cmac_pk(cplx<float> &a, const cplx<float> &b, const cplx<float> &c) {
  a = cunpk(fma2(pk2(-b.im, b.im), cpk_swap(c), fma2(bc2(b.re), cpk(c), cpk(a))));
}
#endif
// double precision has no packed form: the *_pk names fall through to the generic templates
__device__ __forceinline__ void apply_twist_pk(cplx<double> *psi, double p, double q) { apply_twist(psi, p, q); }
template <int MU> __device__ __forceinline__ void project_pk(cplx<double> *h, const cplx<double> *psi, double s) { project<MU>(h, psi, s); }
template <bool SCALED, int MU> __device__ __forceinline__ void reconstruct_acc_pk(cplx<double> *a, const cplx<double> *c, double s, double sc) { reconstruct_acc<SCALED, MU>(a, c, s, sc); }
template <bool DAG> __device__ __forceinline__ void su3_mul_pk(cplx<double> *chi, const cplx<double> *U, const cplx<double> *h) { su3_mul<DAG>(chi, U, h); }
__device__ __forceinline__ void cmac_pk(cplx<double> &a, const cplx<double> &b, const cplx<double> &c) { cmac(a, b, c); }

// checkerboard index -> coordinates and the full lexicographic index
__device__ __forceinline__ void cb_coords(int *x, int &full, int cb, int parity, const Geom &g) {
  const int za = cb / g.Xh;
  const int zb = za / g.X[1];
  x[1] = za - zb * g.X[1];
  x[3] = zb / g.X[2];
  x[2] = zb - x[3] * g.X[2];
  const int odd = (x[1] + x[2] + x[3] + parity) & 1;
  full = 2 * cb + odd;
  x[0] = full - za * g.X[0];
}

// index of a site inside the checkerboarded face orthogonal to MU (3-d lexicographic, x fastest, >> 1)
template <int MU> __device__ __forceinline__ int face_index(const int *x, const Geom &g) {
  if (MU == 0) return (x[1] + g.X[1] * (x[2] + g.X[2] * x[3])) >> 1;
  if (MU == 1) return (x[0] + g.X[0] * (x[2] + g.X[2] * x[3])) >> 1;
  if (MU == 2) return (x[0] + g.X[0] * (x[1] + g.X[1] * x[3])) >> 1;
  return (x[0] + g.X[0] * (x[1] + g.X[1] * x[2])) >> 1;
}

__device__ __forceinline__ int dim_stride(int mu, const Geom &g) {
  return mu == 0 ? 1 : (mu == 1 ? g.X[0] : (mu == 2 ? g.X[0] * g.X[1] : g.X[0] * g.X[1] * g.X[2]));
}

template <typename Store, int RECON> __device__ __forceinline__ size_t link_block_bytes(long stride) {
  return (size_t)RECON * StoreTraits<Store>::real_bytes * stride;
}

// packed-pair math for this storage type?  (QB_PACKED_S / QB_NO_F32X2: tuning builds)
template <typename Store> __host__ __device__ constexpr bool use_packed() {
#ifdef QB_NO_F32X2
  return false;
#else
  return Store::packed_math;
#endif
}

// one hop: MU direction, BACK = 0 forward (x+mu), 1 backward (x-mu); GHOST = false compiles the ghost-zone branches away
// (unpartitioned lattices and the interior launch of partitioned ones: straight-line code, loads hoisted across hops)
template <typename Store, int RECON, bool TWIST_IN, bool GHOST, bool BATCH, int MU, int BACK>
__device__ __forceinline__ void hop(cplx<typename Store::real> *acc, const DslashParam &p, const void *in, const int *x, int full, int cb) {
  typedef typename Store::real real;
  const Geom &g = p.g;
  const int L = g.X[MU];
  const int step = dim_stride(MU, g);
  const bool edge = BACK ? (x[MU] == 0) : (x[MU] == L - 1);
  constexpr bool PK = use_packed<Store>();
  const bool use_ghost = GHOST && edge && g.part[MU];
  const real sigma = BACK ? (real)(-p.sgn_fwd) : (real)p.sgn_fwd;

  cplx<real> h[6];
  int nbr = 0, fidx = 0;
  real sc;  // storage scale of the loaded (half) spinor, folded into the accumulation below
  if (use_ghost) {
    fidx = face_index<MU>(x, g);
    sc = Store::template load<6>(h, p.ghost[MU][BACK ? 0 : 1], p.ghost_norm[MU][BACK ? 0 : 1], g.faceVh[MU], fidx);
  } else {
    const int nfull = BACK ? (edge ? full + (L - 1) * step : full - step) : (edge ? full - (L - 1) * step : full + step);
    nbr = nfull >> 1;
    cplx<real> psi[12];
    sc = Store::template load<12>(psi, in, p.in_norm, p.stride, nbr);
    if (TWIST_IN) { if constexpr (PK) apply_twist_pk(psi, (real)p.cin[0], (real)p.cin[1]); else apply_twist(psi, (real)p.cin[0], (real)p.cin[1]); }
    if constexpr (PK) project_pk<MU>(h, psi, sigma); else project<MU>(h, psi, sigma);
  }

  // link: forward hop uses U_mu(x) (own parity, own site); backward uses U_mu(x-mu)^dag (other parity)
  real raw[RECON];
  if (!BACK) {
    LinkRaw<Store, RECON>::template load<!BATCH>(raw, (const char *)p.gauge_fwd + MU * link_block_bytes<Store, RECON>(p.stride), p.stride, cb);
  } else if (use_ghost) {
    LinkRaw<Store, RECON>::template load<!BATCH>(raw, p.gauge_ghost[MU], g.faceVh[MU], fidx);
  } else {
    LinkRaw<Store, RECON>::template load<!BATCH>(raw, (const char *)p.gauge_bwd + MU * link_block_bytes<Store, RECON>(p.stride), p.stride, nbr);
  }
  real u0;
  const real an = sizeof(real) == 8 ? (real)g.aniso : (real)g.aniso_f;
  if (MU < 3) u0 = RECON == 8 ? (real)1 / an : an;
  else u0 = BACK ? (x[3] == 0 ? (real)g.tb_bwd : (real)1) : (x[3] == L - 1 ? (real)g.tb_fwd : (real)1);
  cplx<real> U[9];
  reconstruct_link<real, RECON, PK>(U, raw, link_u0<Store, RECON>(u0));

  cplx<real> chi[6];
  if constexpr (PK) {
    su3_mul_pk<BACK != 0>(chi, U, h);
    reconstruct_acc_pk<Store::scaled, MU>(acc, chi, sigma, sc * link_scale<Store, RECON>());
  } else {
    su3_mul<BACK != 0>(chi, U, h);
    reconstruct_acc<Store::scaled, MU>(acc, chi, sigma, sc * link_scale<Store, RECON>());
  }
}

// acc <- S acc with S = (C + i a gamma5)^-1 (ADJ = false) or its conjugate transpose (ADJ = true): the site-local factor of the
// twisted-clover even-odd operator fused into the hop's epilogue (the reference fuses it the same way, lib/tmc_dslash_def.h; as a second
// launch it costs one more write and read of the spinor, 192 of 1344 B per site in fp32).  One 6 x 6 block per chirality, streamed row
// by row (12 reals) so that only one row is live.
template <typename real, bool ADJ, bool PK, bool I16>
__device__ __forceinline__ void clover_inv_mul(cplx<real> *acc, const void *Av, const float *Anorm, int parity, long Vh, long cb) {
  constexpr int RP = sizeof(real) == 8 ? 2 : 4;
  real scale = (real)1;
  if constexpr (I16) scale = (real)(ld_nc(Anorm + (size_t)parity * Vh + cb) * (1.0f / HALF_MAX));
#pragma unroll
  for (int chi = 0; chi < 2; chi++) {
    cplx<real> *v = acc + 6 * chi;
    cplx<real> o[6];
#pragma unroll
    for (int i = 0; i < 6; i++) o[i] = cplx<real>((real)0, (real)0);
#pragma unroll
    for (int i = 0; i < 6; i++) {
      real m[12];
#pragma unroll
      for (int q = 0; q < 12 / RP; q++) {
        const size_t plane = (size_t)parity * (144 / RP) + (chi * 72 + i * 12) / RP + q;
        if constexpr (I16) {
          const int2 t = ld_stream((const int2 *)Av + plane * Vh + cb);
          float a0, a1, a2, a3;
          unpack_s16x2_raw(t.x, a0, a1);
          unpack_s16x2_raw(t.y, a2, a3);
          m[4 * q] = (real)a0; m[4 * q + 1] = (real)a1; m[4 * q + 2] = (real)a2; m[4 * q + 3] = (real)a3;
        } else {
          const real *src = (const real *)Av + (plane * Vh + cb) * RP;
          if constexpr (RP == 2) { const double2 t = ld_stream((const double2 *)src); m[2 * q] = t.x; m[2 * q + 1] = t.y; }
          else { const float4 t = ld_stream((const float4 *)src); m[4 * q] = t.x; m[4 * q + 1] = t.y; m[4 * q + 2] = t.z; m[4 * q + 3] = t.w; }
        }
      }
#pragma unroll
      for (int j = 0; j < 6; j++) {
        const cplx<real> b(m[2 * j], m[2 * j + 1]);
        if (!ADJ) { if constexpr (PK) cmac_pk(o[i], b, v[j]); else cmac(o[i], b, v[j]); }
        else cmac_conj(o[j], b, v[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 6; i++) v[i] = I16 ? cplx<real>(o[i].re * scale, o[i].im * scale) : o[i];
  }
}

// Launch bounds from the B200 sweeps of profiles/README_r01.md: the kernel is HBM-latency bound; with the ghost-zone
// branches compiled away fp32 / int16 run best at 72 registers (7 CTAs of 128 threads per SM: the extra registers let
// the compiler keep the next hop's loads in flight), fp64 needs 128 registers to avoid spills.
// QB_DSLASH_MINB overrides for tuning builds.
template <typename Store> struct DslashBounds { static constexpr int max_threads = 128, min_blocks = 7; static constexpr bool prefetch_links = false; };
template <> struct DslashBounds<StoreS> { static constexpr int max_threads = 128, min_blocks = 7; static constexpr bool prefetch_links = true; };
template <> struct DslashBounds<StoreD> { static constexpr int max_threads = 128, min_blocks = 4; static constexpr bool prefetch_links = false; };
#ifdef QB_DSLASH_MINB
#define QB_DSLASH_BOUNDS __launch_bounds__(QB_DSLASH_MAXT, QB_DSLASH_MINB)
#else
#define QB_DSLASH_BOUNDS __launch_bounds__(DslashBounds<Store>::max_threads, DslashBounds<Store>::min_blocks)
#endif
// BATCH: blockDim = (32 sites, nbatch members <= 12), every warp works on one member of a batched field for the same 32 sites, so the
// links of those sites come from HBM once and from L1 for the other members (they are loaded with L1 allocation here, not streamed)
constexpr int DSLASH_BATCH_MAX = 12;
template <typename Store, int RECON, bool TWIST_IN, bool HAS_X, bool GHOST, bool BATCH> struct DslashLaunchBounds {
  static constexpr int max_threads = BATCH ? 32 * DSLASH_BATCH_MAX : DslashBounds<Store>::max_threads;
  static constexpr int min_blocks = BATCH ? (sizeof(typename Store::real) == 8 ? 1 : 2) : DslashBounds<Store>::min_blocks;
};
// CLOVER: 0 none, 1 inverse twisted-clover block on the output, 2 its conjugate transpose
template <typename Store, int RECON, bool TWIST_IN, bool HAS_X, bool GHOST, bool BATCH = false, int CLOVER = 0>
__global__ void
#ifdef QB_DSLASH_MINB
QB_DSLASH_BOUNDS
#else
__launch_bounds__(DslashLaunchBounds<Store, RECON, TWIST_IN, HAS_X, GHOST, BATCH>::max_threads, DslashLaunchBounds<Store, RECON, TWIST_IN, HAS_X, GHOST, BATCH>::min_blocks)
#endif
dslash_kernel(const DslashParam p) {
  typedef typename Store::real real;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= p.site_count) return;
  const int cb = p.site_list ? p.site_list[tid] : p.site_begin + tid;
  const void *in = p.in, *xin = p.x;
  void *out = p.out;
  if (BATCH) {
    in = (const char *)p.in + (size_t)threadIdx.y * p.batch_in;
    out = (char *)p.out + (size_t)threadIdx.y * p.batch_out;
    xin = (const char *)p.x + (size_t)threadIdx.y * p.batch_x;
  }

  int x[4], full;
  cb_coords(x, full, cb, p.parity, p.g);

  cplx<real> acc[12];
#pragma unroll
  for (int k = 0; k < 12; k++) acc[k] = cplx<real>((real)0, (real)0);

  // forward links of this site are pure DRAM streams: pull the y/z/t ones towards L2 before their hops need them
  // (fp32 storage only: 102.7 -> 100.8 us on B200; int16 unchanged, fp64 slower -- profiles/README_r01.md)
  if (DslashBounds<Store>::prefetch_links && RECON != 18) {
    const size_t eb = 4 * StoreTraits<Store>::real_bytes;
#pragma unroll
    for (int mu = 1; mu < 4; mu++)
#pragma unroll
      for (int k = 0; k < RECON / 4; k++) {
        const char *a = (const char *)p.gauge_fwd + mu * link_block_bytes<Store, RECON>(p.stride) + ((size_t)k * p.stride + cb) * eb;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
      }
  }
  hop<Store, RECON, TWIST_IN, GHOST, BATCH, 0, 0>(acc, p, in, x, full, cb);
  hop<Store, RECON, TWIST_IN, GHOST, BATCH, 0, 1>(acc, p, in, x, full, cb);
  hop<Store, RECON, TWIST_IN, GHOST, BATCH, 1, 0>(acc, p, in, x, full, cb);
  hop<Store, RECON, TWIST_IN, GHOST, BATCH, 1, 1>(acc, p, in, x, full, cb);
  hop<Store, RECON, TWIST_IN, GHOST, BATCH, 2, 0>(acc, p, in, x, full, cb);
  hop<Store, RECON, TWIST_IN, GHOST, BATCH, 2, 1>(acc, p, in, x, full, cb);
  hop<Store, RECON, TWIST_IN, GHOST, BATCH, 3, 0>(acc, p, in, x, full, cb);
  hop<Store, RECON, TWIST_IN, GHOST, BATCH, 3, 1>(acc, p, in, x, full, cb);

  // epilogue: out = Cx x + Co acc
  constexpr bool PK = use_packed<Store>();
  if constexpr (PK) apply_twist_pk(acc, (real)p.co[0], (real)p.co[1]); else apply_twist(acc, (real)p.co[0], (real)p.co[1]);
  if constexpr (CLOVER != 0) clover_inv_mul<real, CLOVER == 2, PK, Store::scaled>(acc, p.clover_inv, p.clover_norm, p.parity, p.stride, cb);
  if (HAS_X) {
    cplx<real> xs[12];
    const real xsc = Store::template load<12, false>(xs, xin, p.x_norm, p.stride, cb);
    const cplx<real> cu((real)p.cx[0] * xsc, (real)p.cx[1] * xsc), cl((real)p.cx[0] * xsc, -(real)p.cx[1] * xsc);
#pragma unroll
    for (int k = 0; k < 6; k++) { if constexpr (PK) cmac_pk(acc[k], cu, xs[k]); else cmac(acc[k], cu, xs[k]); }
#pragma unroll
    for (int k = 6; k < 12; k++) { if constexpr (PK) cmac_pk(acc[k], cl, xs[k]); else cmac(acc[k], cl, xs[k]); }
  }
  Store::template store<12>(out, p.out_norm, p.stride, cb, acc);
}

// ---- face packing (replaces lib/dslash_pack.cu:271-339, :609-674) -------------------------------
// One thread per (partitioned dim, direction, face site): load the full spinor of the input field on
// the boundary slice, apply Cin if the operator twists its input, project with the projector the
// *receiver* will use, store the 12-real half spinor into the send buffer [plane][face site].
struct PackParam {
  Geom g;
  int parity;                 // parity of the field being packed (= input parity of the hop)
  const void *in; const float *in_norm;
  long stride;
  void *send[4][2];           // [d][0]: slice x_d = 0 (sent backward), [d][1]: slice x_d = X_d-1 (sent forward)
  float *send_norm[4][2];
  int thread_off[5];          // prefix sums over partitioned dims of 2*faceVh[d]
  double cin[2];
  double sgn_fwd;
};

template <int MU> __device__ __forceinline__ int face_to_cb(int fidx, int slice, int parity, const Geom &g) {
  // invert face_index<MU> for the site of the given parity on the slice x_MU = slice
  int a, b, c;  // the three remaining coordinates, fastest first
  const int d0 = MU == 0 ? 1 : 0, d1 = MU <= 1 ? 2 : 1, d2 = MU <= 2 ? 3 : 2;
  const int L0 = g.X[d0], L1 = g.X[d1];
  const int f2 = 2 * fidx;
  const int row = f2 / L0;
  c = row / L1;
  b = row - c * L1;
  a = f2 - row * L0;
  a += (slice + b + c + parity + a) & 1;  // a is even here; fix the checkerboard offset
  int x[4];
  x[MU] = slice; x[d0] = a; x[d1] = b; x[d2] = c;
  return (((x[3] * g.X[2] + x[2]) * g.X[1] + x[1]) * g.X[0] + x[0]) >> 1;
}

template <typename Store, bool TWIST_IN, int MU>
__device__ __forceinline__ void pack_site(const PackParam &p, int t) {
  typedef typename Store::real real;
  const Geom &g = p.g;
  const int fv = g.faceVh[MU];
  const int dir = t / fv;  // 0: back face (x=0), 1: forward face (x=L-1)
  const int fidx = t - dir * fv;
  const int cb = face_to_cb<MU>(fidx, dir ? g.X[MU] - 1 : 0, p.parity, g);
  cplx<real> psi[12], h[6];
  load_scaled<Store, 12>(psi, p.in, p.in_norm, p.stride, cb);
  if (TWIST_IN) apply_twist(psi, (real)p.cin[0], (real)p.cin[1]);
  // back face feeds the neighbour's forward hop (sigma = sgn_fwd); forward face its backward hop
  const real sigma = dir ? (real)(-p.sgn_fwd) : (real)p.sgn_fwd;
  project<MU>(h, psi, sigma);
  Store::template store<6>(p.send[MU][dir], p.send_norm[MU][dir], fv, fidx, h);
}

template <typename Store, bool TWIST_IN>
__global__ void __launch_bounds__(128) pack_kernel(const PackParam p) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= p.thread_off[4]) return;
  if (tid < p.thread_off[1]) pack_site<Store, TWIST_IN, 0>(p, tid - p.thread_off[0]);
  else if (tid < p.thread_off[2]) pack_site<Store, TWIST_IN, 1>(p, tid - p.thread_off[1]);
  else if (tid < p.thread_off[3]) pack_site<Store, TWIST_IN, 2>(p, tid - p.thread_off[2]);
  else pack_site<Store, TWIST_IN, 3>(p, tid - p.thread_off[3]);
}

// ---- site-local twist (replaces twistGamma5Cuda, lib/dslash_quda.cu:430-463) --------------------
template <typename Store>
__global__ void __launch_bounds__(256) twist_kernel(void *out, float *out_norm, const void *in, const float *in_norm, long stride,
                                                    int n, double pr, double qr) {
  typedef typename Store::real real;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  cplx<real> psi[12];
  load_scaled<Store, 12, false>(psi, in, in_norm, stride, i);
  apply_twist(psi, (real)pr, (real)qr);
  Store::template store<12>(out, out_norm, stride, i, psi);
}

}  // namespace qb
