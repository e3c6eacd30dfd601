// Device-side accessors for the plane layouts described in field.h.
#pragma once
#include <cuda_fp16.h>
#include "field.h"
#include "f32x2.cuh"

namespace qb {

template <typename T> struct cplx {
  T re, im;
  __host__ __device__ cplx() {}
  __host__ __device__ cplx(T r, T i) : re(r), im(i) {}
};
template <typename T> __host__ __device__ inline cplx<T> operator+(cplx<T> a, cplx<T> b) { return cplx<T>(a.re + b.re, a.im + b.im); }
template <typename T> __host__ __device__ inline cplx<T> operator-(cplx<T> a, cplx<T> b) { return cplx<T>(a.re - b.re, a.im - b.im); }
template <typename T> __host__ __device__ inline cplx<T> operator*(cplx<T> a, cplx<T> b) { return cplx<T>(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
template <typename T> __host__ __device__ inline cplx<T> conj(cplx<T> a) { return cplx<T>(a.re, -a.im); }
// a += b*c
template <typename T> __host__ __device__ inline void cmac(cplx<T> &a, cplx<T> b, cplx<T> c) {
  a.re += b.re * c.re; a.re -= b.im * c.im; a.im += b.re * c.im; a.im += b.im * c.re;
}
// a += conj(b)*c
template <typename T> __host__ __device__ inline void cmac_conj(cplx<T> &a, cplx<T> b, cplx<T> c) {
  a.re += b.re * c.re; a.re += b.im * c.im; a.im += b.re * c.im; a.im -= b.im * c.re;
}

// ------------------------------------------------------------------------------------------
// 128-bit global loads.  ld_nc: read-only data path (spinor neighbours, links);
// ld_stream: links are touched exactly once per hop -> do not pollute L1 (ld.global.nc.L1::no_allocate)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld_nc(const float4 *p) { return __ldg(p); }
__device__ __forceinline__ double2 ld_nc(const double2 *p) { return __ldg(p); }
__device__ __forceinline__ float2 ld_nc(const float2 *p) { return __ldg(p); }
__device__ __forceinline__ int4 ld_nc(const int4 *p) { return __ldg(p); }
__device__ __forceinline__ int2 ld_nc(const int2 *p) { return __ldg(p); }
__device__ __forceinline__ int ld_nc(const int *p) { return __ldg(p); }
__device__ __forceinline__ float ld_nc(const float *p) { return __ldg(p); }

__device__ __forceinline__ float4 ld_stream(const float4 *p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ double2 ld_stream(const double2 *p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ float2 ld_stream(const float2 *p) {
  float2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ int2 ld_stream(const int2 *p) {
  int2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ int ld_stream(const int *p) {
  int r;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}

// ------------------------------------------------------------------------------------------
// Spinor storage policies.  NC = complex components handled per call (12 for a fine site,
// 6 for a half spinor in a ghost / face buffer).
// ------------------------------------------------------------------------------------------
constexpr float HALF_MAX = 32767.0f;  // int16 fixed point of the reference (quda_internal.h:30, io_spinor.h:49-62)

// Every policy's load() returns the factor the loaded values still have to be multiplied by
// (1 for fp64/fp32).  The hop kernel folds it into the accumulation FMA, so that int16 storage
// costs no extra multiplies; everything else uses load_scaled().
struct StoreD {
  typedef double real;
  static constexpr Prec prec = PREC_DOUBLE;
  static constexpr bool scaled = false;
  static constexpr bool packed_math = false;
  template <int NC, bool NC_PATH = true>
  __device__ __forceinline__ static double load(cplx<double> *o, const void *base, const float *, long stride, long i) {
    const double2 *p = (const double2 *)base + i;
#pragma unroll
    for (int k = 0; k < NC; k++) {
      double2 t = NC_PATH ? ld_nc(p + (long)k * stride) : p[(long)k * stride];
      o[k] = cplx<double>(t.x, t.y);
    }
    return 1.0;
  }
  template <int NC>
  __device__ __forceinline__ static void store(void *base, float *, long stride, long i, const cplx<double> *v) {
    double2 *p = (double2 *)base + i;
#pragma unroll
    for (int k = 0; k < NC; k++) p[(long)k * stride] = make_double2(v[k].re, v[k].im);
  }
};

struct StoreS {
  typedef float real;
  static constexpr Prec prec = PREC_SINGLE;
  static constexpr bool scaled = false;
#ifdef QB_SCALAR_S
  static constexpr bool packed_math = false;
#else
  static constexpr bool packed_math = true;  // f32x2.cuh
#endif
  template <int NC, bool NC_PATH = true>
  __device__ __forceinline__ static float load(cplx<float> *o, const void *base, const float *, long stride, long i) {
    const float4 *p = (const float4 *)base + i;
#pragma unroll
    for (int k = 0; k < NC / 2; k++) {
      float4 t = NC_PATH ? ld_nc(p + (long)k * stride) : p[(long)k * stride];
      o[2 * k] = cplx<float>(t.x, t.y);
      o[2 * k + 1] = cplx<float>(t.z, t.w);
    }
    return 1.0f;
  }
  template <int NC>
  __device__ __forceinline__ static void store(void *base, float *, long stride, long i, const cplx<float> *v) {
    float4 *p = (float4 *)base + i;
#pragma unroll
    for (int k = 0; k < NC / 2; k++) p[(long)k * stride] = make_float4(v[2 * k].re, v[2 * k].im, v[2 * k + 1].re, v[2 * k + 1].im);
  }
};

// int16 <-> fp32 without the quarter-rate I2F/F2I conversion pipe (the first profile of the half
// kernel showed 387 I2F + 592 FMUL per thread).  The 16-bit integers are kept in HBM in offset-binary
// form u = s + 32768 (same values as the reference's two's-complement shorts, different encoding), so
//   as_float(0x4B000000 | u) = 2^23 + u   exactly   =>   float(s) = that - (2^23 + 32768)
// costs one PRMT and one exact FADD per value on the full-rate pipes.
__device__ __forceinline__ void unpack_s16x2_raw(int w, float &a, float &b) {
  const unsigned lo = __byte_perm((unsigned)w, 0x4B000000u, 0x7610);
  const unsigned hi = __byte_perm((unsigned)w, 0x4B000000u, 0x7632);
#ifndef QB_NO_F32X2
  unpk2(add2(pk2(__uint_as_float(lo), __uint_as_float(hi)), bc2(-8421376.0f)), a, b);
#else
  a = __uint_as_float(lo) - 8421376.0f;
  b = __uint_as_float(hi) - 8421376.0f;
#endif
}
// fp32 -> offset-binary int16 pair, round to nearest even: x*scale + (1.5*2^23 + 32768) leaves u in the low mantissa bits
__device__ __forceinline__ int pack_s16x2(float a, float b, float scale) {
#ifndef QB_NO_F32X2
  float fa, fb;
  unpk2(fma2(pk2(a, b), bc2(scale), bc2(12615680.0f)), fa, fb);
  const unsigned ia = __float_as_uint(fa), ib = __float_as_uint(fb);
#else
  const unsigned ia = __float_as_uint(fmaf(a, scale, 12615680.0f));
  const unsigned ib = __float_as_uint(fmaf(b, scale, 12615680.0f));
#endif
  return (int)__byte_perm(ia, ib, 0x5410);
}

// int16 storage + one float norm per site; arithmetic in fp32
struct StoreH {
  typedef float real;
  static constexpr Prec prec = PREC_HALF;
  static constexpr bool scaled = true;
  static constexpr bool packed_math = true;  // f32x2.cuh
  template <int NC, bool NC_PATH = true>
  __device__ __forceinline__ static float load(cplx<float> *o, const void *base, const float *norm, long stride, long i) {
    const float c = (NC_PATH ? ld_nc(norm + i) : norm[i]) * (1.0f / HALF_MAX);
    if (NC % 4 == 0) {
      const int4 *p = (const int4 *)base + i;
#pragma unroll
      for (int k = 0; k < NC / 4; k++) {
        int4 t = NC_PATH ? ld_nc(p + (long)k * stride) : p[(long)k * stride];
        unpack_s16x2_raw(t.x, o[4 * k].re, o[4 * k].im);
        unpack_s16x2_raw(t.y, o[4 * k + 1].re, o[4 * k + 1].im);
        unpack_s16x2_raw(t.z, o[4 * k + 2].re, o[4 * k + 2].im);
        unpack_s16x2_raw(t.w, o[4 * k + 3].re, o[4 * k + 3].im);
      }
    } else {  // half spinors (6 complex): 8-byte planes of 2 complex
      const int2 *p = (const int2 *)base + i;
#pragma unroll
      for (int k = 0; k < NC / 2; k++) {
        int2 t = NC_PATH ? ld_nc(p + (long)k * stride) : p[(long)k * stride];
        unpack_s16x2_raw(t.x, o[2 * k].re, o[2 * k].im);
        unpack_s16x2_raw(t.y, o[2 * k + 1].re, o[2 * k + 1].im);
      }
    }
    return c;
  }
  template <int NC>
  __device__ __forceinline__ static void store(void *base, float *norm, long stride, long i, const cplx<float> *v) {
    float m = 0.0f;
#pragma unroll
    for (int k = 0; k < NC; k++) m = fmaxf(m, fmaxf(fabsf(v[k].re), fabsf(v[k].im)));
    norm[i] = m;
    const float s = m > 0.0f ? HALF_MAX / m : 0.0f;
    if (NC % 4 == 0) {
      int4 *p = (int4 *)base + i;
#pragma unroll
      for (int k = 0; k < NC / 4; k++)
        p[(long)k * stride] = make_int4(pack_s16x2(v[4 * k].re, v[4 * k].im, s), pack_s16x2(v[4 * k + 1].re, v[4 * k + 1].im, s),
                                         pack_s16x2(v[4 * k + 2].re, v[4 * k + 2].im, s), pack_s16x2(v[4 * k + 3].re, v[4 * k + 3].im, s));
    } else {
      int2 *p = (int2 *)base + i;
#pragma unroll
      for (int k = 0; k < NC / 2; k++)
        p[(long)k * stride] = make_int2(pack_s16x2(v[2 * k].re, v[2 * k].im, s), pack_s16x2(v[2 * k + 1].re, v[2 * k + 1].im, s));
    }
  }
};

// load and apply the storage scale (everything except the hop kernel)
template <typename Store, int NC, bool NC_PATH = true>
__device__ __forceinline__ void load_scaled(cplx<typename Store::real> *o, const void *base, const float *norm, long stride, long i) {
  const typename Store::real sc = Store::template load<NC, NC_PATH>(o, base, norm, stride, i);
  if (Store::scaled) {
#pragma unroll
    for (int k = 0; k < NC; k++) { o[k].re *= sc; o[k].im *= sc; }
  }
}

// bytes of one site-plane element for NC complex per site
template <typename Store> struct StoreTraits;
template <> struct StoreTraits<StoreD> { static constexpr int real_bytes = 8; };
template <> struct StoreTraits<StoreS> { static constexpr int real_bytes = 4; };
template <> struct StoreTraits<StoreH> { static constexpr int real_bytes = 2; };

// ------------------------------------------------------------------------------------------
// Gauge links.  reals per plane: recon 18 -> 2 (one complex); recon 12/8 -> 2 (fp64) or 4 (fp32, half)
// ------------------------------------------------------------------------------------------
__host__ __device__ inline int gauge_reals_per_plane(int prec, int recon) { return recon == 18 ? 2 : (prec == 8 ? 2 : 4); }

// raw load of RECON reals (converted to the compute type) from planes [plane][site]
template <typename Store, int RECON> struct LinkRaw;

template <int RECON> struct LinkRaw<StoreD, RECON> {
  template <bool STREAM = true>
  __device__ __forceinline__ static void load(double *r, const void *base, long stride, long i) {
    const double2 *p = (const double2 *)base + i;
#pragma unroll
    for (int k = 0; k < RECON / 2; k++) {
      double2 t = STREAM ? ld_stream(p + (long)k * stride) : ld_nc(p + (long)k * stride);
      r[2 * k] = t.x; r[2 * k + 1] = t.y;
    }
  }
};
template <int RECON> struct LinkRaw<StoreS, RECON> {
  template <bool STREAM = true>
  __device__ __forceinline__ static void load(float *r, const void *base, long stride, long i) {
    if (RECON == 18) {
      const float2 *p = (const float2 *)base + i;
#pragma unroll
      for (int k = 0; k < 9; k++) {
        float2 t = STREAM ? ld_stream(p + (long)k * stride) : ld_nc(p + (long)k * stride);
        r[2 * k] = t.x; r[2 * k + 1] = t.y;
      }
    } else {
      const float4 *p = (const float4 *)base + i;
#pragma unroll
      for (int k = 0; k < RECON / 4; k++) {
        float4 t = STREAM ? ld_stream(p + (long)k * stride) : ld_nc(p + (long)k * stride);
        r[4 * k] = t.x; r[4 * k + 1] = t.y; r[4 * k + 2] = t.z; r[4 * k + 3] = t.w;
      }
    }
  }
};
// half precision: returns integer-valued floats (the 1/32767 is folded into the hop's accumulation scale,
// see link_scale()); recon 8 is non-linear in the stored numbers and is converted to real units here
template <int RECON> struct LinkRaw<StoreH, RECON> {
  template <bool STREAM = true>
  __device__ __forceinline__ static void load(float *r, const void *base, long stride, long i) {
    if (RECON == 18) {
      const int *p = (const int *)base + i;
#pragma unroll
      for (int k = 0; k < 9; k++) {
        int t = ld_stream(p + (long)k * stride);
        unpack_s16x2_raw(t, r[2 * k], r[2 * k + 1]);
      }
    } else {
      const int2 *p = (const int2 *)base + i;
#pragma unroll
      for (int k = 0; k < RECON / 4; k++) {
        int2 t = ld_stream(p + (long)k * stride);
        unpack_s16x2_raw(t.x, r[4 * k], r[4 * k + 1]);
        unpack_s16x2_raw(t.y, r[4 * k + 2], r[4 * k + 3]);
      }
      if (RECON == 8) {
        const float c = 1.0f / HALF_MAX;
#pragma unroll
        for (int k = 0; k < 6; k++) r[k] *= c;
        r[6] *= c * 3.14159265358979323846f;  // phases are stored / pi
        r[7] *= c * 3.14159265358979323846f;
      }
    }
  }
};

// factor by which a reconstructed link of this storage type still has to be multiplied, and the
// correction of the recon-12 third-row factor u0 (that row is quadratic in the stored numbers)
template <typename Store, int RECON> __device__ __forceinline__ typename Store::real link_scale() {
  return (Store::scaled && RECON != 8) ? (typename Store::real)(1.0f / HALF_MAX) : (typename Store::real)1;
}
template <typename Store, int RECON> __device__ __forceinline__ typename Store::real link_u0(typename Store::real u0) {
  return (Store::scaled && RECON == 12) ? u0 * (typename Store::real)(1.0f / HALF_MAX) : u0;
}

// recon-8 helpers: the reference uses the fast hardware intrinsics for its single-precision reconstruction (lib/read_gauge.h:403-483,
// __sinf / __cosf); they are accurate to ~4e-7 absolute on [-pi, pi], inside the fp32 parity budget (1e-6), and an order of magnitude
// cheaper than the range-reducing sincosf (16 calls per site made recon 8 slower than recon 12)
__device__ __forceinline__ void sincos_r8(double x, double *s, double *c) { sincos(x, s, c); }
__device__ __forceinline__ void sincos_r8(float x, float *s, float *c) { __sincosf(x, s, c); }
__device__ __forceinline__ double neg_rcp_r8(double x) { return -1.0 / x; }
__device__ __forceinline__ float neg_rcp_r8(float x) { return -__frcp_rn(x); }

// Reconstruct the full 3x3 link U[row*3+col] from RECON stored reals.
//  recon 12: rows 0,1 stored; row 2 = conj(row0 x row1) * u0   (cf. lib/read_gauge.h:393-401)
//  recon  8: a2,a3,b1 + phases of a1 and c1 (Bunk/Sommer, cf. lib/read_gauge.h:403-483)
// u0: recon 12 -> factor of the reconstructed row (anisotropy for spatial, boundary sign for temporal links);
//     recon  8 -> factor of the whole link (1/anisotropy resp. boundary sign).
template <typename real, int RECON, bool PK = false>
__device__ __forceinline__ void reconstruct_link(cplx<real> *U, const real *r, real u0) {
  if (RECON == 18) {
#pragma unroll
    for (int k = 0; k < 9; k++) U[k] = cplx<real>(r[2 * k], r[2 * k + 1]);
  } else if (RECON == 12) {
#pragma unroll
    for (int k = 0; k < 6; k++) U[k] = cplx<real>(r[2 * k], r[2 * k + 1]);
#ifndef QB_NO_F32X2
    if constexpr (PK && sizeof(real) == 4) {
      // conj(a b) = a.re (b.re, -b.im) + (-a.im) (b.im, b.re): a from row 0 as broadcast scalars, b from row 1 as pairs
      f2 cb[3], sb[3];
#pragma unroll
      for (int k = 0; k < 3; k++) { cb[k] = mul2(pk2(U[3 + k].re, U[3 + k].im), pk2(1.0f, -1.0f)); sb[k] = pk2(U[3 + k].im, U[3 + k].re); }
      const f2 u02 = bc2(u0);
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const int i = (k + 1) % 3, j = (k + 2) % 3;  // row2[k] = conj(U[i] U[3+j] - U[j] U[3+i]) u0
        f2 t = mul2(bc2(U[i].re), cb[j]);
        t = fma2(bc2(-U[i].im), sb[j], t);
        t = fma2(bc2(-U[j].re), cb[i], t);
        t = fma2(bc2(U[j].im), sb[i], t);
        t = mul2(t, u02);
        unpk2(t, U[6 + k].re, U[6 + k].im);
      }
      return;
    }
#endif
    U[6] = conj(U[1] * U[5] - U[2] * U[4]);
    U[7] = conj(U[2] * U[3] - U[0] * U[5]);
    U[8] = conj(U[0] * U[4] - U[1] * U[3]);
    U[6].re *= u0; U[6].im *= u0; U[7].re *= u0; U[7].im *= u0; U[8].re *= u0; U[8].im *= u0;
  } else {
    // stored (unit-determinant matrix W, scale and boundary sign divided out at import):
    //   r[0..1]=a2, r[2..3]=a3, r[4..5]=b1, r[6]=arg(a1), r[7]=arg(c1)   (rows a,b,c of W);  U = u0 * W
    const cplx<real> a2(r[0], r[1]), a3(r[2], r[3]), b1(r[4], r[5]);
    const real N2 = a2.re * a2.re + a2.im * a2.im + a3.re * a3.re + a3.im * a3.im;
    real a1m2 = (real)1 - N2;
    a1m2 = a1m2 > 0 ? a1m2 : (real)0;
    real sn, cs;
    sincos_r8(r[6], &sn, &cs);
    const real a1m = sqrt(a1m2);
    const cplx<real> a1(a1m * cs, a1m * sn);
    real c1m2 = (real)1 - a1m2 - (b1.re * b1.re + b1.im * b1.im);
    c1m2 = c1m2 > 0 ? c1m2 : (real)0;
    sincos_r8(r[7], &sn, &cs);
    const real c1m = sqrt(c1m2);
    const cplx<real> c1(c1m * cs, c1m * sn);
    const real rN = neg_rcp_r8(N2);
    const cplx<real> A = conj(a1) * b1, B = conj(a1) * c1;
    const cplx<real> cc1 = conj(c1), cb1 = conj(b1), ca2 = conj(a2), ca3 = conj(a3);
    cplx<real> b2 = A * a2 + cc1 * ca3, b3 = A * a3 - cc1 * ca2;
    cplx<real> c2 = B * a2 - cb1 * ca3, c3 = B * a3 + cb1 * ca2;
    U[0] = cplx<real>(u0 * a1.re, u0 * a1.im); U[1] = cplx<real>(u0 * a2.re, u0 * a2.im); U[2] = cplx<real>(u0 * a3.re, u0 * a3.im);
    U[3] = cplx<real>(u0 * b1.re, u0 * b1.im);
    const real f = u0 * rN;
    U[4] = cplx<real>(f * b2.re, f * b2.im); U[5] = cplx<real>(f * b3.re, f * b3.im);
    U[6] = cplx<real>(u0 * c1.re, u0 * c1.im);
    U[7] = cplx<real>(f * c2.re, f * c2.im); U[8] = cplx<real>(f * c3.re, f * c3.im);
  }
}

}  // namespace qb
