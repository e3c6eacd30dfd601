// Packed fp32 pairs for sm_100a: fma/add/mul.f32x2 (SASS FFMA2 / FADD2 / FMUL2) do two fp32 operations per issue
// slot, and the SASS forms take a scalar-broadcast operand (R.F32) and a swapped pair (R.F32x2.LO_HI) for free, so a
// complex multiply-accumulate a += b*c is TWO instructions:
//     a = fma2( (b.re, b.re), (c.re, c.im), a )
//     a = fma2( (b.im, b.im), (-c.im, c.re), a )
// The issue-bound kernels of this library (fine Dslash with int16 storage: ~2600 instructions per site before) are
// written on top of these.  Each lane is an IEEE fp32 fma / add / mul: same arithmetic as the scalar code.
// QB_NO_F32X2 builds the scalar code instead (A/B timing, tools/tune_dslash.py).
#pragma once

namespace qb {

typedef unsigned long long f2;  // (lo, hi) = two floats in an aligned register pair

__device__ __forceinline__ f2 pk2(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f2 bc2(float a) { return pk2(a, a); }
__device__ __forceinline__ void unpk2(f2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

}  // namespace qb
