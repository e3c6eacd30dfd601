// Clover term: import, on-device inversion of the twisted blocks, and the site-local kernel  out = [x +] k S in.
// Reference: lib/clover_field.cpp (storage), lib/clover_invert.cu:32-160 (Cholesky of C^2 + mu^2 per chiral block),
// lib/dslash_core/tmc_core.h / clover_def.h (application), tests/clover_reference.cpp (host order and operator semantics).
#include <cstring>
#include <memory>
#include <vector>
#include "clover.h"
#include "layout.cuh"

namespace qb {

CloverField::CloverField(long Vh_, Prec prec_) : prec(prec_), Vh(Vh_) {
  const size_t rb = prec == PREC_DOUBLE ? 8 : 4;
  QB_CUDA(cudaMalloc(&C, (size_t)2 * Vh * 72 * rb));
  QB_CUDA(cudaMalloc(&Ainv, (size_t)2 * Vh * 144 * rb));
}
CloverField::~CloverField() {
  if (C) cudaFree(C);
  if (Ainv) cudaFree(Ainv);
  if (Ainv16) cudaFree(Ainv16);
  if (Ainv16_norm) cudaFree(Ainv16_norm);
}

// fp32 inverse blocks [parity][36 float4 planes][cb] -> int16 [parity][36 int2 planes][cb] + norm (largest |element| of the site)
__global__ void clover_inv_to_int16_kernel(int2 *out, float *norm, const float4 *in, long Vh) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * Vh) return;
  const long parity = t / Vh, cb = t - parity * Vh;
  float m = 0.0f;
  for (int pl = 0; pl < 36; pl++) {
    const float4 v = in[(parity * 36 + pl) * Vh + cb];
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  norm[t] = m;
  const float s = m > 0.0f ? HALF_MAX / m : 0.0f;
  for (int pl = 0; pl < 36; pl++) {
    const float4 v = in[(parity * 36 + pl) * Vh + cb];
    out[(parity * 36 + pl) * Vh + cb] = make_int2(pack_s16x2(v.x, v.y, s), pack_s16x2(v.z, v.w, s));
  }
}
void CloverField::ainv16() {
  if (prec != PREC_SINGLE) QB_ERROR("int16 inverse clover blocks are derived from the fp32 copy");
  if (have16 && a16 == a) return;
  if (!Ainv16) {
    QB_CUDA(cudaMalloc(&Ainv16, (size_t)2 * Vh * 144 * sizeof(short)));
    QB_CUDA(cudaMalloc((void **)&Ainv16_norm, (size_t)2 * Vh * sizeof(float)));
  }
  clover_inv_to_int16_kernel<<<div_up(2 * Vh, 128), 128, 0, rt().compute>>>((int2 *)Ainv16, Ainv16_norm, (const float4 *)Ainv, Vh);
  QB_CHECK_LAUNCH();
  have16 = true; a16 = a;
}

// index of L(row, col), row > col, in the column-by-column packed lower triangle (clover_reference.cpp:45-52)
__host__ __device__ inline int tri_index(int row, int col) { return 15 - (6 - col) * (5 - col) / 2 + row - col - 1; }

// in-place inverse of a complex 6 x 6 matrix, Gauss-Jordan with partial pivoting, fp64
__device__ void invert6(cplx<double> (*A)[6]) {
  cplx<double> B[6][6];
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) B[i][j] = cplx<double>(i == j ? 1.0 : 0.0, 0.0);
  for (int k = 0; k < 6; k++) {
    int piv = k;
    double best = A[k][k].re * A[k][k].re + A[k][k].im * A[k][k].im;
    for (int r = k + 1; r < 6; r++) {
      const double m = A[r][k].re * A[r][k].re + A[r][k].im * A[r][k].im;
      if (m > best) { best = m; piv = r; }
    }
    if (piv != k)
      for (int c = 0; c < 6; c++) {
        cplx<double> t = A[k][c]; A[k][c] = A[piv][c]; A[piv][c] = t;
        t = B[k][c]; B[k][c] = B[piv][c]; B[piv][c] = t;
      }
    const cplx<double> ip(A[k][k].re / best, -A[k][k].im / best);
    for (int c = 0; c < 6; c++) { A[k][c] = A[k][c] * ip; B[k][c] = B[k][c] * ip; }
    for (int r = 0; r < 6; r++) {
      if (r == k) continue;
      const cplx<double> f = A[r][k];
      for (int c = 0; c < 6; c++) { A[r][c] = A[r][c] - f * A[k][c]; B[r][c] = B[r][c] - f * B[k][c]; }
    }
  }
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) A[i][j] = B[i][j];
}

__device__ inline void unpack_block(cplx<double> (*H)[6], const double *p) {
  for (int i = 0; i < 6; i++) H[i][i] = cplx<double>(p[i], 0.0);
  for (int col = 0; col < 6; col++)
    for (int row = col + 1; row < 6; row++) {
      const int k = tri_index(row, col);
      H[row][col] = cplx<double>(p[6 + 2 * k], p[6 + 2 * k + 1]);
      H[col][row] = cplx<double>(p[6 + 2 * k], -p[6 + 2 * k + 1]);
    }
}

// element e of the per-site record (nreals reals) of site (parity, cb) in the plane layout
template <typename real> __device__ inline real *plane_elem(real *base, int nreals, long Vh, int parity, long cb, int e) {
  constexpr int RP = sizeof(real) == 8 ? 2 : 4;
  const int npl = nreals / RP;
  return base + (((size_t)parity * npl + e / RP) * Vh + cb) * RP + e % RP;
}

// thread = (site, chirality): copy the packed block, build (C + i s a)^-1
template <typename real>
__global__ void clover_prepare_kernel(real *C, real *Ainv, const double *master, long Vh, double a, int write_c) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 4 * Vh) return;
  const int chi = (int)(t & 1);
  const long fs = t >> 1;
  const int parity = fs >= Vh ? 1 : 0;
  const long cb = fs - (long)parity * Vh;
  const double *p = master + (fs * 2 + chi) * 36;
  if (write_c)
    for (int e = 0; e < 36; e++) *plane_elem(C, 72, Vh, parity, cb, chi * 36 + e) = (real)p[e];
  cplx<double> H[6][6];
  unpack_block(H, p);
  const double sa = chi == 0 ? a : -a;
  for (int i = 0; i < 6; i++) H[i][i].im += sa;
  invert6(H);
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) {
      *plane_elem(Ainv, 144, Vh, parity, cb, chi * 72 + (i * 6 + j) * 2) = (real)H[i][j].re;
      *plane_elem(Ainv, 144, Vh, parity, cb, chi * 72 + (i * 6 + j) * 2 + 1) = (real)H[i][j].im;
    }
}

// (C^2 + a2)^-1 per block, packed like the input, fp64 (what the reference's loadCloverQuda returns as the "inverse" of a
// twisted clover term: lib/clover_invert.cu:56-90)
__global__ void clover_sq_inverse_kernel(double *out, const double *master, long nblk, double a2) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nblk) return;
  cplx<double> H[6][6], S[6][6];
  unpack_block(H, master + t * 36);
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) {
      cplx<double> s(i == j ? a2 : 0.0, 0.0);
      for (int k = 0; k < 6; k++) cmac(s, H[i][k], H[k][j]);
      S[i][j] = s;
    }
  invert6(S);
  double *o = out + t * 36;
  for (int i = 0; i < 6; i++) o[i] = S[i][i].re;
  for (int col = 0; col < 6; col++)
    for (int row = col + 1; row < 6; row++) {
      const int k = tri_index(row, col);
      o[6 + 2 * k] = S[row][col].re; o[6 + 2 * k + 1] = S[row][col].im;
    }
}

void CloverSet::load(const void *h_clover, Prec host_prec, long Vh_) {
  release();
  Vh = Vh_;
  const size_t n = (size_t)2 * Vh * 72;
  QB_CUDA(cudaMalloc((void **)&master, n * sizeof(double)));
  if (host_prec == PREC_DOUBLE) {
    QB_CUDA(cudaMemcpy(master, h_clover, n * sizeof(double), cudaMemcpyHostToDevice));
  } else if (host_prec == PREC_SINGLE) {
    std::vector<double> tmp(n);
    const float *f = (const float *)h_clover;
    for (size_t i = 0; i < n; i++) tmp[i] = f[i];
    QB_CUDA(cudaMemcpy(master, tmp.data(), n * sizeof(double), cudaMemcpyHostToDevice));
  } else QB_ERROR("clover_cpu_prec must be double or single");
  loaded = true;
}

__global__ void clover_to_float_kernel(float *out, const double *in, long n) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) out[t] = (float)in[t];
}
float *CloverSet::site_major_f32() const {
  if (!loaded) QB_ERROR("no clover field resident");
  const long n = 2 * Vh * 72;
  float *out = (float *)pool_malloc(sizeof(float) * n);
  clover_to_float_kernel<<<div_up(n, 256), 256, 0, rt().compute>>>(out, master, n);
  QB_CHECK_LAUNCH();
  return out;
}

void CloverSet::release() {
  if (master) cudaFree(master);
  master = nullptr;
  for (int i = 0; i < CLOVER_CACHE; i++) { d64[i].f.reset(); f32[i].f.reset(); }
  loaded = false;
}

const CloverField &CloverSet::get(Prec prec, double a) {
  if (!loaded) QB_ERROR("no clover field resident: call loadCloverQuda first");
  Slot *slots = prec == PREC_DOUBLE ? d64 : f32;
  // the copy built for this twist, else an empty slot, else the least recently used one
  int pick = -1;
  for (int i = 0; i < CLOVER_CACHE; i++)
    if (slots[i].f && slots[i].f->a == a) pick = i;
  bool rebuild = false;
  if (pick < 0) {
    rebuild = true;
    for (int i = 0; i < CLOVER_CACHE && pick < 0; i++)
      if (!slots[i].f) pick = i;
    if (pick < 0) {
      pick = 0;
      for (int i = 1; i < CLOVER_CACHE; i++)
        if (slots[i].last_use < slots[pick].last_use) pick = i;
    }
  }
  std::unique_ptr<CloverField> &f = slots[pick].f;
  slots[pick].last_use = ++use_clock;
  const bool fresh = !f;
  if (fresh) f.reset(new CloverField(Vh, prec == PREC_DOUBLE ? PREC_DOUBLE : PREC_SINGLE));
  if (fresh || rebuild) {
    const long n = 4 * Vh;
    if (f->prec == PREC_DOUBLE) clover_prepare_kernel<double><<<div_up(n, 128), 128, 0, rt().compute>>>((double *)f->C, (double *)f->Ainv, master, Vh, a, fresh ? 1 : 0);
    else clover_prepare_kernel<float><<<div_up(n, 128), 128, 0, rt().compute>>>((float *)f->C, (float *)f->Ainv, master, Vh, a, fresh ? 1 : 0);
    QB_CHECK_LAUNCH();
    f->a = a;
  }
  return *f;
}

void CloverSet::inverse_to_host(void *h_clovinv, Prec host_prec, double a2) {
  if (!loaded) QB_ERROR("no clover field resident");
  const long nblk = 4 * Vh;
  double *d;
  QB_CUDA(cudaMalloc((void **)&d, (size_t)nblk * 36 * sizeof(double)));
  clover_sq_inverse_kernel<<<div_up(nblk, 128), 128, 0, rt().compute>>>(d, master, nblk, a2);
  QB_CHECK_LAUNCH();
  std::vector<double> h((size_t)nblk * 36);
  QB_CUDA(cudaMemcpyAsync(h.data(), d, h.size() * sizeof(double), cudaMemcpyDeviceToHost, rt().compute));
  QB_CUDA(cudaStreamSynchronize(rt().compute));
  QB_CUDA(cudaFree(d));
  if (host_prec == PREC_DOUBLE) memcpy(h_clovinv, h.data(), h.size() * sizeof(double));
  else {
    float *f = (float *)h_clovinv;
    for (size_t i = 0; i < h.size(); i++) f[i] = (float)h[i];
  }
}

// ---- application ------------------------------------------------------------------------------------------------------------
template <typename creal, int N> __device__ __forceinline__ void load_reals(creal *dst, const creal *base, int first, int nreals, long Vh, int parity, long cb) {
  constexpr int RP = sizeof(creal) == 8 ? 2 : 4;
  const int npl = nreals / RP;
#pragma unroll
  for (int p = 0; p < N / RP; p++) {
    const creal *src = base + (((size_t)parity * npl + first / RP + p) * Vh + cb) * RP;
    if (RP == 2) { const double2 v = __ldg((const double2 *)src); dst[2 * p] = (creal)v.x; dst[2 * p + 1] = (creal)v.y; }
    else { const float4 v = __ldg((const float4 *)src); dst[4 * p] = (creal)v.x; dst[4 * p + 1] = (creal)v.y; dst[4 * p + 2] = (creal)v.z; dst[4 * p + 3] = (creal)v.w; }
  }
}

template <typename Store, int MODE, bool HAS_X>
__global__ void __launch_bounds__(128) clover_apply_kernel(void *out, float *out_norm, const void *in, const float *in_norm, const void *x, const float *x_norm,
                                                           const void *Cv, const void *Av, long Vh, int parity, double a_, double k_) {
  typedef typename Store::real real;
  typedef real creal;  // the clover copy is kept in the arithmetic type of the operator
  const long cb = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (cb >= Vh) return;
  cplx<real> psi[12], res[12];
  load_scaled<Store, 12>(psi, in, in_norm, Vh, cb);
#pragma unroll
  for (int chi = 0; chi < 2; chi++) {
    const cplx<real> *v = psi + 6 * chi;
    cplx<real> *o = res + 6 * chi;
    if (MODE == CLOVER_DIRECT) {
      creal m[36];
      load_reals<creal, 36>(m, (const creal *)Cv, chi * 36, 72, Vh, parity, cb);
      const real sa = (real)(chi == 0 ? a_ : -a_);
#pragma unroll
      for (int i = 0; i < 6; i++) o[i] = cplx<real>((real)m[i] * v[i].re - sa * v[i].im, (real)m[i] * v[i].im + sa * v[i].re);
#pragma unroll
      for (int col = 0; col < 6; col++)
#pragma unroll
        for (int row = col + 1; row < 6; row++) {
          const int kk = tri_index(row, col);
          const cplx<real> l((real)m[6 + 2 * kk], (real)m[6 + 2 * kk + 1]);
          cmac(o[row], l, v[col]);        // H[row][col] = L
          cmac_conj(o[col], l, v[row]);   // H[col][row] = conj(L)
        }
    } else {
#pragma unroll
      for (int i = 0; i < 6; i++) o[i] = cplx<real>((real)0, (real)0);
#pragma unroll
      for (int i = 0; i < 6; i++) {
        creal m[12];
        load_reals<creal, 12>(m, (const creal *)Av, chi * 72 + i * 12, 144, Vh, parity, cb);
#pragma unroll
        for (int j = 0; j < 6; j++) {
          const cplx<real> b((real)m[2 * j], (real)m[2 * j + 1]);
          if (MODE == CLOVER_INVERSE) cmac(o[i], b, v[j]);
          else cmac_conj(o[j], b, v[i]);
        }
      }
    }
  }
  const real k = (real)k_;
  if (HAS_X) {
    cplx<real> xs[12];
    load_scaled<Store, 12, false>(xs, x, x_norm, Vh, cb);
#pragma unroll
    for (int c = 0; c < 12; c++) res[c] = cplx<real>(xs[c].re + k * res[c].re, xs[c].im + k * res[c].im);
  } else {
#pragma unroll
    for (int c = 0; c < 12; c++) res[c] = cplx<real>(k * res[c].re, k * res[c].im);
  }
  Store::template store<12>(out, out_norm, Vh, cb, res);
}

template <typename Store, int MODE>
static void launch_clover(SpinorField &out, const SpinorField &in, const CloverField &cl, int parity, double a, const SpinorField *x, double k) {
  const int nb = div_up(in.Vh, 128);
  cudaStream_t s = rt().compute;
  if (x) clover_apply_kernel<Store, MODE, true><<<nb, 128, 0, s>>>(out.parity_ptr(0), out.parity_norm(0), in.parity_ptr(0), in.parity_norm(0), x->parity_ptr(0),
                                                                     x->parity_norm(0), cl.C, cl.Ainv, in.Vh, parity, a, k);
  else clover_apply_kernel<Store, MODE, false><<<nb, 128, 0, s>>>(out.parity_ptr(0), out.parity_norm(0), in.parity_ptr(0), in.parity_norm(0), nullptr, nullptr, cl.C,
                                                                    cl.Ainv, in.Vh, parity, a, k);
  QB_CHECK_LAUNCH();
}

template <typename Store>
static void launch_clover_mode(SpinorField &out, const SpinorField &in, const CloverField &cl, int parity, CloverMode mode, double a, const SpinorField *x, double k) {
  if (mode == CLOVER_DIRECT) launch_clover<Store, CLOVER_DIRECT>(out, in, cl, parity, a, x, k);
  else if (mode == CLOVER_INVERSE) launch_clover<Store, CLOVER_INVERSE>(out, in, cl, parity, a, x, k);
  else launch_clover<Store, CLOVER_INVERSE_ADJ>(out, in, cl, parity, a, x, k);
}

void clover_apply(SpinorField &out, const SpinorField &in, const CloverField &cl, int parity, CloverMode mode, double a, const SpinorField *x, double k) {
  if (in.nparity != 1 || out.nparity != 1 || (x && x->nparity != 1)) QB_ERROR("clover_apply works on single-parity fields");
  if (out.prec != in.prec || out.Vh != in.Vh || in.Vh != cl.Vh || (x && (x->prec != in.prec || x->Vh != in.Vh))) QB_ERROR("clover_apply: field mismatch");
  if ((in.prec == PREC_DOUBLE) != (cl.prec == PREC_DOUBLE)) QB_ERROR("clover_apply: clover copy in the wrong precision");
  if (in.prec == PREC_DOUBLE) launch_clover_mode<StoreD>(out, in, cl, parity, mode, a, x, k);
  else if (in.prec == PREC_SINGLE) launch_clover_mode<StoreS>(out, in, cl, parity, mode, a, x, k);
  else launch_clover_mode<StoreH>(out, in, cl, parity, mode, a, x, k);
}

}  // namespace qb
