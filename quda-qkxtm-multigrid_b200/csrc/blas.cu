// BLAS-1 / reduction kernels (see blas.h).  Compiled with --extended-lambda.
#include "blas.h"
#include "comm.h"
#include "layout.cuh"
#include "peer_reduce.cuh"

namespace qb {
namespace blas {

unsigned long long flops = 0, bytes = 0;
static bool global_reduction = true;
void set_global_reduction(bool on) { global_reduction = on; }

template <typename real> struct alignas(16) Pack {
  static constexpr int N = 8 / sizeof(real);  // complex numbers per 16 bytes
  cplx<real> c[N];
};

// Field accessors: the BLAS functors below are written once, on "packs" of complex numbers, and run on
//   Acc<double> / Acc<float> : flat arrays of 16-byte packs (1 / 2 complex numbers), any field shape;
//   AccH                     : int16 + norm fields (lib/blas_core.h:12-52 handles them as short4 + norm, M = 6): one pack = one SITE of 12
//                              complex numbers, decoded to fp32 on load, re-normalised and re-quantised on store (fp32 arithmetic, as the
//                              reference does for half precision).
template <typename real_> struct Acc {
  typedef real_ real;
  typedef Pack<real_> pack;
  Pack<real_> *p;
  Acc() : p(nullptr) {}
  explicit Acc(const SpinorField &f) : p((Pack<real_> *)f.v) {}
  static long count(const SpinorField &f) { return f.reals() * (long)sizeof(real_) / 16; }
  __device__ __forceinline__ pack load(long i) const { return p[i]; }
  __device__ __forceinline__ void store(long i, const pack &v) const { p[i] = v; }
};
struct PackH {
  static constexpr int N = 12;
  cplx<float> c[N];
};
struct AccH {
  typedef float real;
  typedef PackH pack;
  char *v; float *norm; long Vh; size_t parity_bytes;
  AccH() : v(nullptr), norm(nullptr), Vh(0), parity_bytes(0) {}
  explicit AccH(const SpinorField &f) : v((char *)f.v), norm(f.norm), Vh(f.Vh), parity_bytes(f.parity_bytes) {
    if (f.ncomplex != 12 || f.nflavor != 1 || f.nbatch != 1) QB_ERROR("blas: half-precision vectors are supported for single fine-grid fields (4 spins x 3 colours) only");
  }
  static long count(const SpinorField &f) { return (long)f.nparity * f.Vh; }
  __device__ __forceinline__ pack load(long i) const {
    const long par = i / Vh, cb = i - par * Vh;
    pack u;
    load_scaled<StoreH, 12, false>(u.c, v + par * parity_bytes, norm + par * Vh, Vh, cb);
    return u;
  }
  __device__ __forceinline__ void store(long i, const pack &u) const {
    const long par = i / Vh, cb = i - par * Vh;
    StoreH::store<12>(v + par * parity_bytes, norm + par * Vh, Vh, cb, u.c);
  }
};

static const int BLOCK = 256;
static int grid_for(long n) {
  const long want = (n + BLOCK - 1) / BLOCK;
  const long cap = (long)rt().num_sms * 8;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

template <typename F> __global__ void __launch_bounds__(BLOCK) ew_kernel(long n, F f) {
  for (long i = (long)blockIdx.x * BLOCK + threadIdx.x; i < n; i += (long)gridDim.x * BLOCK) f(i);
}

// ---- reduction machinery --------------------------------------------------------------------
static const int MAX_RED = 64;           // doubles per reduction (block dot of up to 32 vectors)
static double *d_partial = nullptr;      // [grid][MAX_RED]
static unsigned *d_counter = nullptr;
static double *h_result = nullptr;       // mapped pinned host memory, written by the last block
static double *d_result = nullptr;       // device alias of h_result
static double *d_sum = nullptr;          // device-memory result for the ncclAllReduce fallback
static int max_grid = 0;

void init() {
  if (d_partial) return;
  max_grid = rt().num_sms * 8;
  QB_CUDA(cudaMalloc((void **)&d_partial, sizeof(double) * max_grid * MAX_RED));
  QB_CUDA(cudaMalloc((void **)&d_counter, sizeof(unsigned)));
  QB_CUDA(cudaMemset(d_counter, 0, sizeof(unsigned)));
  QB_CUDA(cudaHostAlloc((void **)&h_result, sizeof(double) * MAX_RED, cudaHostAllocMapped));
  QB_CUDA(cudaHostGetDevicePointer((void **)&d_result, h_result, 0));
  QB_CUDA(cudaMalloc((void **)&d_sum, sizeof(double) * MAX_RED));
}

void end() {
  if (d_partial) cudaFree(d_partial);
  if (d_counter) cudaFree(d_counter);
  if (h_result) cudaFreeHost(h_result);
  if (d_sum) cudaFree(d_sum);
  d_sum = nullptr;
  d_partial = nullptr; d_counter = nullptr; h_result = nullptr; d_result = nullptr;
}

template <int NR, typename F>
__global__ void __launch_bounds__(BLOCK) red_kernel(long n, F f, double *partial, unsigned *counter, double *result, const PeerReduce pr) {
  double acc[NR];
_Pragma("unroll")
  for (int k = 0; k < NR; k++) acc[k] = 0.0;
  for (long i = (long)blockIdx.x * BLOCK + threadIdx.x; i < n; i += (long)gridDim.x * BLOCK) f(i, acc);

  __shared__ double sm[NR][BLOCK / 32];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
_Pragma("unroll")
  for (int k = 0; k < NR; k++) {
    double v = acc[k];
_Pragma("unroll")
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sm[k][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < NR) {
    double v = 0.0;
_Pragma("unroll")
    for (int w = 0; w < BLOCK / 32; w++) v += sm[threadIdx.x][w];
    partial[(long)blockIdx.x * NR + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicInc(counter, gridDim.x - 1) == gridDim.x - 1);
  __syncthreads();
  if (last) {
    // deterministic final pass: fixed order over CTAs
    __shared__ double fin[NR];
    for (int k = warp; k < NR; k += BLOCK / 32) {
      double v = 0.0;
      for (int b = lane; b < (int)gridDim.x; b += 32) v += partial[(long)b * NR + k];
_Pragma("unroll")
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) { if (pr.size > 1) fin[k] = v; else result[k] = v; }
    }
    if (pr.size > 1) {
      // all-reduce over the ranks inside the kernel (peer_reduce.cuh): my sums into every rank's mailbox over NVLink, then the flags,
      // then wait for everybody's flag in my own mailbox and add in rank order
      peer_allreduce_cta(pr, fin, NR);
      if ((int)threadIdx.x < NR) result[threadIdx.x] = fin[threadIdx.x];
    }
  }
}

template <int NR, typename F> static void reduce(double *out, long n, F f) {
  Runtime &r = rt();
  init();
  const int grid = grid_for(n);
  const bool global = global_reduction && r.size > 1;
  PeerReduce pr;   // size 1: the kernel writes this rank's sums into the mapped host result
  if (global && comm_peer_reduce_ready()) {
    // the all-reduce happens inside the kernel over the peers' mailboxes: when the stream is idle the global sums are on the host
    pr = comm_peer_reduce_next();
    red_kernel<NR, F><<<grid, BLOCK, 0, r.compute>>>(n, f, d_partial, d_counter, d_result, pr);
    QB_CHECK_LAUNCH();
    QB_CUDA(cudaStreamSynchronize(r.compute));
    for (int k = 0; k < NR; k++) out[k] = h_result[k];
    return;
  }
  if (global) {
    // no peer mapping: sums stay on the device, ncclAllReduce on the compute stream, one copy to the host, one synchronisation
    red_kernel<NR, F><<<grid, BLOCK, 0, r.compute>>>(n, f, d_partial, d_counter, d_sum, pr);
    QB_CHECK_LAUNCH();
    comm_allreduce_sum_device(d_sum, NR, r.compute);
    QB_CUDA(cudaMemcpyAsync(h_result, d_sum, sizeof(double) * NR, cudaMemcpyDeviceToHost, r.compute));
    QB_CUDA(cudaStreamSynchronize(r.compute));
    for (int k = 0; k < NR; k++) out[k] = h_result[k];
    return;
  }
  red_kernel<NR, F><<<grid, BLOCK, 0, r.compute>>>(n, f, d_partial, d_counter, d_result, pr);
  QB_CHECK_LAUNCH();
  QB_CUDA(cudaStreamSynchronize(r.compute));
  for (int k = 0; k < NR; k++) out[k] = h_result[k];
}

template <typename F> static void elementwise(long n, F f) {
  ew_kernel<F><<<grid_for(n), BLOCK, 0, rt().compute>>>(n, f);
  QB_CHECK_LAUNCH();
}

static void check_same(const SpinorField &a, const SpinorField &b) {
  if (a.prec != b.prec) QB_ERROR("blas: precision mismatch (%d vs %d)", (int)a.prec, (int)b.prec);
  if (a.reals() != b.reals()) QB_ERROR("blas: field length mismatch (%ld vs %ld)", a.reals(), b.reals());
}
#define BY_PREC(field, ...)                                \
  do {                                                     \
    if ((field).prec == PREC_DOUBLE) { typedef double real; typedef Acc<double> FA; __VA_ARGS__ } \
    else if ((field).prec == PREC_SINGLE) { typedef float real; typedef Acc<float> FA; __VA_ARGS__ } \
    else { typedef float real; typedef AccH FA; __VA_ARGS__ }               \
  } while (0)

template <typename real> __host__ __device__ inline cplx<real> cmul(cplx<real> a, cplx<real> b) { return a * b; }

void zero(SpinorField &a) { a.zero(rt().compute); }
void copy(SpinorField &dst, const SpinorField &src) { copy_spinor(dst, src, rt().compute); }

// ---- elementwise ------------------------------------------------------------------------------
void ax(double a, SpinorField &x) {
  BY_PREC(x, const FA X(x); const real A = (real)a;
          elementwise(FA::count(x), [=] __device__(long i) {
            FA::pack v = X.load(i);
_Pragma("unroll")
            for (int k = 0; k < FA::pack::N; k++) { v.c[k].re *= A; v.c[k].im *= A; }
            X.store(i, v);
          }););
  flops += x.reals(); bytes += 2 * x.bytes();
}

void axpby(double a, const SpinorField &x, double b, SpinorField &y) {
  check_same(x, y);
  BY_PREC(x, const FA X(x); const FA Y(y); const real A = (real)a, B = (real)b;
          elementwise(FA::count(x), [=] __device__(long i) {
            FA::pack u = X.load(i), v = Y.load(i);
_Pragma("unroll")
            for (int k = 0; k < FA::pack::N; k++) { v.c[k].re = A * u.c[k].re + B * v.c[k].re; v.c[k].im = A * u.c[k].im + B * v.c[k].im; }
            Y.store(i, v);
          }););
  flops += 3 * x.reals(); bytes += 3 * x.bytes();
}

void axpy(double a, const SpinorField &x, SpinorField &y) {
  check_same(x, y);
  BY_PREC(x, const FA X(x); const FA Y(y); const real A = (real)a;
          elementwise(FA::count(x), [=] __device__(long i) {
            FA::pack u = X.load(i), v = Y.load(i);
_Pragma("unroll")
            for (int k = 0; k < FA::pack::N; k++) { v.c[k].re += A * u.c[k].re; v.c[k].im += A * u.c[k].im; }
            Y.store(i, v);
          }););
  flops += 2 * x.reals(); bytes += 3 * x.bytes();
}

void xpy(const SpinorField &x, SpinorField &y) { axpy(1.0, x, y); }
void mxpy(const SpinorField &x, SpinorField &y) { axpy(-1.0, x, y); }
void xpay(const SpinorField &x, double a, SpinorField &y) { axpby(1.0, x, a, y); }

void caxpy(Complex a, const SpinorField &x, SpinorField &y) {
  check_same(x, y);
  BY_PREC(x, const FA X(x); const FA Y(y); const cplx<real> A((real)a.real(), (real)a.imag());
          elementwise(FA::count(x), [=] __device__(long i) {
            FA::pack u = X.load(i), v = Y.load(i);
_Pragma("unroll")
            for (int k = 0; k < FA::pack::N; k++) cmac(v.c[k], A, u.c[k]);
            Y.store(i, v);
          }););
  flops += 4 * x.reals(); bytes += 3 * x.bytes();
}

void caxpby(Complex a, const SpinorField &x, Complex b, SpinorField &y) {
  check_same(x, y);
  BY_PREC(x, const FA X(x); const FA Y(y);
          const cplx<real> A((real)a.real(), (real)a.imag()), B((real)b.real(), (real)b.imag());
          elementwise(FA::count(x), [=] __device__(long i) {
            FA::pack u = X.load(i), v = Y.load(i);
_Pragma("unroll")
            for (int k = 0; k < FA::pack::N; k++) { cplx<real> t = B * v.c[k]; cmac(t, A, u.c[k]); v.c[k] = t; }
            Y.store(i, v);
          }););
  flops += 7 * x.reals(); bytes += 3 * x.bytes();
}

void cxpaypbz(const SpinorField &x, Complex a, const SpinorField &y, Complex b, SpinorField &z) {
  check_same(x, y); check_same(x, z);
  BY_PREC(x, const FA X(x); const FA Y(y); const FA Z(z);
          const cplx<real> A((real)a.real(), (real)a.imag()), B((real)b.real(), (real)b.imag());
          elementwise(FA::count(x), [=] __device__(long i) {
            FA::pack u = X.load(i), v = Y.load(i), w = Z.load(i);
_Pragma("unroll")
            for (int k = 0; k < FA::pack::N; k++) { cplx<real> t = u.c[k]; cmac(t, A, v.c[k]); cmac(t, B, w.c[k]); w.c[k] = t; }
            Z.store(i, w);
          }););
  flops += 8 * x.reals(); bytes += 4 * x.bytes();
}

void caxpbypz(Complex a, const SpinorField &x, Complex b, const SpinorField &y, SpinorField &z) {
  check_same(x, y); check_same(x, z);
  BY_PREC(x, const FA X(x); const FA Y(y); const FA Z(z);
          const cplx<real> A((real)a.real(), (real)a.imag()), B((real)b.real(), (real)b.imag());
          elementwise(FA::count(x), [=] __device__(long i) {
            FA::pack u = X.load(i), v = Y.load(i), w = Z.load(i);
_Pragma("unroll")
            for (int k = 0; k < FA::pack::N; k++) { cmac(w.c[k], A, u.c[k]); cmac(w.c[k], B, v.c[k]); }
            Z.store(i, w);
          }););
  flops += 8 * x.reals(); bytes += 4 * x.bytes();
}

void caxpbypzYmbw(Complex a, const SpinorField &x, Complex b, SpinorField &y, SpinorField &z, const SpinorField &w) {
  check_same(x, y); check_same(x, z); check_same(x, w);
  BY_PREC(x, const FA X(x); const FA Y(y); const FA Z(z);
          const FA W(w);
          const cplx<real> A((real)a.real(), (real)a.imag()), B((real)b.real(), (real)b.imag()), mB(-(real)b.real(), -(real)b.imag());
          elementwise(FA::count(x), [=] __device__(long i) {
            FA::pack u = X.load(i), v = Y.load(i), zz = Z.load(i), ww = W.load(i);
_Pragma("unroll")
            for (int k = 0; k < FA::pack::N; k++) { cmac(zz.c[k], A, u.c[k]); cmac(zz.c[k], B, v.c[k]); cmac(v.c[k], mB, ww.c[k]); }
            Z.store(i, zz); Y.store(i, v);
          }););
  flops += 12 * x.reals(); bytes += 6 * x.bytes();
}

void cabxpyAx(double a, Complex b, SpinorField &x, SpinorField &y) {
  check_same(x, y);
  BY_PREC(x, const FA X(x); const FA Y(y);
          const real A = (real)a; const cplx<real> AB((real)(a * b.real()), (real)(a * b.imag()));
          elementwise(FA::count(x), [=] __device__(long i) {
            FA::pack u = X.load(i), v = Y.load(i);
_Pragma("unroll")
            for (int k = 0; k < FA::pack::N; k++) { cmac(v.c[k], AB, u.c[k]); u.c[k].re *= A; u.c[k].im *= A; }
            X.store(i, u); Y.store(i, v);
          }););
  flops += 5 * x.reals(); bytes += 4 * x.bytes();
}

void caxpyXmaz(Complex a, SpinorField &x, SpinorField &y, const SpinorField &z) {
  check_same(x, y); check_same(x, z);
  BY_PREC(x, const FA X(x); const FA Y(y); const FA Z(z);
          const cplx<real> A((real)a.real(), (real)a.imag()), mA(-(real)a.real(), -(real)a.imag());
          elementwise(FA::count(x), [=] __device__(long i) {
            FA::pack u = X.load(i), v = Y.load(i), w = Z.load(i);
_Pragma("unroll")
            for (int k = 0; k < FA::pack::N; k++) { cmac(v.c[k], A, u.c[k]); cmac(u.c[k], mA, w.c[k]); }
            X.store(i, u); Y.store(i, v);
          }););
  flops += 8 * x.reals(); bytes += 5 * x.bytes();
}

// first step of a fixed-iteration MR smoother started from the source itself (no copy of b into r, no normalisation pass):
//   x (+)= a b,   r = b - a Ab          [reads b, Ab (and x), writes x, r]
void mrFirstStep(Complex a, const SpinorField &b, const SpinorField &Ab, SpinorField &x, SpinorField &r, bool accumulate) {
  check_same(b, Ab); check_same(b, x); check_same(b, r);
  BY_PREC(b, const FA B(b); const FA AB(Ab); const FA X(x);
          const FA R(r);
          const cplx<real> A((real)a.real(), (real)a.imag()), mA(-(real)a.real(), -(real)a.imag());
          elementwise(FA::count(b), [=] __device__(long i) {
            FA::pack u = B.load(i), w = AB.load(i), v;
            if (accumulate) v = X.load(i);
_Pragma("unroll")
            for (int k = 0; k < FA::pack::N; k++) {
              if (!accumulate) v.c[k] = cplx<real>((real)0, (real)0);
              cmac(v.c[k], A, u.c[k]);
              cmac(u.c[k], mA, w.c[k]);
            }
            X.store(i, v); R.store(i, u);
          }););
  flops += 8 * b.reals(); bytes += (accumulate ? 5 : 4) * b.bytes();
}

// y = a x
void cax(Complex a, const SpinorField &x, SpinorField &y) {
  check_same(x, y);
  BY_PREC(x, const FA X(x); const FA Y(y); const cplx<real> A((real)a.real(), (real)a.imag());
          elementwise(FA::count(x), [=] __device__(long i) {
            FA::pack u = X.load(i), v;
_Pragma("unroll")
            for (int k = 0; k < FA::pack::N; k++) v.c[k] = A * u.c[k];
            Y.store(i, v);
          }););
  flops += 6 * x.reals(); bytes += 2 * x.bytes();
}

// ---- reductions -------------------------------------------------------------------------------
double norm2(const SpinorField &x) {
  double out[1];
  BY_PREC(x, const FA X(x);
          reduce<1>(out, FA::count(x), [=] __device__(long i, double *acc) {
            FA::pack u = X.load(i);
_Pragma("unroll")
            for (int k = 0; k < FA::pack::N; k++) acc[0] += (double)u.c[k].re * u.c[k].re + (double)u.c[k].im * u.c[k].im;
          }););
  flops += 2 * x.reals(); bytes += x.bytes();
  return out[0];
}

double reDotProduct(const SpinorField &x, const SpinorField &y) {
  check_same(x, y);
  double out[1];
  BY_PREC(x, const FA X(x); const FA Y(y);
          reduce<1>(out, FA::count(x), [=] __device__(long i, double *acc) {
            FA::pack u = X.load(i), v = Y.load(i);
_Pragma("unroll")
            for (int k = 0; k < FA::pack::N; k++) acc[0] += (double)u.c[k].re * v.c[k].re + (double)u.c[k].im * v.c[k].im;
          }););
  flops += 2 * x.reals(); bytes += 2 * x.bytes();
  return out[0];
}

// acc[0..1] += conj(u) v
template <typename real> __device__ __forceinline__ void cdot_acc(double *acc, cplx<real> u, cplx<real> v) {
  acc[0] += (double)u.re * v.re + (double)u.im * v.im;
  acc[1] += (double)u.re * v.im - (double)u.im * v.re;
}
template <typename real> __device__ __forceinline__ double norm_c(cplx<real> u) { return (double)u.re * u.re + (double)u.im * u.im; }

Complex cDotProduct(const SpinorField &x, const SpinorField &y) {
  check_same(x, y);
  double out[2];
  BY_PREC(x, const FA X(x); const FA Y(y);
          reduce<2>(out, FA::count(x), [=] __device__(long i, double *acc) {
            FA::pack u = X.load(i), v = Y.load(i);
_Pragma("unroll")
            for (int k = 0; k < FA::pack::N; k++) cdot_acc(acc, u.c[k], v.c[k]);
          }););
  flops += 4 * x.reals(); bytes += 2 * x.bytes();
  return Complex(out[0], out[1]);
}

double3_ cDotProductNormA(const SpinorField &x, const SpinorField &y) {
  check_same(x, y);
  double out[3];
  BY_PREC(x, const FA X(x); const FA Y(y);
          reduce<3>(out, FA::count(x), [=] __device__(long i, double *acc) {
            FA::pack u = X.load(i), v = Y.load(i);
_Pragma("unroll")
            for (int k = 0; k < FA::pack::N; k++) { cdot_acc(acc, u.c[k], v.c[k]); acc[2] += norm_c(u.c[k]); }
          }););
  flops += 6 * x.reals(); bytes += 2 * x.bytes();
  return double3_{out[0], out[1], out[2]};
}

double3_ cDotProductNormB(const SpinorField &x, const SpinorField &y) {
  check_same(x, y);
  double out[3];
  BY_PREC(x, const FA X(x); const FA Y(y);
          reduce<3>(out, FA::count(x), [=] __device__(long i, double *acc) {
            FA::pack u = X.load(i), v = Y.load(i);
_Pragma("unroll")
            for (int k = 0; k < FA::pack::N; k++) { cdot_acc(acc, u.c[k], v.c[k]); acc[2] += norm_c(v.c[k]); }
          }););
  flops += 6 * x.reals(); bytes += 2 * x.bytes();
  return double3_{out[0], out[1], out[2]};
}

double axpyNorm(double a, const SpinorField &x, SpinorField &y) {
  check_same(x, y);
  double out[1];
  BY_PREC(x, const FA X(x); const FA Y(y); const real A = (real)a;
          reduce<1>(out, FA::count(x), [=] __device__(long i, double *acc) {
            FA::pack u = X.load(i), v = Y.load(i);
_Pragma("unroll")
            for (int k = 0; k < FA::pack::N; k++) { v.c[k].re += A * u.c[k].re; v.c[k].im += A * u.c[k].im; acc[0] += norm_c(v.c[k]); }
            Y.store(i, v);
          }););
  flops += 4 * x.reals(); bytes += 3 * x.bytes();
  return out[0];
}

double xmyNorm(const SpinorField &x, SpinorField &y) {
  check_same(x, y);
  double out[1];
  BY_PREC(x, const FA X(x); const FA Y(y);
          reduce<1>(out, FA::count(x), [=] __device__(long i, double *acc) {
            FA::pack u = X.load(i), v = Y.load(i);
_Pragma("unroll")
            for (int k = 0; k < FA::pack::N; k++) { v.c[k].re = u.c[k].re - v.c[k].re; v.c[k].im = u.c[k].im - v.c[k].im; acc[0] += norm_c(v.c[k]); }
            Y.store(i, v);
          }););
  flops += 3 * x.reals(); bytes += 3 * x.bytes();
  return out[0];
}

double caxpyNorm(Complex a, const SpinorField &x, SpinorField &y) {
  check_same(x, y);
  double out[1];
  BY_PREC(x, const FA X(x); const FA Y(y); const cplx<real> A((real)a.real(), (real)a.imag());
          reduce<1>(out, FA::count(x), [=] __device__(long i, double *acc) {
            FA::pack u = X.load(i), v = Y.load(i);
_Pragma("unroll")
            for (int k = 0; k < FA::pack::N; k++) { cmac(v.c[k], A, u.c[k]); acc[0] += norm_c(v.c[k]); }
            Y.store(i, v);
          }););
  flops += 6 * x.reals(); bytes += 3 * x.bytes();
  return out[0];
}

// BiCGStab tail in one pass: x += a p + w r; r -= w t; returns (<r0, r>, |r|^2) of the new r  (5 reads + 2 writes instead of the 7 + 2
// of caxpbypz, caxpyNorm and the next iteration's cDotProduct; the reference fuses the same three, caxpbypzYmbwcDotProductUYNormY,
// lib/inv_bicgstab_quda.cpp)
double3_ bicgstabUpdate(Complex a, const SpinorField &p, Complex w, SpinorField &r, const SpinorField &t, SpinorField &x, const SpinorField &r0) {
  check_same(p, r); check_same(p, t); check_same(p, x); check_same(p, r0);
  double out[3];
  BY_PREC(p, const FA P(p); const FA R(r); const FA T(t); const FA X(x); const FA R0(r0);
          const cplx<real> A((real)a.real(), (real)a.imag()), W((real)w.real(), (real)w.imag()), mW(-(real)w.real(), -(real)w.imag());
          reduce<3>(out, FA::count(p), [=] __device__(long i, double *acc) {
            FA::pack pp = P.load(i), rr = R.load(i), tt = T.load(i), xx = X.load(i), zz = R0.load(i);
_Pragma("unroll")
            for (int k = 0; k < FA::pack::N; k++) {
              cmac(xx.c[k], A, pp.c[k]); cmac(xx.c[k], W, rr.c[k]);
              cmac(rr.c[k], mW, tt.c[k]);
              cdot_acc(acc, zz.c[k], rr.c[k]); acc[2] += norm_c(rr.c[k]);
            }
            X.store(i, xx); R.store(i, rr);
          }););
  flops += 18 * p.reals(); bytes += 7 * p.bytes();
  return double3_{out[0], out[1], out[2]};
}

double cabxpyAxNorm(double a, Complex b, SpinorField &x, SpinorField &y) {
  check_same(x, y);
  double out[1];
  BY_PREC(x, const FA X(x); const FA Y(y);
          const real A = (real)a; const cplx<real> AB((real)(a * b.real()), (real)(a * b.imag()));
          reduce<1>(out, FA::count(x), [=] __device__(long i, double *acc) {
            FA::pack u = X.load(i), v = Y.load(i);
_Pragma("unroll")
            for (int k = 0; k < FA::pack::N; k++) { cmac(v.c[k], AB, u.c[k]); u.c[k].re *= A; u.c[k].im *= A; acc[0] += norm_c(v.c[k]); }
            X.store(i, u); Y.store(i, v);
          }););
  flops += 7 * x.reals(); bytes += 4 * x.bytes();
  return out[0];
}

Complex caxpyDotzy(Complex a, const SpinorField &x, SpinorField &y, const SpinorField &z) {
  check_same(x, y); check_same(x, z);
  double out[2];
  BY_PREC(x, const FA X(x); const FA Y(y); const FA Z(z);
          const cplx<real> A((real)a.real(), (real)a.imag());
          reduce<2>(out, FA::count(x), [=] __device__(long i, double *acc) {
            FA::pack u = X.load(i), v = Y.load(i), w = Z.load(i);
_Pragma("unroll")
            for (int k = 0; k < FA::pack::N; k++) { cmac(v.c[k], A, u.c[k]); cdot_acc(acc, w.c[k], v.c[k]); }
            Y.store(i, v);
          }););
  flops += 8 * x.reals(); bytes += 4 * x.bytes();
  return Complex(out[0], out[1]);
}

double caxpyXmazNormX(Complex a, SpinorField &x, SpinorField &y, const SpinorField &z) {
  check_same(x, y); check_same(x, z);
  double out[1];
  BY_PREC(x, const FA X(x); const FA Y(y); const FA Z(z);
          const cplx<real> A((real)a.real(), (real)a.imag()), mA(-(real)a.real(), -(real)a.imag());
          reduce<1>(out, FA::count(x), [=] __device__(long i, double *acc) {
            FA::pack u = X.load(i), v = Y.load(i), w = Z.load(i);
_Pragma("unroll")
            for (int k = 0; k < FA::pack::N; k++) { cmac(v.c[k], A, u.c[k]); cmac(u.c[k], mA, w.c[k]); acc[0] += norm_c(u.c[k]); }
            X.store(i, u); Y.store(i, v);
          }););
  flops += 10 * x.reals(); bytes += 5 * x.bytes();
  return out[0];
}

Complex xpaycDotzy(const SpinorField &x, double a, SpinorField &y, const SpinorField &z) {
  check_same(x, y); check_same(x, z);
  double out[2];
  BY_PREC(x, const FA X(x); const FA Y(y); const FA Z(z);
          const real A = (real)a;
          reduce<2>(out, FA::count(x), [=] __device__(long i, double *acc) {
            FA::pack u = X.load(i), v = Y.load(i), w = Z.load(i);
_Pragma("unroll")
            for (int k = 0; k < FA::pack::N; k++) { v.c[k].re = u.c[k].re + A * v.c[k].re; v.c[k].im = u.c[k].im + A * v.c[k].im; cdot_acc(acc, w.c[k], v.c[k]); }
            Y.store(i, v);
          }););
  flops += 6 * x.reals(); bytes += 4 * x.bytes();
  return Complex(out[0], out[1]);
}

// ---- block variants -----------------------------------------------------------------------------
static const int MAX_BLOCK_VEC = 32;
template <typename A> struct PtrList { A p[MAX_BLOCK_VEC]; cplx<typename A::real> a[MAX_BLOCK_VEC]; int n; };

template <int NV, typename A> static void block_cdot(double *out, const PtrList<A> &L, const A &Y, long n) {
  const PtrList<A> l = L;
  reduce<2 * NV>(out, n, [=] __device__(long i, double *acc) {
    typename A::pack v = Y.load(i);
_Pragma("unroll")
    for (int j = 0; j < NV; j++) {
      if (j < l.n) {
        typename A::pack u = l.p[j].load(i);
_Pragma("unroll")
        for (int k = 0; k < A::pack::N; k++) cdot_acc(acc + 2 * j, u.c[k], v.c[k]);
      }
    }
  });
}

void cDotProduct(Complex *result, const std::vector<SpinorField *> &x, const SpinorField &y) {
  const int n = (int)x.size();
  if (n == 0) return;
  if (n > MAX_BLOCK_VEC) QB_ERROR("block cDotProduct supports at most %d vectors", MAX_BLOCK_VEC);
  for (auto *f : x) check_same(*f, y);
  double out[2 * MAX_BLOCK_VEC];
  BY_PREC(y, PtrList<FA> L; L.n = n; for (int j = 0; j < n; j++) L.p[j] = FA(*x[j]);
          const FA Y(y); const long np = FA::count(y);
          if (n <= 4) block_cdot<4>(out, L, Y, np);
          else if (n <= 8) block_cdot<8>(out, L, Y, np);
          else if (n <= 16) block_cdot<16>(out, L, Y, np);
          else block_cdot<32>(out, L, Y, np););
  for (int j = 0; j < n; j++) result[j] = Complex(out[2 * j], out[2 * j + 1]);
  flops += 4ull * n * y.reals(); bytes += (unsigned long long)(n + 1) * y.bytes();
}

void caxpy(const Complex *a, const std::vector<SpinorField *> &x, SpinorField &y) {
  const int n = (int)x.size();
  if (n == 0) return;
  if (n > MAX_BLOCK_VEC) QB_ERROR("block caxpy supports at most %d vectors", MAX_BLOCK_VEC);
  for (auto *f : x) check_same(*f, y);
  BY_PREC(y, PtrList<FA> L; L.n = n;
          for (int j = 0; j < n; j++) { L.p[j] = FA(*x[j]); L.a[j] = cplx<real>((real)a[j].real(), (real)a[j].imag()); }
          const FA Y(y);
          elementwise(FA::count(y), [=] __device__(long i) {
            FA::pack v = Y.load(i);
            for (int j = 0; j < L.n; j++) {
              FA::pack u = L.p[j].load(i);
_Pragma("unroll")
              for (int k = 0; k < FA::pack::N; k++) cmac(v.c[k], L.a[j], u.c[k]);
            }
            Y.store(i, v);
          }););
  flops += 4ull * n * y.reals(); bytes += (unsigned long long)(n + 2) * y.bytes();
}

}  // namespace blas
}  // namespace qb
