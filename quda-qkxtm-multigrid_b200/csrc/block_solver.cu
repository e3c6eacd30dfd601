// Multi-right-hand-side (block) BLAS and the batched BiCGStab used to generate the near-null vectors of a coarse level.
//
// The reference generates the n_vec near-null vectors of every level one after the other: n_vec BiCGStab solves of
// M x = 0 from random initial guesses (lib/multigrid.cpp:693-779).  On a coarse level every iteration of every solve
// streams the whole coarse operator (166 KB per site for n_vec = 24).  Here all n_vec solves of a coarse level advance in
// lock-step on block fields ([site][component pair][rhs]), so that the operator is applied by the tensor-core multi-RHS
// kernel (coarse_mrhs.cu) and its links are read once per iteration for all right-hand sides.  Each right-hand side keeps
// its own Krylov scalars and its own convergence test; a converged column is frozen (its coefficients become zero).
// The arithmetic per column is that of BiCGStab::operator() in solver.cu.
#include <cmath>
#include <complex>
#include <cstdlib>
#include <vector>
#include "coarse.h"
#include "multigrid.h"

namespace qb {

namespace {

constexpr int MAXR = 64;
struct Coef { float2 a[MAXR], b[MAXR]; };

// x, y: [n4 / R][R] float4 (two complex numbers each); every thread keeps one column: the stride is a multiple of R.
// partial[block][r] = {sum conj(x) y (re, im), sum |x|^2}
__global__ void block_cdot_kernel(const float4 *x, const float4 *y, long n4, int R, double *partial) {
  extern __shared__ double sh[];  // [blockDim][3]
  const long stride = (long)gridDim.x * blockDim.x;
  double re = 0, im = 0, nx = 0;
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += stride) {
    const float4 a = x[e], b = y[e];
    re += (double)a.x * b.x + (double)a.y * b.y + (double)a.z * b.z + (double)a.w * b.w;
    im += (double)a.x * b.y - (double)a.y * b.x + (double)a.z * b.w - (double)a.w * b.z;
    nx += (double)a.x * a.x + (double)a.y * a.y + (double)a.z * a.z + (double)a.w * a.w;
  }
  sh[threadIdx.x * 3] = re; sh[threadIdx.x * 3 + 1] = im; sh[threadIdx.x * 3 + 2] = nx;
  __syncthreads();
  if (threadIdx.x < R) {  // fixed order: deterministic
    double s0 = 0, s1 = 0, s2 = 0;
    for (int t = threadIdx.x; t < blockDim.x; t += R) { s0 += sh[t * 3]; s1 += sh[t * 3 + 1]; s2 += sh[t * 3 + 2]; }
    double *o = partial + ((size_t)blockIdx.x * R + threadIdx.x) * 3;
    o[0] = s0; o[1] = s1; o[2] = s2;
  }
}
__global__ void block_cdot_final_kernel(const double *partial, int nblk, int R, double *out) {
  const int t = threadIdx.x;
  if (t >= 3 * R) return;
  double s = 0;
  for (int b = 0; b < nblk; b++) s += partial[(size_t)b * R * 3 + t];
  out[t] = s;
}

// op 0: y += a x            op 1: z = x + a y + b z (p = r + a v + b p)        op 2: z += a x + b y         op 3: y = a x
template <int OP>
__global__ void block_axpy_kernel(const Coef c, const float4 *x, const float4 *y, float4 *z, long n4, int R) {
  const long stride = (long)gridDim.x * blockDim.x;
  const int r = threadIdx.x % R;  // blockDim is a multiple of R
  const float2 a = c.a[r], b = c.b[r];
  auto cm = [](float2 s, float4 v) { return make_float4(s.x * v.x - s.y * v.y, s.x * v.y + s.y * v.x, s.x * v.z - s.y * v.w, s.x * v.w + s.y * v.z); };
  for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += stride) {
    if (OP == 0) {
      const float4 t = cm(a, x[e]); float4 o = z[e];
      o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w; z[e] = o;
    } else if (OP == 1) {
      const float4 t = cm(a, y[e]), u = cm(b, z[e]), xx = x[e];
      z[e] = make_float4(xx.x + t.x + u.x, xx.y + t.y + u.y, xx.z + t.z + u.z, xx.w + t.w + u.w);
    } else if (OP == 2) {
      const float4 t = cm(a, x[e]), u = cm(b, y[e]); float4 o = z[e];
      o.x += t.x + u.x; o.y += t.y + u.y; o.z += t.z + u.z; o.w += t.w + u.w; z[e] = o;
    } else {
      z[e] = cm(a, x[e]);
    }
  }
}

struct BlockBlas {
  int R, threads, nblk;
  long n4;
  double *partial = nullptr, *result_d = nullptr;
  std::vector<double> result;
  BlockBlas(long n4_, int R_) : R(R_), n4(n4_) {
    threads = (256 / R) * R;
    if (threads == 0) QB_ERROR("block BLAS: too many right-hand sides");
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    nblk = (int)std::min<long>((n4 + threads - 1) / threads, 4L * nsm);
    partial = (double *)pool_malloc(sizeof(double) * 3 * R * nblk);
    result_d = (double *)pool_malloc(sizeof(double) * 3 * R);
    result.resize(3 * R);
  }
  ~BlockBlas() { pool_free(partial); pool_free(result_d); }
  // dot[r] = <x_r, y_r> = sum conj(x) y,  nx[r] = |x_r|^2
  void cdot(const float *x, const float *y, std::vector<std::complex<double>> &dot, std::vector<double> &nx) {
    cudaStream_t s = rt().compute;
    block_cdot_kernel<<<nblk, threads, sizeof(double) * 3 * threads, s>>>((const float4 *)x, (const float4 *)y, n4, R, partial);
    block_cdot_final_kernel<<<1, 3 * MAXR, 0, s>>>(partial, nblk, R, result_d);
    QB_CHECK_LAUNCH();
    QB_CUDA(cudaMemcpyAsync(result.data(), result_d, sizeof(double) * 3 * R, cudaMemcpyDeviceToHost, s));
    QB_CUDA(cudaStreamSynchronize(s));
    dot.resize(R); nx.resize(R);
    for (int r = 0; r < R; r++) { dot[r] = std::complex<double>(result[3 * r], result[3 * r + 1]); nx[r] = result[3 * r + 2]; }
  }
  template <int OP> void axpy(const std::vector<std::complex<double>> &a, const std::vector<std::complex<double>> &b, const float *x, const float *y, float *z) {
    Coef c;
    for (int r = 0; r < R; r++) {
      c.a[r] = make_float2((float)a[r].real(), (float)a[r].imag());
      c.b[r] = b.empty() ? make_float2(0.f, 0.f) : make_float2((float)b[r].real(), (float)b[r].imag());
    }
    block_axpy_kernel<OP><<<nblk, threads, 0, rt().compute>>>(c, (const float4 *)x, (const float4 *)y, (float4 *)z, n4, R);
    QB_CHECK_LAUNCH();
  }
};

// even-odd preconditioned coarse operator on block fields of parity p:  out = in - Xinv_p Y_pq Xinv_q Y_qp in
struct BlockMhat {
  const CoarseOperator &op;
  int p, R, mode;
  float *t1, *t2;  // parity-q scratch
  size_t bytes;
  BlockMhat(const CoarseOperator &op_, int p_, int R_, int mode_) : op(op_), p(p_), R(R_), mode(mode_) {
    bytes = (size_t)op.geom.Vh * (op.N / 2) * R * 16;
    t1 = (float *)pool_malloc(bytes); t2 = (float *)pool_malloc(bytes);
  }
  ~BlockMhat() { pool_free(t1); pool_free(t2); }
  void launch(float *out, const float *in_hop, const float *in_diag, const float *xpay, int parity, bool use_y, bool use_xinv, float a, float b) const {
    CoarseMrhsArgs k{};
    k.op = &op; k.out = out; k.in_hop = in_hop; k.in_diag = in_diag; k.xpay = xpay;  // all single-parity buffers: offsets stay 0
    k.parity = parity; k.use_y = use_y; k.use_x = false; k.use_xinv = use_xinv; k.a = a; k.b = b; k.R = R; k.mode = mode;
    coarse_apply_mrhs(k);
  }
  void hop(float *out, const float *in, int out_parity) const { launch(out, in, nullptr, nullptr, out_parity, true, false, 1.f, 0.f); }
  void operator()(float *out, const float *in) const {
    const int q = 1 - p;
    hop(t1, in, q);                                                   // Y_qp in
    launch(t2, nullptr, t1, nullptr, q, false, true, 1.f, 0.f);       // Xinv_q .
    hop(t1, t2, p);                                                   // Y_pq .   (t1 reused as a parity-p buffer: same size)
    launch(out, nullptr, t1, in, p, false, true, -1.f, 1.f);          // in - Xinv_p .
  }
};

}  // namespace

bool block_null_vectors_supported(const Dirac *matSmooth, int nvec) {
  if (getenv("QB_BLOCK_SETUP") && atoi(getenv("QB_BLOCK_SETUP")) == 0) return false;
  const DiracCoarse *d = dynamic_cast<const DiracCoarse *>(matSmooth);
  if (!d || !d->pc || d->op->geom.partitioned()) return false;
  const int mode = getenv("QB_BLOCK_SETUP_MODE") ? atoi(getenv("QB_BLOCK_SETUP_MODE")) : 3;
  return coarse_mrhs_max_rhs(d->op->N, mode) >= 1 && nvec >= 2;
}

// x[r]: full coarse fields holding the random initial guesses; on return the approximate null vectors of the smoother's
// operator (M x = 0 solved through the even-odd preconditioned system, exactly as prepare / BiCGStab / reconstruct do for one vector).
// Returns the iteration count of the longest-running column.
int block_null_vectors(const Dirac *matSmooth, std::vector<SpinorField *> &x, int maxiter, double tol) {
  const DiracCoarse *d = dynamic_cast<const DiracCoarse *>(matSmooth);
  CoarseOperator &op = *d->op;
  const int mode = getenv("QB_BLOCK_SETUP_MODE") ? atoi(getenv("QB_BLOCK_SETUP_MODE")) : 3;
  const int Rmax = coarse_mrhs_max_rhs(op.N, mode);
  if (!op.Xinv) op.compute_xinv();
  op.prepare_mrhs();
  const int p = d->p_parity(), q = 1 - p;
  const long Vh = op.geom.Vh;
  const int NKC = op.N / 2;
  int longest = 0;
  for (size_t first = 0; first < x.size(); first += Rmax) {
    const int R = (int)std::min<size_t>(Rmax, x.size() - first);
    const long n4 = Vh * NKC * R;
    const size_t bytes = (size_t)n4 * 16;
    CoarseBlockField xs(Vh, 1, op.N, R), r(Vh, 1, op.N, R), r0(Vh, 1, op.N, R), pp(Vh, 1, op.N, R), v(Vh, 1, op.N, R), t(Vh, 1, op.N, R);
    std::vector<const void *> ptr_p(R), ptr_q(R);
    for (int c = 0; c < R; c++) { ptr_p[c] = x[first + c]->parity_ptr(p); ptr_q[c] = x[first + c]->parity_ptr(q); }
    xs.pack_ptrs(ptr_p.data());
    BlockMhat A(op, p, R, mode);
    BlockBlas blas_(n4, R);
    cudaStream_t s = rt().compute;
    typedef std::complex<double> Cx;
    std::vector<Cx> dot, rho(R, Cx(1, 0)), rho0(R, Cx(1, 0)), alpha(R, Cx(1, 0)), omega(R, Cx(1, 0)), ca(R), cb(R), none;
    std::vector<double> nx, r2(R), stop(R);
    std::vector<char> done(R, 0);
    // r = 0 - A x0
    A(r.v, xs.v);
    for (int c = 0; c < R; c++) ca[c] = Cx(-1, 0);
    blas_.axpy<3>(ca, none, r.v, nullptr, r.v);
    blas_.cdot(r.v, r.v, dot, nx);
    for (int c = 0; c < R; c++) { r2[c] = nx[c]; stop[c] = tol * tol * nx[c]; if (!(nx[c] > 0)) done[c] = 1; }
    QB_CUDA(cudaMemcpyAsync(r0.v, r.v, bytes, cudaMemcpyDeviceToDevice, s));
    QB_CUDA(cudaMemsetAsync(pp.v, 0, bytes, s));
    QB_CUDA(cudaMemsetAsync(v.v, 0, bytes, s));
    int k = 0;
    auto all_done = [&]() { for (int c = 0; c < R; c++) if (!done[c]) return false; return true; };
    while (!all_done() && k < maxiter) {
      blas_.cdot(r0.v, r.v, dot, nx);
      for (int c = 0; c < R; c++) {
        rho0[c] = rho[c]; rho[c] = dot[c];
        if (!done[c] && (std::abs(rho0[c]) == 0.0 || std::abs(omega[c]) == 0.0)) done[c] = 1;
        const Cx beta = done[c] ? Cx(0, 0) : (rho[c] / rho0[c]) * (alpha[c] / omega[c]);
        ca[c] = -beta * omega[c]; cb[c] = beta;
      }
      blas_.axpy<1>(ca, cb, r.v, v.v, pp.v);  // p = r + beta (p - omega v)
      A(v.v, pp.v);
      blas_.cdot(r0.v, v.v, dot, nx);
      for (int c = 0; c < R; c++) {
        if (!done[c] && std::abs(dot[c]) == 0.0) done[c] = 1;
        alpha[c] = done[c] ? Cx(0, 0) : rho[c] / dot[c];
        ca[c] = -alpha[c];
      }
      blas_.axpy<0>(ca, none, v.v, nullptr, r.v);  // s = r - alpha v (kept in r)
      A(t.v, r.v);
      blas_.cdot(t.v, r.v, dot, nx);               // <t, s>, |t|^2
      for (int c = 0; c < R; c++) {
        omega[c] = (done[c] || nx[c] == 0.0) ? Cx(0, 0) : dot[c] / nx[c];
        ca[c] = alpha[c]; cb[c] = omega[c];
      }
      blas_.axpy<2>(ca, cb, pp.v, r.v, xs.v);      // x += alpha p + omega s
      for (int c = 0; c < R; c++) ca[c] = -omega[c];
      blas_.axpy<0>(ca, none, t.v, nullptr, r.v);  // r = s - omega t
      blas_.cdot(r.v, r.v, dot, nx);
      k++;
      for (int c = 0; c < R; c++) {
        if (done[c]) continue;
        r2[c] = nx[c];
        if (r2[c] <= stop[c] || omega[c] == Cx(0, 0)) done[c] = 1;
      }
    }
    longest = std::max(longest, k);
    // reconstruct (b = 0):  x_q = -Xinv_q Y_qp x_p
    xs.unpack_ptrs(ptr_p.data());
    A.hop(t.v, xs.v, q);
    A.launch(r.v, nullptr, t.v, nullptr, q, false, true, -1.f, 0.f);
    r.unpack_ptrs(ptr_q.data());
    QB_CUDA(cudaStreamSynchronize(s));
  }
  return longest;
}

}  // namespace qb
