// Multi-right-hand-side (block) BLAS and the batched BiCGStab used to generate the near-null vectors of a coarse level.
//
// The reference generates the n_vec near-null vectors of every level one after the other: n_vec BiCGStab solves of
// M x = 0 from random initial guesses (lib/multigrid.cpp:693-779).  On a coarse level every iteration of every solve
// streams the whole coarse operator (166 KB per site for n_vec = 24).  Here all n_vec solves of a coarse level advance in
// lock-step on block fields ([site][component pair][rhs]), so that the operator is applied by the tensor-core multi-RHS
// kernel (coarse_mrhs.cu) and its links are read once per iteration for all right-hand sides.  Each right-hand side keeps
// its own Krylov scalars and its own convergence test; a converged column is frozen (its coefficients become zero).
// The arithmetic per column is that of BiCGStab::operator() in solver.cu.
#include <chrono>
#include <cmath>
#include <complex>
#include <functional>
#include <cstdlib>
#include <vector>
#include "coarse.h"
#include "comm.h"
#include "multigrid.h"
#include "peer_reduce.cuh"

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
// one CTA of 3 * MAXR threads; pr.size > 1: the sums are all-reduced over the ranks inside the kernel (peer_reduce.cuh)
__global__ void block_cdot_final_kernel(const double *partial, int nblk, int R, double *out, const PeerReduce pr) {
  __shared__ double fin[3 * MAXR];
  const int t = threadIdx.x;
  if (t < 3 * R) {
    double s = 0;
    for (int b = 0; b < nblk; b++) s += partial[(size_t)b * R * 3 + t];
    fin[t] = s;
  }
  if (pr.size > 1) peer_allreduce_cta(pr, fin, 3 * R);
  if (t < 3 * R) out[t] = fin[t];
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
    // global sums on partitioned lattices: inside the kernel over the peer mailboxes, else ncclAllReduce on the stream
    PeerReduce pr;
    const bool global = rt().size > 1;
    if (global && comm_peer_reduce_ready()) pr = comm_peer_reduce_next();
    block_cdot_final_kernel<<<1, 3 * MAXR, 0, s>>>(partial, nblk, R, result_d, pr);
    QB_CHECK_LAUNCH();
    if (global && pr.size == 1) comm_allreduce_sum_device(result_d, 3 * R, s);
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
  if (!d || !d->pc) return false;
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


// =====================================================================================================================
// Block multigrid: R right-hand sides through the K-cycle together (reference: invertMultiSrcQuda
// include/quda.h:647 + the composite ColorSpinorField).  On the coarse levels the R vectors live side by side in block
// fields, every operator application is ONE launch of the tensor-core multi-RHS kernel (links read once for all R), and
// the Krylov scalars are kept per column; a converged column gets zero coefficients.  The fine level keeps one field per
// right-hand side and runs the single-RHS kernels column by column (the fine hop is HBM-bound on its own vectors).
// The per-column arithmetic is that of MR::operator() / GCR::operator() in solver.cu.
// =====================================================================================================================
namespace {

typedef std::complex<double> Cx;
typedef std::vector<char> Mask;

// optional wall-clock profile of the block cycle (QUDA_B200_MG_PROFILE=1): sections bracketed by stream syncs
static double bnow_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static bool block_profile_on() { static int on = -1; if (on < 0) on = (getenv("QUDA_B200_MG_PROFILE") && atoi(getenv("QUDA_B200_MG_PROFILE"))) ? 1 : 0; return on == 1; }
struct BSection {
  double *acc; double t0; bool on;
  explicit BSection(double *a) : acc(a), on(block_profile_on()) { if (on) { cudaStreamSynchronize(rt().compute); t0 = bnow_s(); } }
  ~BSection() { if (on) { cudaStreamSynchronize(rt().compute); *acc += bnow_s() - t0; } }
};

static bool any_active(const Mask &m) { for (char c : m) if (c) return true; return false; }

// ---- vector-set policies -------------------------------------------------------------------------------------------
// R coarse vectors in one block buffer of n4 float4 (one parity or a full field)
struct BlockOps {
  typedef float *Vec;
  int R; long n4; BlockBlas blas_;
  std::vector<Cx> none;
  BlockOps(long n4_, int R_) : R(R_), n4(n4_), blas_(n4_, R_) {}
  Vec make() { Vec v = (Vec)pool_malloc((size_t)n4 * 16); QB_CUDA(cudaMemsetAsync(v, 0, (size_t)n4 * 16, rt().compute)); return v; }
  void release(Vec v) { pool_free(v); }
  void copy(Vec d, Vec s_, const Mask &) { QB_CUDA(cudaMemcpyAsync(d, s_, (size_t)n4 * 16, cudaMemcpyDeviceToDevice, rt().compute)); }
  void zero(Vec d, const Mask &) { QB_CUDA(cudaMemsetAsync(d, 0, (size_t)n4 * 16, rt().compute)); }
  void cdot(Vec x, Vec y, std::vector<Cx> &dot, std::vector<double> &nx, const Mask &) { blas_.cdot(x, y, dot, nx); }
  void axpy(const std::vector<Cx> &a, Vec x, Vec y, const Mask &) { blas_.axpy<0>(a, none, x, nullptr, y); }   // y += a x
  void scale(const std::vector<double> &a, Vec x, const Mask &) {
    std::vector<Cx> c(R);
    for (int r = 0; r < R; r++) c[r] = Cx(a[r], 0.0);
    blas_.axpy<3>(c, none, x, nullptr, x);
  }
  // x *= sc; y += coef x (the scaled x); r2 = |y|^2 per column
  void scale_axpy_norm(const std::vector<double> &sc, const std::vector<Cx> &coef, Vec x, Vec y, std::vector<double> &r2, const Mask &m) {
    std::vector<Cx> dot;
    scale(sc, x, m);
    axpy(coef, x, y, m);
    cdot(y, y, dot, r2, m);
  }
  // beta[i][c] = <v_i, y>_c for i < k, then y -= sum_i beta_i v_i   (one vector at a time: modified Gram-Schmidt)
  void ortho(const std::vector<Vec> &v, int k, Vec y, std::vector<std::vector<Cx>> &beta, const Mask &live) {
    std::vector<Cx> dot, coef(R);
    std::vector<double> nx;
    beta.assign(k, std::vector<Cx>(R));
    for (int i = 0; i < k; i++) {
      blas_.cdot(v[i], y, dot, nx);
      for (int c = 0; c < R; c++) { beta[i][c] = live[c] ? dot[c] : Cx(0, 0); coef[c] = -beta[i][c]; }
      axpy(coef, v[i], y, live);
    }
  }
  // y += sum_i a[i][c] v_i
  void multi_axpy(const std::vector<std::vector<Cx>> &a, const std::vector<Vec> &v, int k, Vec y, const Mask &m) {
    for (int i = 0; i < k; i++) axpy(a[i], v[i], y, m);
  }
};

// R separate fine-grid fields (single precision, full); inactive columns are skipped altogether
// QB_BLOCK_BATCH=0: the R columns of a fine-grid block vector as separate fields and every fine operator application one column at a time
static bool block_batch_on() {
  static int on = -1;
  if (on < 0) { const char *e = getenv("QB_BLOCK_BATCH"); on = (e && atoi(e) == 0) ? 0 : 1; }
  return on == 1;
}

struct FieldOps {
  typedef std::vector<SpinorField *> *Vec;
  int R; long Vh; int nparity;
  std::vector<std::unique_ptr<std::vector<SpinorField *>>> owned;
  // single-parity vectors: the R columns are the members of ONE batch field, so that the fine operator can run on all of them in one
  // launch per hop (links read once per group of 12 columns); BLAS works on the member views
  std::vector<std::pair<Vec, SpinorField *>> batches;
  FieldOps(int R_, long Vh_, int nparity_) : R(R_), Vh(Vh_), nparity(nparity_) {}
  ~FieldOps() {
    for (auto &v : owned) for (SpinorField *f : *v) delete f;
    for (auto &b : batches) delete b.second;
  }
  Vec make() {
    owned.emplace_back(new std::vector<SpinorField *>(R));
    SpinorField *bf = (nparity == 1 && R > 1 && block_batch_on()) ? new SpinorField(Vh, 1, PREC_SINGLE, 4, 3, R) : nullptr;
    for (int r = 0; r < R; r++) {
      SpinorField *f;
      if (bf) { f = new SpinorField(); bf->member(*f, r); }
      else f = new SpinorField(Vh, nparity, PREC_SINGLE);
      (*owned.back())[r] = f;
      blas::zero(*f);
    }
    if (bf) batches.emplace_back(owned.back().get(), bf);
    return owned.back().get();
  }
  SpinorField *batch(Vec v) const {
    for (auto &b : batches) if (b.first == v) return b.second;
    return nullptr;
  }
  void release(Vec) {}
  void copy(Vec d, Vec s_, const Mask &m) { for (int r = 0; r < R; r++) if (m[r]) blas::copy(*(*d)[r], *(*s_)[r]); }
  void zero(Vec d, const Mask &m) { for (int r = 0; r < R; r++) if (m[r]) blas::zero(*(*d)[r]); }
  void cdot(Vec x, Vec y, std::vector<Cx> &dot, std::vector<double> &nx, const Mask &m) {
    dot.assign(R, Cx(0, 0)); nx.assign(R, 0.0);
    for (int r = 0; r < R; r++) if (m[r]) { const blas::double3_ d = blas::cDotProductNormA(*(*x)[r], *(*y)[r]); dot[r] = Cx(d.x, d.y); nx[r] = d.z; }
  }
  void axpy(const std::vector<Cx> &a, Vec x, Vec y, const Mask &m) { for (int r = 0; r < R; r++) if (m[r] && a[r] != Cx(0, 0)) blas::caxpy(a[r], *(*x)[r], *(*y)[r]); }
  void scale(const std::vector<double> &a, Vec x, const Mask &m) { for (int r = 0; r < R; r++) if (m[r] && a[r] != 1.0) blas::ax(a[r], *(*x)[r]); }
  // x *= sc; y += coef x; r2 = |y|^2: the fused kernel of the single-source GCR (cabxpyAxNorm: 4 field streams instead of 7)
  void scale_axpy_norm(const std::vector<double> &sc, const std::vector<Cx> &coef, Vec x, Vec y, std::vector<double> &r2, const Mask &m) {
    r2.assign(R, 0.0);
    for (int r = 0; r < R; r++) if (m[r]) r2[r] = blas::cabxpyAxNorm(sc[r], coef[r], *(*x)[r], *(*y)[r]);
  }
  // per column the fused multi-vector kernels of blas.cu (y is read once for all k dot products / updated once for all k terms)
  void ortho(const std::vector<Vec> &v, int k, Vec y, std::vector<std::vector<Cx>> &beta, const Mask &live) {
    beta.assign(k, std::vector<Cx>(R, Cx(0, 0)));
    if (k == 0) return;
    std::vector<SpinorField *> prev(k);
    std::vector<Cx> bt(k);
    for (int c = 0; c < R; c++) {
      if (!live[c]) continue;
      for (int i = 0; i < k; i++) prev[i] = (*v[i])[c];
      blas::cDotProduct(bt.data(), prev, *(*y)[c]);
      for (int i = 0; i < k; i++) { beta[i][c] = bt[i]; bt[i] = -bt[i]; }
      blas::caxpy(bt.data(), prev, *(*y)[c]);
    }
  }
  void multi_axpy(const std::vector<std::vector<Cx>> &a, const std::vector<Vec> &v, int k, Vec y, const Mask &m) {
    if (k == 0) return;
    std::vector<SpinorField *> vs(k);
    std::vector<Cx> co(k);
    for (int c = 0; c < R; c++) {
      if (!m[c]) continue;
      for (int i = 0; i < k; i++) { vs[i] = (*v[i])[c]; co[i] = a[i][c]; }
      blas::caxpy(co.data(), vs, *(*y)[c]);
    }
  }
};

// ---- lock-step Krylov methods ----------------------------------------------------------------------------------------
template <class Ops> struct BlockKrylov {
  typedef typename Ops::Vec Vec;
  typedef std::function<void(Vec, Vec, const Mask &)> Op;
  Ops &ops;
  std::vector<Vec> p, Ap;
  Vec r = Vec(), y = Vec(), Ar = Vec();
  explicit BlockKrylov(Ops &o) : ops(o) {}
  ~BlockKrylov() {
    for (Vec v : p) ops.release(v);
    for (Vec v : Ap) ops.release(v);
    if (r) ops.release(r);
    if (y) ops.release(y);
    if (Ar) ops.release(Ar);
  }

  // MR::operator() of solver.cu on every active column: niter steps with relaxation omega
  void mr(const Op &A, Vec x, Vec b, int niter, double omega, bool use_init_guess, const Mask &active) {
    const int R = ops.R;
    if (!r) r = ops.make();
    if (!y) y = ops.make();
    if (!Ar) Ar = ops.make();
    std::vector<Cx> dot, a(R);
    std::vector<double> nx, c2, sc(R);
    if (use_init_guess) {
      A(Ar, x, active);
      ops.copy(r, b, active);
      for (int c = 0; c < R; c++) a[c] = Cx(-1, 0);
      ops.axpy(a, Ar, r, active);
    } else {
      ops.copy(r, b, active);
      ops.zero(x, active);
    }
    ops.cdot(r, r, dot, c2, active);
    for (int c = 0; c < R; c++) sc[c] = (active[c] && c2[c] > 0.0) ? 1.0 / sqrt(c2[c]) : 1.0;
    ops.scale(sc, r, active);
    ops.zero(y, active);
    for (int k = 0; k < niter; k++) {
      A(Ar, r, active);
      ops.cdot(Ar, r, dot, nx, active);
      for (int c = 0; c < R; c++) a[c] = (active[c] && c2[c] > 0.0 && nx[c] > 0.0) ? omega * dot[c] / nx[c] : Cx(0, 0);
      ops.axpy(a, r, y, active);                       // y += omega alpha r
      for (int c = 0; c < R; c++) a[c] = -a[c];
      ops.axpy(a, Ar, r, active);                      // r -= omega alpha A r
    }
    for (int c = 0; c < R; c++) a[c] = (active[c] && c2[c] > 0.0) ? Cx(sqrt(c2[c]), 0) : Cx(0, 0);
    ops.axpy(a, y, x, active);
  }

  // GCR::operator() of solver.cu on every active column, zero initial guess.  stop2[c]: target <r,r> of column c.
  // Restarts after nK iterations with the residual recomputed from A (unless the iteration budget is used up).
  // r2 returns the last (recursive, or after a restart true) <r,r>.  Returns the number of iterations of the longest column.
  // delta > 0: hand back to the caller (who recomputes the true residual) as soon as a running column has dropped by that
  // factor since the start of the cycle -- the reliable update of GCR::operator() (reliable_delta).
  int gcr(const Op &A, const Op *K, Vec x, Vec b, const std::vector<double> &stop2, int nK, int maxiter, const Mask &active, std::vector<double> &r2,
          double delta = 0.0) {
    const int R = ops.R;
    if (!r) r = ops.make();
    Mask live = active;
    std::vector<Cx> dot, coef(R);
    std::vector<double> nx, sc(R);
    std::vector<std::vector<Cx>> alpha(nK, std::vector<Cx>(R)), beta((size_t)nK * nK, std::vector<Cx>(R));
    std::vector<std::vector<double>> gamma(nK, std::vector<double>(R));
    ops.copy(r, b, active);
    ops.zero(x, active);
    ops.cdot(r, r, dot, r2, active);
    for (int c = 0; c < R; c++) if (live[c] && !(r2[c] > stop2[c])) live[c] = 0;
    const std::vector<double> r2_start = r2;
    bool reliable = false;
    int total = 0;
    while (any_active(live) && total < maxiter) {
      int k = 0;
      while (k < nK && total < maxiter && any_active(live) && !reliable) {
        while ((int)p.size() <= k) { p.push_back(ops.make()); Ap.push_back(ops.make()); }
        if (K) (*K)(p[k], r, live);
        else ops.copy(p[k], r, live);
        A(Ap[k], p[k], live);
        {
          std::vector<std::vector<Cx>> bk;
          ops.ortho(Ap, k, Ap[k], bk, live);
          for (int i = 0; i < k; i++) beta[(size_t)i * nK + k] = bk[i];
        }
        ops.cdot(Ap[k], r, dot, nx, live);
        for (int c = 0; c < R; c++) {
          if (live[c] && nx[c] > 0.0) { gamma[k][c] = sqrt(nx[c]); alpha[k][c] = dot[c] / gamma[k][c]; sc[c] = 1.0 / gamma[k][c]; }
          else { gamma[k][c] = 1.0; alpha[k][c] = Cx(0, 0); sc[c] = 1.0; }
          coef[c] = -alpha[k][c];
        }
        ops.scale_axpy_norm(sc, coef, Ap[k], r, nx, live);   // Ap_k /= gamma_k; r -= alpha_k Ap_k; |r|^2
        k++; total++;
        for (int c = 0; c < R; c++) if (live[c]) {
          r2[c] = nx[c];
          if (!(r2[c] > stop2[c])) live[c] = 0;
          else if (delta > 0.0 && sqrt(r2[c] / r2_start[c]) < delta) reliable = true;
        }
      }
      // x += sum_i delta_i p_i   (back substitution per column; columns that stopped early have alpha = beta = 0 from then on)
      std::vector<std::vector<Cx>> delta(k, std::vector<Cx>(R));
      for (int c = 0; c < R; c++)
        for (int i = k - 1; i >= 0; i--) {
          Cx d = alpha[i][c];
          for (int j = i + 1; j < k; j++) d -= beta[(size_t)i * nK + j][c] * delta[j][c];
          delta[i][c] = d / gamma[i][c];
        }
      ops.multi_axpy(delta, p, k, x, active);
      if (!any_active(live) || total >= maxiter || reliable) break;
      // restart from the true residual of the columns still running
      A(Ap[0], x, live);
      ops.copy(r, b, live);
      for (int c = 0; c < R; c++) coef[c] = Cx(-1, 0);
      ops.axpy(coef, Ap[0], r, live);
      ops.cdot(r, r, dot, nx, live);
      for (int c = 0; c < R; c++) if (live[c]) { r2[c] = nx[c]; if (!(r2[c] > stop2[c])) live[c] = 0; }
      for (auto &a_ : alpha) a_.assign(R, Cx(0, 0));
      for (auto &b_ : beta) b_.assign(R, Cx(0, 0));
    }
    return total;
  }
};

}  // namespace

// One coarse level (l >= 1) of the block cycle
struct BlockMGLevel {
  MG *mg;
  const DiracCoarse *dres;
  CoarseOperator *op;
  int R, mode, p, q;
  long Vh; int N; size_t pf4;    // float4 per parity block
  std::unique_ptr<CoarseBlockField> x, b, r, t, pview;   // pview: non-owning single-parity view used for pack / unpack of scratch buffers
  std::unique_ptr<BlockOps> ops_full, ops_par;
  std::unique_ptr<BlockKrylov<BlockOps>> smoother, kcycle_gcr, kcycle_gcr_par;
  std::unique_ptr<BlockMhat> mhat;
  float *s1 = nullptr, *s2 = nullptr, *src = nullptr, *rp = nullptr, *s_k = nullptr;   // parity scratch (rp: residual, s_k: source of the K-cycle's even-odd system)
  std::vector<std::unique_ptr<SpinorField>> fx, fb;     // R single fields of this level (transfers to / from the neighbours)
  double t_prof[6] = {0, 0, 0, 0, 0, 0};                // smooth-pre, residual, restrict, coarse solve, prolong, smooth-post

  BlockMGLevel(MG *mg_, int R_, int mode_) : mg(mg_), R(R_), mode(mode_) {
    dres = dynamic_cast<const DiracCoarse *>(mg->matResidual);
    const DiracCoarse *ds = dynamic_cast<const DiracCoarse *>(mg->matSmooth);
    op = dres->op.get();
    if (!op->Xinv) op->compute_xinv();
    op->prepare_mrhs();
    p = ds->p_parity(); q = 1 - p;
    Vh = op->geom.Vh; N = op->N; pf4 = (size_t)Vh * (N / 2) * R;
    x.reset(new CoarseBlockField(Vh, 2, N, R)); b.reset(new CoarseBlockField(Vh, 2, N, R));
    r.reset(new CoarseBlockField(Vh, 2, N, R)); t.reset(new CoarseBlockField(Vh, 2, N, R));
    for (CoarseBlockField *f : {x.get(), b.get(), r.get(), t.get()}) QB_CUDA(cudaMemsetAsync(f->v, 0, f->bytes(), rt().compute));
    ops_full.reset(new BlockOps((long)(2 * pf4), R));
    ops_par.reset(new BlockOps((long)pf4, R));
    smoother.reset(new BlockKrylov<BlockOps>(*ops_par));
    kcycle_gcr.reset(new BlockKrylov<BlockOps>(*ops_full));
    kcycle_gcr_par.reset(new BlockKrylov<BlockOps>(*ops_par));
    mhat.reset(new BlockMhat(*op, p, R, mode));
    s1 = ops_par->make(); s2 = ops_par->make(); src = ops_par->make(); rp = ops_par->make(); s_k = ops_par->make();
    pview.reset(new CoarseBlockField(Vh, 1, N, R, rp));
    for (int c = 0; c < R; c++) {
      fx.emplace_back(new SpinorField(Vh, 2, PREC_SINGLE, 2, op->nvec));
      fb.emplace_back(new SpinorField(Vh, 2, PREC_SINGLE, 2, op->nvec));
      blas::zero(*fx.back()); blas::zero(*fb.back());
    }
  }
  ~BlockMGLevel() { pool_free(s1); pool_free(s2); pool_free(src); pool_free(rp); pool_free(s_k); }
  // source of the even-odd system from a full-lattice right-hand side: out = Xinv_p (b_p - Y_pq Xinv_q b_q)   (DiracCoarsePC::prepare)
  void prepare(float *out, CoarseBlockField &bb) {
    float *bp = par(bb, p), *bq = par(bb, q);
    mhat->launch(s1, nullptr, bq, nullptr, q, false, true, 1.f, 0.f);
    mhat->launch(s2, s1, nullptr, bp, p, true, false, -1.f, 1.f);
    mhat->launch(out, nullptr, s2, nullptr, p, false, true, 1.f, 0.f);
  }
  // x_q = Xinv_q (b_q - Y_qp x_p)   (DiracCoarsePC::reconstruct)
  void reconstruct(CoarseBlockField &xx, CoarseBlockField &bb) {
    mhat->launch(s1, par(xx, p), nullptr, par(bb, q), q, true, false, -1.f, 1.f);
    mhat->launch(par(xx, q), nullptr, s1, nullptr, q, false, true, 1.f, 0.f);
  }
  float *par(CoarseBlockField &f, int parity) const { return f.v + (size_t)parity * pf4 * 4; }

  // out = M in on full block fields
  void full_M(float *out, const float *in) const {
    CoarseMrhsArgs k{};
    k.op = op; k.out = out; k.in_hop = in; k.in_diag = in; k.xpay = nullptr;
    for (int pp = 0; pp < 2; pp++) k.out_poff[pp] = k.hop_poff[pp] = k.diag_poff[pp] = k.xpay_poff[pp] = (long)(pp * pf4);
    k.parity = -1; k.use_y = true; k.use_x = true; k.use_xinv = false; k.a = 1.f; k.b = 0.f; k.R = R; k.mode = mode;
    coarse_apply_mrhs(k);
  }
  // x <- smoother / coarsest solver on M x = b through the even-odd preconditioned system (MG::smooth + DiracCoarse::prepare / reconstruct)
  void smooth(CoarseBlockField &xx, CoarseBlockField &bb, bool coarsest, int niter, bool init_guess, const Mask &active) {
    const MGLevelParam &lp = mg->mp.level[mg->level];
    float *xp = par(xx, p), *xq = par(xx, q), *bp = par(bb, p), *bq = par(bb, q);
    mhat->launch(s1, nullptr, bq, nullptr, q, false, true, 1.f, 0.f);      // Xinv_q b_q
    mhat->launch(s2, s1, nullptr, bp, p, true, false, -1.f, 1.f);          // b_p - Y_pq .
    mhat->launch(src, nullptr, s2, nullptr, p, false, true, 1.f, 0.f);     // Xinv_p .
    BlockKrylov<BlockOps>::Op A = [this](float *o, float *i, const Mask &) { (*mhat)(o, i); };
    if (coarsest) {
      std::vector<Cx> dot; std::vector<double> n2, stop(R), r2;
      ops_par->cdot(src, src, dot, n2, active);
      for (int c = 0; c < R; c++) stop[c] = lp.smoother_tol * lp.smoother_tol * n2[c];
      smoother->gcr(A, nullptr, xp, src, stop, 20, 1000, active, r2);
    } else {
      smoother->mr(A, xp, src, niter, lp.omega, init_guess, active);
    }
    mhat->launch(s1, xp, nullptr, bq, q, true, false, -1.f, 1.f);          // b_q - Y_qp x_p
    mhat->launch(xq, nullptr, s1, nullptr, q, false, true, 1.f, 0.f);      // x_q = Xinv_q .
  }
};

class BlockMG {
 public:
  MG &top;
  int R, mode;
  std::vector<std::unique_ptr<BlockMGLevel>> lv;   // lv[l - 1] <-> level l
  std::vector<std::unique_ptr<SpinorField>> rres;      // R fine residuals (input of the multi-vector restrictor)
  BlockMG(MG &top_, int R_, int mode_) : top(top_), R(R_), mode(mode_) {
    for (MG *m = top.coarse.get(); m; m = m->coarse.get()) lv.emplace_back(new BlockMGLevel(m, R, mode));
  }

  // block cycle on coarse level l (MG::cycle)
  void cycle(int l, CoarseBlockField &x, CoarseBlockField &b, const Mask &active) {
    BlockMGLevel &L = *lv[l - 1];
    MG &m = *L.mg;
    const MGLevelParam &lp = m.mp.level[m.level];
    if (!m.coarse) { BSection s_(&L.t_prof[0]); L.smooth(x, b, true, 0, false, active); return; }
    if (m.pc_coarsen) {
      // this level coarsens its even-odd system: reduce the full system to it exactly, run the single-parity cycle, reconstruct
      L.prepare(L.src, b);
      cycle_pc(l, L.par(x, L.p), L.src, active);
      L.reconstruct(x, b);
      return;
    }
    BlockMGLevel &C = *lv[l];
    if (lp.nu_pre > 0) {
      { BSection s_(&L.t_prof[0]); L.smooth(x, b, false, lp.nu_pre, false, active); }
      BSection s_(&L.t_prof[1]);
      L.full_M(L.r->v, x.v);
      std::vector<Cx> a(R, Cx(-1, 0)), one(R, Cx(1, 0)), none;
      L.ops_full->blas_.axpy<3>(a, none, L.r->v, nullptr, L.r->v);   // r = -M x
      L.ops_full->axpy(one, b.v, L.r->v, active);                    // r += b
    } else {
      L.ops_full->zero(x.v, active);
      L.ops_full->copy(L.r->v, b.v, active);
    }
    { BSection s_(&L.t_prof[2]); restrict_to(l, *L.r, *C.b, active); }
    { BSection s_(&L.t_prof[3]); coarse_solve(l + 1, *C.x, *C.b, active); }
    { BSection s_(&L.t_prof[4]); prolong_add(l, *C.x, x, active); }
    if (lp.nu_post > 0) { BSection s_(&L.t_prof[5]); L.smooth(x, b, false, lp.nu_post, true, active); }
  }

  // MG::cycle_pc on level l >= 1: the cycle on the even-odd system Mhat x_p = src of a level that coarsens it (single-parity block
  // buffers); the next level receives the residual of this parity only and solves its full system
  void cycle_pc(int l, float *xp, float *srcp, const Mask &active) {
    BlockMGLevel &L = *lv[l - 1], &C = *lv[l];
    MG &m = *L.mg;
    const MGLevelParam &lp = m.mp.level[m.level];
    BlockKrylov<BlockOps>::Op A = [&L](float *o, float *i, const Mask &) { (*L.mhat)(o, i); };
    if (lp.nu_pre > 0) {
      { BSection s_(&L.t_prof[0]); L.smoother->mr(A, xp, srcp, lp.nu_pre, lp.omega, false, active); }
      BSection s_(&L.t_prof[1]);
      (*L.mhat)(L.rp, xp);
      std::vector<Cx> a(R, Cx(-1, 0)), one(R, Cx(1, 0)), none;
      L.ops_par->blas_.axpy<3>(a, none, L.rp, nullptr, L.rp);     // rp = -Mhat x
      L.ops_par->axpy(one, srcp, L.rp, active);                   // rp += src
    } else {
      L.ops_par->zero(xp, active);
      L.ops_par->copy(L.rp, srcp, active);
    }
    { BSection s_(&L.t_prof[2]); restrict_to(l, L.rp, *C.b, active, L.p); }
    { BSection s_(&L.t_prof[3]); coarse_solve(l + 1, *C.x, *C.b, active); }
    { BSection s_(&L.t_prof[4]); prolong_add(l, *C.x, xp, active, L.p); }
    if (lp.nu_post > 0) { BSection s_(&L.t_prof[5]); L.smoother->mr(A, xp, srcp, lp.nu_post, lp.omega, true, active); }
  }

  // solve on level l as seen from level l - 1: the level's cycle, wrapped in GCR(10) when the parent runs a K-cycle
  void coarse_solve(int l, CoarseBlockField &x, CoarseBlockField &b, const Mask &active) {
    BlockMGLevel &L = *lv[l - 1];
    MG *parent = l == 1 ? &top : lv[l - 2]->mg;
    if (parent->coarse_solver_pc) {
      // K-cycle on the even-odd system of this level (MG::MG: GCR on the even-odd operator inside prepare / reconstruct)
      const SolverParam &sp = parent->param_coarse_solver;
      BlockKrylov<BlockOps>::Op A = [&L](float *o, float *i, const Mask &) { (*L.mhat)(o, i); };
      BlockKrylov<BlockOps>::Op K = [this, l](float *o, float *i, const Mask &m) { cycle_pc(l, o, i, m); };
      L.prepare(L.s_k, b);
      std::vector<Cx> dot; std::vector<double> n2, stop(R), r2;
      L.ops_par->cdot(L.s_k, L.s_k, dot, n2, active);
      for (int c = 0; c < R; c++) stop[c] = sp.tol * sp.tol * n2[c];
      L.kcycle_gcr_par->gcr(A, &K, L.par(x, L.p), L.s_k, stop, sp.Nkrylov, sp.maxiter, active, r2);
      L.reconstruct(x, b);
      return;
    }
    if (!parent->coarse_solver_gcr) { cycle(l, x, b, active); return; }
    const SolverParam &sp = parent->param_coarse_solver;
    BlockKrylov<BlockOps>::Op A = [&L](float *o, float *i, const Mask &) { L.full_M(o, i); };
    // preconditioner of the K-cycle: this level's cycle on the GCR's own vectors (the cycle only reads its source)
    BlockKrylov<BlockOps>::Op K = [this, l](float *o, float *i, const Mask &m) { cycle_raw(l, o, i, m); };
    std::vector<Cx> dot; std::vector<double> n2, stop(R), r2;
    L.ops_full->cdot(b.v, b.v, dot, n2, active);
    for (int c = 0; c < R; c++) stop[c] = sp.tol * sp.tol * n2[c];
    L.kcycle_gcr->gcr(A, &K, x.v, b.v, stop, sp.Nkrylov, sp.maxiter, active, r2);
  }

  // cycle() on raw full-field buffers of level l (used as the K-cycle preconditioner): borrows the level's x / b objects' storage
  void cycle_raw(int l, float *x, float *b, const Mask &active) {
    BlockMGLevel &L = *lv[l - 1];
    float *sx = L.x->v, *sb = L.b->v;
    L.x->v = x; L.b->v = b;
    cycle(l, *L.x, *L.b, active);
    L.x->v = sx; L.b->v = sb;
  }

  // level l (coarse) -> level l + 1 through the single-vector transfer kernels
  void restrict_to(int l, CoarseBlockField &fine, CoarseBlockField &coarse, const Mask &active) {
    BlockMGLevel &L = *lv[l - 1], &C = *lv[l];
    std::vector<SpinorField *> pf(R), pc(R);
    for (int c = 0; c < R; c++) { pf[c] = L.fb[c].get(); pc[c] = C.fb[c].get(); }
    fine.unpack(pf.data());
    for (int c = 0; c < R; c++) if (active[c]) L.mg->transfer->R(*pc[c], *pf[c]);
    coarse.pack(pc.data());
  }
  // single-parity variants (levels that coarsen their even-odd system): fine_par = block buffer of parity `parity`
  void restrict_to(int l, float *fine_par, CoarseBlockField &coarse, const Mask &active, int parity) {
    BlockMGLevel &L = *lv[l - 1], &C = *lv[l];
    std::vector<SpinorField *> pc(R);
    std::vector<const void *> ptr(R);
    for (int c = 0; c < R; c++) { pc[c] = C.fb[c].get(); ptr[c] = L.fb[c]->parity_ptr(parity); }
    L.pview->v = fine_par;
    L.pview->unpack_ptrs(ptr.data());
    for (int c = 0; c < R; c++) if (active[c]) L.mg->transfer->R(*pc[c], *L.fb[c], parity);
    coarse.pack(pc.data());
  }
  void prolong_add(int l, CoarseBlockField &coarse, float *fine_par, const Mask &active, int parity) {
    BlockMGLevel &L = *lv[l - 1], &C = *lv[l];
    std::vector<SpinorField *> pc(R);
    std::vector<const void *> ptr(R);
    for (int c = 0; c < R; c++) { pc[c] = C.fx[c].get(); ptr[c] = L.fx[c]->parity_ptr(parity); }
    coarse.unpack(pc.data());
    for (int c = 0; c < R; c++) if (active[c]) {
      SpinorField *fo = L.fx[c].get();
      const SpinorField *ci = pc[c];
      L.mg->transfer->P_multi(&fo, &ci, 1, false, parity);
    }
    L.pview->v = L.rp;   // rp is free again at this point of the cycle
    L.pview->pack_ptrs(ptr.data());
    std::vector<Cx> one(R, Cx(1, 0));
    L.ops_par->axpy(one, L.rp, fine_par, active);
  }
  void prolong_add(int l, CoarseBlockField &coarse, CoarseBlockField &fine, const Mask &active) {
    BlockMGLevel &L = *lv[l - 1], &C = *lv[l];
    std::vector<SpinorField *> pf(R), pc(R);
    for (int c = 0; c < R; c++) { pf[c] = L.fx[c].get(); pc[c] = C.fx[c].get(); }
    coarse.unpack(pc.data());
    for (int c = 0; c < R; c++) if (active[c]) L.mg->transfer->P(*pf[c], *pc[c]);
    L.t->pack(pf.data());
    std::vector<Cx> one(R, Cx(1, 0));
    L.ops_full->axpy(one, L.t->v, fine.v, active);
  }

  // ---- level-0 smoother of the even-odd cycle, all columns in lock-step ---------------------------------------------------------
  // scratch block vectors (residual, A r) as batch fields + member views
  std::unique_ptr<SpinorField> sB[2];
  std::vector<std::unique_ptr<SpinorField>> sV[2];
  void ensure_smoother_scratch(long Vh) {
    for (int i = 0; i < 2; i++) {
      if (sB[i] && sB[i]->Vh == Vh) continue;
      sV[i].clear();
      sB[i].reset(new SpinorField(Vh, 1, PREC_SINGLE, 4, 3, R));
      for (int c = 0; c < R; c++) { sV[i].emplace_back(new SpinorField()); sB[i]->member(*sV[i].back(), c); }
    }
  }
  // out = M_pc in on every active column: one batched application (one launch per hop for all columns, links read once per group of 12)
  // when both sides are batch fields and every column is active, else column by column
  void apply_smoother_op(SpinorField *outB, SpinorField *const *out, const SpinorField *inB, SpinorField *const *in, const Mask &active) {
    bool all = true;
    for (int c = 0; c < R; c++) all = all && active[c];
    if (all && outB && inB) { top.matSmooth->M(*outB, *inB); return; }
    for (int c = 0; c < R; c++) if (active[c]) top.matSmooth->M(*out[c], *in[c]);
  }
  // The smoother fast path of MR::operator() (solver.cu: fixed iteration count, source preserved, no normalisation passes, last residual
  // update dropped unless it is wanted) on R columns at once.  keep_residual: sV[0][c] holds b - M_pc x of column c on return.
  void mr_lockstep(std::vector<SpinorField *> &x, SpinorField *xB, std::vector<SpinorField *> &b, SpinorField *bB, int nu, double omega, bool init_guess,
                   bool keep_residual, bool global_reduction, const Mask &active) {
    ensure_smoother_scratch(x[0]->Vh);
    blas::set_global_reduction(global_reduction);
    std::vector<SpinorField *> r(R), Ar(R), rc(R);
    for (int c = 0; c < R; c++) { r[c] = sV[0][c].get(); Ar[c] = sV[1][c].get(); rc[c] = b[c]; }
    SpinorField *rB = sB[0].get(), *ArB = sB[1].get();
    const SpinorField *rcB = bB;
    bool from_b = true, x_valid = init_guess;
    if (init_guess) {
      apply_smoother_op(rB, r.data(), xB, x.data(), active);
      for (int c = 0; c < R; c++) if (active[c]) blas::axpby(1.0, *b[c], -1.0, *r[c]);   // r = b - A x0
      rc = r; rcB = rB; from_b = false;
    }
    for (int k = 0; k < nu; k++) {
      apply_smoother_op(ArB, Ar.data(), rcB, rc.data(), active);
      const bool last = k == nu - 1;
      for (int c = 0; c < R; c++) {
        if (!active[c]) continue;
        const blas::double3_ d = blas::cDotProductNormA(*Ar[c], *rc[c]);
        const Cx alpha = d.z > 0.0 ? omega * Cx(d.x, d.y) / d.z : Cx(0, 0);
        if (last && !keep_residual) {
          if (x_valid) blas::caxpy(alpha, *rc[c], *x[c]);
          else blas::cax(alpha, *rc[c], *x[c]);
        } else if (from_b) blas::mrFirstStep(alpha, *b[c], *Ar[c], *x[c], *r[c], x_valid);
        else blas::caxpyXmaz(alpha, *r[c], *x[c], *Ar[c]);   // x += alpha r; r -= alpha A r
      }
      if (from_b && !(last && !keep_residual)) { rc = r; rcB = rB; from_b = false; }
      x_valid = true;
    }
    if (!x_valid) for (int c = 0; c < R; c++) if (active[c]) blas::zero(*x[c]);
    blas::set_global_reduction(true);
  }

  // level-0 cycle for R fine right-hand sides (single precision, full fields)
  void apply(std::vector<SpinorField *> &x, std::vector<SpinorField *> &b, const Mask &active, SpinorField *xB = nullptr, SpinorField *bB = nullptr) {
    const MGLevelParam &lp = top.mp.level[0];
    BlockMGLevel &L1 = *lv[0];
    std::vector<SpinorField *> pb(R), px(R);
    for (int c = 0; c < R; c++) { pb[c] = L1.fb[c].get(); px[c] = L1.fx[c].get(); }
    std::vector<SpinorField *> act_res, act_pb, act_px, act_x;
    // single-parity fields: the cycle on the even-odd system (MG::cycle_pc): smoother and residual on M_pc directly, transfers restricted
    // to that parity
    const bool pc = x[0]->nparity == 1;
    const int tpar = pc ? top.pc_parity : -1;
    if (pc && !top.pc_coarsen) QB_ERROR("block multigrid: single-parity fields need a hierarchy coarsened on the even-odd system");
    while ((int)rres.size() < R) rres.emplace_back(new SpinorField(top.r->Vh, pc ? 1 : 2, PREC_SINGLE));
    // even-odd cycle with an MR smoother: all columns in lock-step, the operator on batch fields, the residual after pre-smoothing
    // left behind by the smoother's last step (use_solver_residual, as MG::cycle_pc)
    const bool lockstep = pc && block_batch_on() && top.mp.level[0].smoother == INV_MR && x[0]->prec == PREC_SINGLE && b[0]->prec == PREC_SINGLE;
    if (lockstep && lp.nu_pre > 0) {
      BSection s_(&t_prof[0]);
      mr_lockstep(x, xB, b, bB, lp.nu_pre, lp.omega, false, true, lp.global_reduction, active);
    }
    for (int c = 0; c < R; c++) {
      if (!active[c]) continue;
      SpinorField *res = rres[c].get();
      if (lockstep && lp.nu_pre > 0) res = sV[0][c].get();
      else if (lp.nu_pre > 0) {
        { BSection s_(&t_prof[0]); if (pc) (*top.presmoother)(*x[c], *b[c]); else top.smooth(*top.presmoother, *x[c], *b[c]); }
        BSection s_(&t_prof[1]);
        const MR *mr = pc ? dynamic_cast<const MR *>(top.presmoother.get()) : nullptr;
        if (mr && mr->residual()) blas::copy(*rres[c], *mr->residual());   // b - M_pc x left behind by the smoother's last step
        else {
          (pc ? top.matSmooth : top.matResidual)->M(*rres[c], *x[c]);
          blas::axpby(1.0, *b[c], -1.0, *rres[c]);
        }
      } else {
        blas::zero(*x[c]);
        blas::copy(*rres[c], *b[c]);
      }
      act_res.push_back(res); act_pb.push_back(pb[c]); act_px.push_back(px[c]); act_x.push_back(x[c]);
    }
    { BSection s_(&t_prof[2]); top.transfer->R_multi(act_pb.data(), act_res.data(), (int)act_res.size(), tpar); }
    {
      BSection s_(&t_prof[3]);
      L1.b->pack(pb.data());
      coarse_solve(1, *L1.x, *L1.b, active);
      L1.x->unpack(px.data());
    }
    { BSection s_(&t_prof[4]); top.transfer->P_multi(act_x.data(), act_px.data(), (int)act_x.size(), true, tpar); }
    if (lp.nu_post > 0) {
      if (lockstep) { BSection s_(&t_prof[5]); mr_lockstep(x, xB, b, bB, lp.nu_post, lp.omega, true, false, lp.global_reduction, active); }
      else
        for (int c = 0; c < R; c++) if (active[c]) { BSection s_(&t_prof[5]); if (pc) (*top.postsmoother)(*x[c], *b[c]); else top.smooth(*top.postsmoother, *x[c], *b[c]); }
    }
    ncycle++;
  }
  double t_prof[6] = {0, 0, 0, 0, 0, 0};
  long ncycle = 0;
  void print_profile() {
    if (!block_profile_on()) return;
    auto line = [&](int level, const double *t) {
      log_msg(0, "block MG level %d profile (%d rhs, %ld cycles): smooth-pre %.4f s, residual %.4f s, restrict %.4f s, coarse-solve %.4f s, prolong %.4f s, smooth-post %.4f s\n",
              level, R, ncycle, t[0], t[1], t[2], t[3], t[4], t[5]);
    };
    line(1, t_prof);
    for (size_t i = 0; i < lv.size(); i++) line((int)i + 2, lv[i]->t_prof);
  }
};

bool block_mg_supported(const MG &mg, int R, int mode) {
  if (getenv("QB_BLOCK_MG") && atoi(getenv("QB_BLOCK_MG")) == 0) return false;
  if (R < 2 || R > MAXR || !mg.coarse) return false;
  for (const MG *m = mg.coarse.get(); m; m = m->coarse.get()) {
    const DiracCoarse *dr = dynamic_cast<const DiracCoarse *>(m->matResidual), *ds = dynamic_cast<const DiracCoarse *>(m->matSmooth);
    if (!dr || !ds || !ds->pc) return false;
    const InverterType sm = m->mp.level[m->level].smoother;
    if (m->coarse ? sm != INV_MR : (sm != INV_MR && sm != INV_GCR)) return false;   // the coarsest level solves with GCR(20) either way
    const int N = dr->op->N;
    if (N != 16 && N != 32 && N != 48 && N != 64) return false;
    if (coarse_mrhs_max_rhs(N, mode) < R) return false;
  }
  return true;
}

// R solves of M x = b by flexible GCR preconditioned with the block multigrid cycle, all columns in lock-step.
// x, b: full fields in the outer precision.  The Krylov space is single precision (the reference's cuda_prec_sloppy); the true
// residual is recomputed in the outer precision at every restart (defect correction), exactly where GCR::operator() does its
// reliable update.  Returns the iteration count of the longest column; true_res[c] = |b - M x| / |b|.
int block_mg_gcr_solve(MG &mg, const DiracMatrix &mat, const DiracMatrix &matSloppy, std::vector<SpinorField *> &x, std::vector<SpinorField *> &b,
                       const SolverParam &sp, int mode, std::vector<double> &true_res) {
  const int R = (int)x.size();
  BlockMG bmg(mg, R, mode);
  FieldOps ops(R, x[0]->Vh, x[0]->nparity);
  BlockKrylov<FieldOps> kry(ops);
  typedef BlockKrylov<FieldOps>::Op Op;
  Op A = [&](FieldOps::Vec o, FieldOps::Vec i, const Mask &m) {
    SpinorField *oB = ops.batch(o), *iB = ops.batch(i);
    bool all = oB && iB;
    for (int c = 0; c < R; c++) all = all && m[c];
    if (all) { matSloppy(*oB, *iB); return; }   // one launch per hop for all columns
    for (int c = 0; c < R; c++) if (m[c]) matSloppy(*(*o)[c], *(*i)[c]);
  };
  Op K = [&](FieldOps::Vec o, FieldOps::Vec i, const Mask &m) { bmg.apply(*o, *i, m, ops.batch(o), ops.batch(i)); };
  FieldOps::Vec rS = ops.make(), e = ops.make();
  std::vector<std::unique_ptr<SpinorField>> r(R), tmp(R);
  std::vector<double> b2(R), r2(R), stop(R), stop_inner(R), r2_inner;
  Mask live(R, 1);
  true_res.assign(R, 0.0);
  for (int c = 0; c < R; c++) {
    r[c].reset(new_like(*b[c], b[c]->prec)); tmp[c].reset(new_like(*b[c], b[c]->prec));
    blas::copy(*r[c], *b[c]);
    blas::zero(*x[c]);
    b2[c] = r2[c] = blas::norm2(*b[c]);
    stop[c] = sp.tol * sp.tol * b2[c];
    if (!(b2[c] > 0.0)) live[c] = 0;
  }
  int total = 0;
  while (any_active(live) && total < sp.maxiter) {
    for (int c = 0; c < R; c++) if (live[c]) { blas::copy(*(*rS)[c], *r[c]); stop_inner[c] = stop[c]; }
    const int budget = std::min(sp.Nkrylov, sp.maxiter - total);
    total += kry.gcr(A, &K, e, rS, stop_inner, sp.Nkrylov, budget, live, r2_inner, sp.delta);
    for (int c = 0; c < R; c++) {
      if (!live[c]) continue;
      blas::copy(*tmp[c], *(*e)[c]);          // to the outer precision
      blas::xpy(*tmp[c], *x[c]);
      mat(*r[c], *x[c]);
      r2[c] = blas::xmyNorm(*b[c], *r[c]);    // r = b - M x
      if (!(r2[c] > stop[c])) live[c] = 0;
    }
  }
  for (int c = 0; c < R; c++) true_res[c] = b2[c] > 0.0 ? sqrt(r2[c] / b2[c]) : 0.0;
  bmg.print_profile();
  return total;
}

}  // namespace qb
