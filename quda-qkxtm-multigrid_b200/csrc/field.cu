// Field allocation and host <-> device marshalling kernels.
#include "layout.cuh"

namespace qb {

// ---------------------------------------------------------------------------------------------
SpinorField::SpinorField(long Vh_, int nparity_, Prec prec_, int nspin_, int ncolor_, int nbatch_, int nflavor_)
    : prec(prec_), nparity(nparity_), ncomplex(nspin_ * ncolor_), nspin(nspin_), ncolor(ncolor_), Vh(Vh_), nbatch(nbatch_), nflavor(nflavor_) {
  const int sb = prec == PREC_HALF ? 2 : (int)prec;
  if ((ncomplex * 2 * sb) % 16) QB_ERROR("site size %d B is not a multiple of the 16-B plane", ncomplex * 2 * sb);
  if (nbatch < 1 || (nbatch > 1 && prec == PREC_HALF)) QB_ERROR("batched fields are fp32 / fp64 only");
  if (nflavor != 1 && nflavor != 2) QB_ERROR("nflavor must be 1 or 2");
  if (nflavor == 2 && (prec == PREC_HALF || nbatch > 1)) QB_ERROR("flavour-doublet fields are fp32 / fp64 and not batched");
  parity_bytes = (size_t)Vh * ncomplex * 2 * sb * nflavor;
  batch_bytes = parity_bytes * nparity;
  v = pool_malloc(batch_bytes * nbatch);
  if (prec == PREC_HALF) norm = (float *)pool_malloc(sizeof(float) * Vh * nparity);
  owner = true;
}

SpinorField::~SpinorField() {
  if (owner) {
    if (v) pool_free(v);
    if (norm) pool_free(norm);
  }
}

void SpinorField::view_parity(SpinorField &dst, int p) const {
  dst.prec = prec; dst.nparity = 1; dst.ncomplex = ncomplex; dst.nspin = nspin; dst.ncolor = ncolor; dst.Vh = Vh;
  dst.v = parity_ptr(p); dst.norm = parity_norm(p); dst.parity_bytes = parity_bytes; dst.owner = false;
  dst.nbatch = nbatch; dst.batch_bytes = batch_bytes;   // a parity view of a batch is a batch of parity views
  dst.nflavor = nflavor;
}

void SpinorField::member(SpinorField &dst, int c) const {
  if (c < 0 || c >= nbatch) QB_ERROR("batch member %d out of range (%d members)", c, nbatch);
  dst.prec = prec; dst.nparity = nparity; dst.ncomplex = ncomplex; dst.nspin = nspin; dst.ncolor = ncolor; dst.Vh = Vh;
  dst.v = (char *)v + (size_t)c * batch_bytes; dst.norm = nullptr; dst.parity_bytes = parity_bytes; dst.owner = false;
  dst.nbatch = 1; dst.batch_bytes = batch_bytes;
}

void SpinorField::zero(cudaStream_t s) {
  QB_CUDA(cudaMemsetAsync(v, 0, bytes(), s));
  if (norm) QB_CUDA(cudaMemsetAsync(norm, 0, sizeof(float) * Vh * nparity, s));
}

GaugeField::GaugeField(long Vh_, Prec prec_, int recon_) : prec(prec_), recon(recon_), Vh(Vh_) {
  if (recon != 18 && recon != 12 && recon != 8) QB_ERROR("unsupported gauge reconstruct %d", recon);
  QB_CUDA(cudaMalloc(&data, bytes()));
}

GaugeField::~GaugeField() {
  if (data) cudaFree(data);
  for (int d = 0; d < 4; d++)
    if (ghost[d]) cudaFree(ghost[d]);
}

// ---------------------------------------------------------------------------------------------
// gauge import: QDP host order -> planes, with 12/8 compression
// ---------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ void st_real(void *base, long idx, T v);
template <> __device__ __forceinline__ void st_real<double>(void *base, long idx, double v) { ((double *)base)[idx] = v; }

template <typename Host, int STORE_BYTES>
__device__ __forceinline__ void store_link_real(void *block, long stride, int rpp, int k, long site, double val, bool phase) {
  // real number k of the compressed link goes to plane k / rpp, slot k % rpp
  const long idx = ((long)(k / rpp) * stride + site) * rpp + (k % rpp);
  if (STORE_BYTES == 8) ((double *)block)[idx] = val;
  else if (STORE_BYTES == 4) ((float *)block)[idx] = (float)val;
  else {
    const double s = phase ? val * (1.0 / 3.14159265358979323846) : val;
    int q = __double2int_rn(s * 32767.0);
    q = q > 32767 ? 32767 : (q < -32767 ? -32767 : q);
    ((unsigned short *)block)[idx] = (unsigned short)(q + 32768);  // offset-binary, see layout.cuh
  }
}

template <typename Host, int STORE_BYTES, int RECON>
__global__ void import_gauge_kernel(void *dst, const Host *stage, Geom g, long Vh, int order) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 8 * Vh) return;
  const int pm = (int)(t / Vh);  // parity*4 + mu  (matches the destination block order)
  const long cb = t - (long)pm * Vh;
  const int parity = pm >> 2, mu = pm & 3;
  const Host *src = order == GAUGE_ORDER_QDP ? stage + ((long)mu * 2 * Vh + (long)parity * Vh + cb) * 18 : stage + (((long)parity * Vh + cb) * 4 + mu) * 18;
  double m[18];
  if (order == GAUGE_ORDER_CPS) {   // column-row colour order, links divided by the anisotropy on load (gauge_field_order.h:1094-1112)
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) { m[(i * 3 + j) * 2] = (double)src[(j * 3 + i) * 2] / (double)g.aniso; m[(i * 3 + j) * 2 + 1] = (double)src[(j * 3 + i) * 2 + 1] / (double)g.aniso; }
  } else {
#pragma unroll
    for (int k = 0; k < 18; k++) m[k] = (double)src[k];
  }
  const int rpp = gauge_reals_per_plane(STORE_BYTES == 2 ? 2 : STORE_BYTES, RECON);
  void *block = (char *)dst + (size_t)pm * RECON * STORE_BYTES * Vh;
  if (RECON == 18 || RECON == 12) {
#pragma unroll
    for (int k = 0; k < RECON; k++) store_link_real<Host, STORE_BYTES>(block, Vh, rpp, k, cb, m[k], false);
  } else {
    // divide out the scale (anisotropy) and the antiperiodic sign so that a unit-determinant matrix is compressed
    double scale = mu < 3 ? (double)g.aniso : 1.0;
    if (mu == 3 && g.tb_fwd < 0) {
      const long za = cb / g.Xh, zb = za / g.X[1];
      const int tt = (int)(zb / g.X[2]);
      if (tt == g.X[3] - 1) scale = -1.0;
    }
#pragma unroll
    for (int k = 0; k < 18; k++) m[k] *= scale;
    double r[8];
    r[0] = m[2]; r[1] = m[3]; r[2] = m[4]; r[3] = m[5]; r[4] = m[6]; r[5] = m[7];
    r[6] = atan2(m[1], m[0]);
    r[7] = atan2(m[13], m[12]);
#pragma unroll
    for (int k = 0; k < 8; k++) store_link_real<Host, STORE_BYTES>(block, Vh, rpp, k, cb, r[k], k >= 6);
  }
}

template <typename Host>
static void import_gauge_host(GaugeField &gf, void *const *h_gauge, const Geom &geom, cudaStream_t s, HostGaugeOrder order) {
  const long Vh = gf.Vh;
  Host *stage;
  const size_t per_dir = sizeof(Host) * 2 * Vh * 18;
  QB_CUDA(cudaMalloc((void **)&stage, 4 * per_dir));
  if (order == GAUGE_ORDER_QDP) {
    for (int mu = 0; mu < 4; mu++)
      QB_CUDA(cudaMemcpyAsync((char *)stage + mu * per_dir, h_gauge[mu], per_dir, cudaMemcpyHostToDevice, s));
  } else {  // one contiguous host array
    QB_CUDA(cudaMemcpyAsync(stage, (const void *)h_gauge, 4 * per_dir, cudaMemcpyHostToDevice, s));
  }
  const int bs = 256;
  const int nb = div_up(8 * Vh, bs);
  const int sb = gf.store_bytes();
#define LAUNCH(SB, RC) import_gauge_kernel<Host, SB, RC><<<nb, bs, 0, s>>>(gf.data, stage, geom, Vh, (int)order)
#define BY_RECON(SB)                         \
  if (gf.recon == 18) LAUNCH(SB, 18);        \
  else if (gf.recon == 12) LAUNCH(SB, 12);   \
  else LAUNCH(SB, 8)
  if (sb == 8) { BY_RECON(8); }
  else if (sb == 4) { BY_RECON(4); }
  else { BY_RECON(2); }
#undef BY_RECON
#undef LAUNCH
  QB_CHECK_LAUNCH();
  QB_CUDA(cudaStreamSynchronize(s));
  QB_CUDA(cudaFree(stage));
}

void import_gauge(GaugeField &g, void *const *h_gauge, Prec host_prec, const Geom &geom, cudaStream_t s, HostGaugeOrder order) {
  if (host_prec == PREC_DOUBLE) import_gauge_host<double>(g, h_gauge, geom, s, order);
  else if (host_prec == PREC_SINGLE) import_gauge_host<float>(g, h_gauge, geom, s, order);
  else QB_ERROR("host gauge precision %d not supported", (int)host_prec);
}

// export (saveGaugeQuda): reconstruct every link and write the QDP host order
template <typename Store, int RECON, typename Host>
__global__ void export_gauge_kernel(Host *stage, const void *src, Geom g, long Vh, int order) {
  typedef typename Store::real real;
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 8 * Vh) return;
  const int pm = (int)(t / Vh);
  const long cb = t - (long)pm * Vh;
  const int parity = pm >> 2, mu = pm & 3;
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
  Host *dst = order == GAUGE_ORDER_QDP ? stage + ((long)mu * 2 * Vh + (long)parity * Vh + cb) * 18 : stage + (((long)parity * Vh + cb) * 4 + mu) * 18;
  if (order == GAUGE_ORDER_CPS) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++) { dst[(j * 3 + i) * 2] = (Host)(U[i * 3 + j].re * ls * (real)g.aniso); dst[(j * 3 + i) * 2 + 1] = (Host)(U[i * 3 + j].im * ls * (real)g.aniso); }
  } else {
#pragma unroll
    for (int k = 0; k < 9; k++) { dst[2 * k] = (Host)(U[k].re * ls); dst[2 * k + 1] = (Host)(U[k].im * ls); }
  }
}

template <typename Host>
static void export_gauge_host(void *const *h_gauge, const GaugeField &gf, const Geom &geom, cudaStream_t s, HostGaugeOrder order) {
  const long Vh = gf.Vh;
  Host *stage;
  const size_t per_dir = sizeof(Host) * 2 * Vh * 18;
  QB_CUDA(cudaMalloc((void **)&stage, 4 * per_dir));
  const int bs = 256, nb = div_up(8 * Vh, bs);
#define LAUNCH(ST, RC) export_gauge_kernel<ST, RC, Host><<<nb, bs, 0, s>>>(stage, gf.data, geom, Vh, (int)order)
#define BY_RECON(ST)                       \
  if (gf.recon == 18) LAUNCH(ST, 18);      \
  else if (gf.recon == 12) LAUNCH(ST, 12); \
  else LAUNCH(ST, 8)
  if (gf.prec == PREC_DOUBLE) { BY_RECON(StoreD); }
  else if (gf.prec == PREC_SINGLE) { BY_RECON(StoreS); }
  else { BY_RECON(StoreH); }
#undef BY_RECON
#undef LAUNCH
  QB_CHECK_LAUNCH();
  if (order == GAUGE_ORDER_QDP) {
    for (int mu = 0; mu < 4; mu++)
      QB_CUDA(cudaMemcpyAsync(h_gauge[mu], (char *)stage + mu * per_dir, per_dir, cudaMemcpyDeviceToHost, s));
  } else {
    QB_CUDA(cudaMemcpyAsync((void *)h_gauge, stage, 4 * per_dir, cudaMemcpyDeviceToHost, s));
  }
  QB_CUDA(cudaStreamSynchronize(s));
  QB_CUDA(cudaFree(stage));
}

void export_gauge(void *const *h_gauge, const GaugeField &g, Prec host_prec, const Geom &geom, cudaStream_t s, HostGaugeOrder order) {
  if (host_prec == PREC_DOUBLE) export_gauge_host<double>(h_gauge, g, geom, s, order);
  else if (host_prec == PREC_SINGLE) export_gauge_host<float>(h_gauge, g, geom, s, order);
  else QB_ERROR("host gauge precision %d not supported", (int)host_prec);
}

// ---------------------------------------------------------------------------------------------
// ghost links: for a partitioned dimension d the backward hop of sites at x_d = 0 needs U_d at
// x_d = X_d - 1 of the backward neighbour.  Single process / forced self-exchange: that is our own
// last slice.  (Multi-rank exchange is layered on top in comm.cu.)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int slice_site_cb(int mu, int fidx, int slice, int parity, const Geom &g) {
  const int d0 = mu == 0 ? 1 : 0, d1 = mu <= 1 ? 2 : 1, d2 = mu <= 2 ? 3 : 2;
  const int L0 = g.X[d0], L1 = g.X[d1];
  const int f2 = 2 * fidx;
  const int row = f2 / L0;
  const int c = row / L1, b = row - c * L1;
  int a = f2 - row * L0;
  a += (slice + b + c + parity + a) & 1;
  int x[4];
  x[mu] = slice; x[d0] = a; x[d1] = b; x[d2] = c;
  return (((x[3] * g.X[2] + x[2]) * g.X[1] + x[1]) * g.X[0] + x[0]) >> 1;
}

// copies the compressed link planes of the slice x_mu = X_mu-1 into [parity][plane][faceVh]
__global__ void gather_ghost_links_kernel(char *dst, const char *src, Geom g, long Vh, int mu, int planes, int elem_bytes, size_t dir_bytes) {
  const int fv = g.faceVh[mu];
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2L * fv) return;
  const int parity = (int)(t / fv), fidx = (int)(t - (long)parity * fv);
  const int cb = slice_site_cb(mu, fidx, g.X[mu] - 1, parity, g);
  const char *sblock = src + (size_t)(parity * 4 + mu) * dir_bytes;
  char *dblock = dst + (size_t)parity * planes * fv * elem_bytes;
  for (int k = 0; k < planes; k++)
    for (int b = 0; b < elem_bytes; b += 4)
      *(int *)(dblock + ((size_t)k * fv + fidx) * elem_bytes + b) = *(const int *)(sblock + ((size_t)k * Vh + cb) * elem_bytes + b);
}

void gather_ghost_links(GaugeField &gf, const Geom &geom, int mu, void *dst, cudaStream_t s) {
  const int fv = geom.faceVh[mu];
  gather_ghost_links_kernel<<<div_up(2L * fv, 256), 256, 0, s>>>((char *)dst, (const char *)gf.data, geom, gf.Vh, mu, gf.planes(),
                                                                 (int)gf.plane_elem_bytes(), gf.dir_bytes());
  QB_CHECK_LAUNCH();
}

// ---------------------------------------------------------------------------------------------
// spinor import / export.  Host layout [parity][cb][spin][color][re,im] (QUDA_DIRAC_ORDER) or
// [cb][color][spin] (QUDA_QDP_DIRAC_ORDER); optional UKQCD <-> DeGrand-Rossi rotation
// (norm-preserving 1/sqrt2 form, the matrices of lib/copy_color_spinor.cuh:49-92).
// ---------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ void ukqcd_to_dr(cplx<T> *o, const cplx<T> *in) {
  const T k = (T)0.70710678118654752440;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    o[0 * 3 + c] = cplx<T>(-k * (in[1 * 3 + c].re + in[3 * 3 + c].re), -k * (in[1 * 3 + c].im + in[3 * 3 + c].im));
    o[1 * 3 + c] = cplx<T>(k * (in[2 * 3 + c].re + in[0 * 3 + c].re), k * (in[2 * 3 + c].im + in[0 * 3 + c].im));
    o[2 * 3 + c] = cplx<T>(k * (in[3 * 3 + c].re - in[1 * 3 + c].re), k * (in[3 * 3 + c].im - in[1 * 3 + c].im));
    o[3 * 3 + c] = cplx<T>(k * (in[0 * 3 + c].re - in[2 * 3 + c].re), k * (in[0 * 3 + c].im - in[2 * 3 + c].im));
  }
}
template <typename T> __device__ __forceinline__ void dr_to_ukqcd(cplx<T> *o, const cplx<T> *in) {
  const T k = (T)0.70710678118654752440;
#pragma unroll
  for (int c = 0; c < 3; c++) {
    o[0 * 3 + c] = cplx<T>(k * (in[1 * 3 + c].re + in[3 * 3 + c].re), k * (in[1 * 3 + c].im + in[3 * 3 + c].im));
    o[1 * 3 + c] = cplx<T>(-k * (in[2 * 3 + c].re + in[0 * 3 + c].re), -k * (in[2 * 3 + c].im + in[0 * 3 + c].im));
    o[2 * 3 + c] = cplx<T>(k * (in[1 * 3 + c].re - in[3 * 3 + c].re), k * (in[1 * 3 + c].im - in[3 * 3 + c].im));
    o[3 * 3 + c] = cplx<T>(k * (in[2 * 3 + c].re - in[0 * 3 + c].re), k * (in[2 * 3 + c].im - in[0 * 3 + c].im));
  }
}

template <typename Store, typename Host>
__global__ void import_spinor_kernel(void *dst, float *dnorm, const Host *stage, long Vh, long nsites, size_t parity_bytes, int basis, int order, long site_begin = 0) {
  typedef typename Store::real real;
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nsites) return;
  t += site_begin;
  const int parity = (int)(t / Vh);
  const long cb = t - (long)parity * Vh;
  const Host *src = stage + t * 24;
  cplx<real> psi[12];
#pragma unroll
  for (int s = 0; s < 4; s++)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const int k = order == ORDER_SPIN_COLOR ? s * 3 + c : c * 4 + s;
      psi[s * 3 + c] = cplx<real>((real)src[2 * k], (real)src[2 * k + 1]);
    }
  if (basis == BASIS_UKQCD) {
    cplx<real> r[12];
    ukqcd_to_dr(r, psi);
#pragma unroll
    for (int k = 0; k < 12; k++) psi[k] = r[k];
  }
  Store::template store<12>((char *)dst + parity_bytes * parity, dnorm ? dnorm + (long)parity * Vh : nullptr, Vh, cb, psi);
}

template <typename Store, typename Host>
__global__ void export_spinor_kernel(Host *stage, const void *src, const float *snorm, long Vh, long nsites, size_t parity_bytes, int basis, int order, long site_begin = 0) {
  typedef typename Store::real real;
  long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nsites) return;
  t += site_begin;
  const int parity = (int)(t / Vh);
  const long cb = t - (long)parity * Vh;
  cplx<real> psi[12];
  load_scaled<Store, 12, false>(psi, (const char *)src + parity_bytes * parity, snorm ? snorm + (long)parity * Vh : nullptr, Vh, cb);
  if (basis == BASIS_UKQCD) {
    cplx<real> r[12];
    dr_to_ukqcd(r, psi);
#pragma unroll
    for (int k = 0; k < 12; k++) psi[k] = r[k];
  }
  Host *dst = stage + t * 24;
#pragma unroll
  for (int s = 0; s < 4; s++)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const int k = order == ORDER_SPIN_COLOR ? s * 3 + c : c * 4 + s;
      dst[2 * k] = (Host)psi[s * 3 + c].re;
      dst[2 * k + 1] = (Host)psi[s * 3 + c].im;
    }
}

// reusable device staging buffer for host-order data (grown on demand, freed at endQuda)
static void *stage_buf = nullptr;
static size_t stage_bytes = 0;
void *staging(size_t bytes) {
  if (bytes > stage_bytes) {
    if (stage_buf) QB_CUDA(cudaFree(stage_buf));
    QB_CUDA(cudaMalloc(&stage_buf, bytes));
    stage_bytes = bytes;
  }
  return stage_buf;
}
void free_staging() {
  if (stage_buf) cudaFree(stage_buf);
  stage_buf = nullptr;
  stage_bytes = 0;
}

template <typename Host>
static void import_spinor_host(SpinorField &f, const void *h, HostBasis basis, HostSpinorOrder order, cudaStream_t s) {
  if (f.ncomplex != 12) QB_ERROR("host import only for nSpin=4, nColor=3 fields");
  const long nsites = f.Vh * f.nparity * f.nflavor;   // doublet: 2 * nparity blocks of Vh sites, see export_spinor_host
  const size_t hb = sizeof(Host) * nsites * 24;
  Host *stage = (Host *)staging(hb);
  QB_CUDA(cudaMemcpyAsync(stage, h, hb, cudaMemcpyHostToDevice, s));
  const int bs = 128, nb = div_up(nsites, bs);
  if (f.prec == PREC_DOUBLE) import_spinor_kernel<StoreD, Host><<<nb, bs, 0, s>>>(f.v, f.norm, stage, f.Vh, nsites, f.flavor_bytes(), basis, order);
  else if (f.prec == PREC_SINGLE) import_spinor_kernel<StoreS, Host><<<nb, bs, 0, s>>>(f.v, f.norm, stage, f.Vh, nsites, f.flavor_bytes(), basis, order);
  else import_spinor_kernel<StoreH, Host><<<nb, bs, 0, s>>>(f.v, f.norm, stage, f.Vh, nsites, f.flavor_bytes(), basis, order);
  QB_CHECK_LAUNCH();
}

void import_spinor(SpinorField &f, const void *h, Prec host_prec, HostBasis basis, HostSpinorOrder order, cudaStream_t s) {
  if (host_prec == PREC_DOUBLE) import_spinor_host<double>(f, h, basis, order, s);
  else if (host_prec == PREC_SINGLE) import_spinor_host<float>(f, h, basis, order, s);
  else QB_ERROR("host spinor precision %d not supported", (int)host_prec);
}

template <typename Host>
static void export_spinor_host(void *h, const SpinorField &f, HostBasis basis, HostSpinorOrder order, cudaStream_t s) {
  if (f.ncomplex != 12) QB_ERROR("host export only for nSpin=4, nColor=3 fields");
  // a flavour doublet is 2 * nparity consecutive blocks of Vh sites: [parity][flavour][site] on the host and on the device
  const long nsites = f.Vh * f.nparity * f.nflavor;
  const size_t hb = sizeof(Host) * nsites * 24;
  Host *stage = (Host *)staging(hb);
  const int bs = 128, nb = div_up(nsites, bs);
  if (f.prec == PREC_DOUBLE) export_spinor_kernel<StoreD, Host><<<nb, bs, 0, s>>>(stage, f.v, f.norm, f.Vh, nsites, f.flavor_bytes(), basis, order);
  else if (f.prec == PREC_SINGLE) export_spinor_kernel<StoreS, Host><<<nb, bs, 0, s>>>(stage, f.v, f.norm, f.Vh, nsites, f.flavor_bytes(), basis, order);
  else export_spinor_kernel<StoreH, Host><<<nb, bs, 0, s>>>(stage, f.v, f.norm, f.Vh, nsites, f.flavor_bytes(), basis, order);
  QB_CHECK_LAUNCH();
  QB_CUDA(cudaMemcpyAsync(h, stage, hb, cudaMemcpyDeviceToHost, s));
  QB_CUDA(cudaStreamSynchronize(s));
}

void export_spinor(void *h, const SpinorField &f, Prec host_prec, HostBasis basis, HostSpinorOrder order, cudaStream_t s) {
  if (host_prec == PREC_DOUBLE) export_spinor_host<double>(h, f, basis, order, s);
  else if (host_prec == PREC_SINGLE) export_spinor_host<float>(h, f, basis, order, s);
  else QB_ERROR("host spinor precision %d not supported", (int)host_prec);
}

// range variants for the pipelined host path: `stage` is the device copy of the WHOLE host array, only
// sites [begin, begin + count) are converted
template <typename Host>
static void import_range_T(SpinorField &f, const void *stage, HostBasis basis, HostSpinorOrder order, long begin, long count, cudaStream_t s) {
  const int bs = 128, nb = div_up(count, bs);
  const Host *st = (const Host *)stage;
  if (f.prec == PREC_DOUBLE) import_spinor_kernel<StoreD, Host><<<nb, bs, 0, s>>>(f.v, f.norm, st, f.Vh, count, f.parity_bytes, basis, order, begin);
  else if (f.prec == PREC_SINGLE) import_spinor_kernel<StoreS, Host><<<nb, bs, 0, s>>>(f.v, f.norm, st, f.Vh, count, f.parity_bytes, basis, order, begin);
  else import_spinor_kernel<StoreH, Host><<<nb, bs, 0, s>>>(f.v, f.norm, st, f.Vh, count, f.parity_bytes, basis, order, begin);
  QB_CHECK_LAUNCH();
}
template <typename Host>
static void export_range_T(void *stage, const SpinorField &f, HostBasis basis, HostSpinorOrder order, long begin, long count, cudaStream_t s) {
  const int bs = 128, nb = div_up(count, bs);
  Host *st = (Host *)stage;
  if (f.prec == PREC_DOUBLE) export_spinor_kernel<StoreD, Host><<<nb, bs, 0, s>>>(st, f.v, f.norm, f.Vh, count, f.parity_bytes, basis, order, begin);
  else if (f.prec == PREC_SINGLE) export_spinor_kernel<StoreS, Host><<<nb, bs, 0, s>>>(st, f.v, f.norm, f.Vh, count, f.parity_bytes, basis, order, begin);
  else export_spinor_kernel<StoreH, Host><<<nb, bs, 0, s>>>(st, f.v, f.norm, f.Vh, count, f.parity_bytes, basis, order, begin);
  QB_CHECK_LAUNCH();
}
void import_spinor_range(SpinorField &f, const void *stage, Prec host_prec, HostBasis basis, HostSpinorOrder order, long begin, long count, cudaStream_t s) {
  if (host_prec == PREC_DOUBLE) import_range_T<double>(f, stage, basis, order, begin, count, s);
  else if (host_prec == PREC_SINGLE) import_range_T<float>(f, stage, basis, order, begin, count, s);
  else QB_ERROR("host spinor precision %d not supported", (int)host_prec);
}
void export_spinor_range(void *stage, const SpinorField &f, Prec host_prec, HostBasis basis, HostSpinorOrder order, long begin, long count, cudaStream_t s) {
  if (host_prec == PREC_DOUBLE) export_range_T<double>(stage, f, basis, order, begin, count, s);
  else if (host_prec == PREC_SINGLE) export_range_T<float>(stage, f, basis, order, begin, count, s);
  else QB_ERROR("host spinor precision %d not supported", (int)host_prec);
}

// ---------------------------------------------------------------------------------------------
// generic fp32 fields (any nSpin x nColor): host order [parity][cb][component][re,im] <-> planes
// ---------------------------------------------------------------------------------------------
__global__ void generic_reorder_kernel(float4 *planes, float *host, long Vh, int nplanes, int nparity, bool to_device) {
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long)nparity * nplanes * Vh) return;
  const long cb = t % Vh;
  const int pl = (int)((t / Vh) % nplanes), parity = (int)(t / (Vh * nplanes));
  float4 *h = (float4 *)(host + (((size_t)parity * Vh + cb) * nplanes + pl) * 4);
  if (to_device) planes[t] = *h;
  else *h = planes[t];
}

void import_generic(SpinorField &f, const float *h, cudaStream_t s) {
  if (f.prec != PREC_SINGLE) QB_ERROR("import_generic: single precision fields only");
  const size_t hb = f.bytes();
  float *stage = (float *)staging(hb);
  QB_CUDA(cudaMemcpyAsync(stage, h, hb, cudaMemcpyHostToDevice, s));
  const long n = (long)f.nparity * f.planes() * f.Vh;
  generic_reorder_kernel<<<div_up(n, 256), 256, 0, s>>>((float4 *)f.v, stage, f.Vh, f.planes(), f.nparity, true);
  QB_CHECK_LAUNCH();
  QB_CUDA(cudaStreamSynchronize(s));
}

void export_generic(float *h, const SpinorField &f, cudaStream_t s) {
  if (f.prec != PREC_SINGLE) QB_ERROR("export_generic: single precision fields only");
  const size_t hb = f.bytes();
  float *stage = (float *)staging(hb);
  const long n = (long)f.nparity * f.planes() * f.Vh;
  generic_reorder_kernel<<<div_up(n, 256), 256, 0, s>>>((float4 *)f.v, stage, f.Vh, f.planes(), f.nparity, false);
  QB_CHECK_LAUNCH();
  QB_CUDA(cudaMemcpyAsync(h, stage, hb, cudaMemcpyDeviceToHost, s));
  QB_CUDA(cudaStreamSynchronize(s));
}

// ---------------------------------------------------------------------------------------------
// resident copy with precision change (fine fields), plain memcpy otherwise
// ---------------------------------------------------------------------------------------------
template <typename Dst, typename Src>
__global__ void convert_spinor_kernel(void *dst, float *dnorm, const void *src, const float *snorm, long Vh, long nsites, size_t dpb, size_t spb) {
  typedef typename Dst::real dreal;
  typedef typename Src::real sreal;
  const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nsites) return;
  const int parity = (int)(t / Vh);
  const long cb = t - (long)parity * Vh;
  cplx<sreal> a[12];
  load_scaled<Src, 12, false>(a, (const char *)src + spb * parity, snorm ? snorm + (long)parity * Vh : nullptr, Vh, cb);
  cplx<dreal> b[12];
#pragma unroll
  for (int k = 0; k < 12; k++) b[k] = cplx<dreal>((dreal)a[k].re, (dreal)a[k].im);
  Dst::template store<12>((char *)dst + dpb * parity, dnorm ? dnorm + (long)parity * Vh : nullptr, Vh, cb, b);
}

void copy_spinor(SpinorField &dst, const SpinorField &src, cudaStream_t s) {
  if (dst.Vh != src.Vh || dst.nparity != src.nparity || dst.ncomplex != src.ncomplex || dst.nflavor != src.nflavor) QB_ERROR("copy_spinor: geometry mismatch");
  if (dst.v == src.v) return;
  if (dst.prec == src.prec) {
    QB_CUDA(cudaMemcpyAsync(dst.v, src.v, src.bytes(), cudaMemcpyDeviceToDevice, s));
    if (src.norm) QB_CUDA(cudaMemcpyAsync(dst.norm, src.norm, sizeof(float) * src.Vh * src.nparity, cudaMemcpyDeviceToDevice, s));
    return;
  }
  if (src.ncomplex != 12) QB_ERROR("precision-changing copy only for fine fields");
  const long nsites = src.Vh * src.nparity * src.nflavor;
  const int bs = 128, nb = div_up(nsites, bs);
#define CV(D, S) convert_spinor_kernel<D, S><<<nb, bs, 0, s>>>(dst.v, dst.norm, src.v, src.norm, src.Vh, nsites, dst.flavor_bytes(), src.flavor_bytes())
  if (dst.prec == PREC_DOUBLE && src.prec == PREC_SINGLE) CV(StoreD, StoreS);
  else if (dst.prec == PREC_DOUBLE && src.prec == PREC_HALF) CV(StoreD, StoreH);
  else if (dst.prec == PREC_SINGLE && src.prec == PREC_DOUBLE) CV(StoreS, StoreD);
  else if (dst.prec == PREC_SINGLE && src.prec == PREC_HALF) CV(StoreS, StoreH);
  else if (dst.prec == PREC_HALF && src.prec == PREC_DOUBLE) CV(StoreH, StoreD);
  else CV(StoreH, StoreS);
#undef CV
  QB_CHECK_LAUNCH();
}

}  // namespace qb
