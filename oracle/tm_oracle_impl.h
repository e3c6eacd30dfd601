/* TEST INFRASTRUCTURE ONLY (oracle).  Included twice by tm_oracle.c with REAL = double / float
 * and SUF = _d / _f.  CPU restatement of the reference's host verification path for the
 * twisted-mass Wilson operator:
 *   hop        : /root/reference/tests/wilson_dslash_reference.cpp:106-133 (dslashReference)
 *   links      : /root/reference/tests/dslash_util.h:104-124 (gaugeLink), :126-145 (spinorNeighbor)
 *   SU(3) mul  : /root/reference/tests/dslash_util.h:56-93
 *   projectors : /root/reference/tests/wilson_dslash_reference.cpp:21-70 (here: gamma table, P = 1 -/+ gamma)
 *   twist      : /root/reference/tests/wilson_dslash_reference.cpp:233-263
 *   tm_dslash  : :276-293, tm_mat :310-330, tm_matpc :357-408, wil_mat :295-308, wil_matpc :333-354
 * Floating-point operation order follows the reference so that fp64 results agree to the bit
 * (compile with -ffp-contract=off).  Unlike the reference the inputs are never modified in place.
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

/* res(3) = U * v  or  U^dagger * v for one colour vector (re,im interleaved) */
static inline void FN(su3_apply)(REAL *res, const REAL *U, const REAL *v, int dag)
{
  for (int n = 0; n < 3; n++) {
    REAL re = 0, im = 0;
    for (int m = 0; m < 3; m++) {
      REAL ar, ai;
      if (!dag) { ar = U[(n * 3 + m) * 2]; ai = U[(n * 3 + m) * 2 + 1]; }
      else      { ar = U[(m * 3 + n) * 2]; ai = -U[(m * 3 + n) * 2 + 1]; }
      REAL br = v[2 * m], bi = v[2 * m + 1];
      re += ar * br - ai * bi;
      im += ar * bi + ai * br;
    }
    res[2 * n] = re;
    res[2 * n + 1] = im;
  }
}

/* out_parity(x) = sum_{mu} [ P(-/+)_mu U_mu(x) in(x+mu) + P(+/-)_mu U_mu(x-mu)^dag in(x-mu) ] */
void FN(orc_wil_dslash)(REAL *out, REAL *const *gauge, const REAL *in, int parity, int dagger)
{
  const orc_lattice_t *L = &orc_lat;
  const long Vh = L->Vh;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < Vh; i++) {
    int x[4];
    orc_coords(x, i, parity);
    REAL acc[24];
    for (int k = 0; k < 24; k++) acc[k] = 0;
    for (int dir = 0; dir < 8; dir++) {
      const int mu = dir >> 1, back = dir & 1;
      int y[4] = {x[0], x[1], x[2], x[3]};
      y[mu] = (y[mu] + (back ? L->X[mu] - 1 : 1)) % L->X[mu];
      const long j = orc_cb_index(y);
      const REAL *U = back ? gauge[mu] + ((long)(1 - parity) * Vh + j) * 18
                           : gauge[mu] + ((long)parity * Vh + i) * 18;
      const REAL *psi = in + j * 24;
      /* sign = -1 for (1 - gamma), +1 for (1 + gamma); forward non-dagger is (1 - gamma) */
      const int plus = (back + dagger) & 1;
      REAL proj[24], hop[24];
      for (int s = 0; s < 4; s++) {
        const int t = orc_gamma_col[mu][s];
        REAL gr = orc_gamma_val[mu][s][0], gi = orc_gamma_val[mu][s][1];
        if (!plus) { gr = -gr; gi = -gi; }
        for (int c = 0; c < 3; c++) {
          /* same accumulation order as a dense 4x4 projector applied column by column */
          REAL re = 0, im = 0;
          if (t < s) { re += gr * psi[t * 6 + c * 2] - gi * psi[t * 6 + c * 2 + 1];
                       im += gr * psi[t * 6 + c * 2 + 1] + gi * psi[t * 6 + c * 2]; }
          re += psi[s * 6 + c * 2];
          im += psi[s * 6 + c * 2 + 1];
          if (t > s) { re += gr * psi[t * 6 + c * 2] - gi * psi[t * 6 + c * 2 + 1];
                       im += gr * psi[t * 6 + c * 2 + 1] + gi * psi[t * 6 + c * 2]; }
          proj[s * 6 + c * 2] = re;
          proj[s * 6 + c * 2 + 1] = im;
        }
      }
      for (int s = 0; s < 4; s++) FN(su3_apply)(hop + s * 6, U, proj + s * 6, back);
      for (int k = 0; k < 24; k++) acc[k] = acc[k] + hop[k];
    }
    for (int k = 0; k < 24; k++) out[i * 24 + k] = acc[k];
  }
}

/* out = b (1 + i a gamma5) in, gamma5 = diag(+,+,-,-) (DeGrand-Rossi); in-place safe */
void FN(orc_twist)(REAL *out, const REAL *in, REAL a, REAL b, long nsites)
{
#pragma omp parallel for schedule(static)
  for (long i = 0; i < nsites; i++)
    for (int s = 0; s < 4; s++) {
      const REAL a5 = (s >= 2 ? (REAL)-1.0 : (REAL)1.0) * a;
      for (int c = 0; c < 3; c++) {
        const long k = i * 24 + s * 6 + c * 2;
        const REAL re = in[k], im = in[k + 1];
        out[k] = b * (re - a5 * im);
        out[k + 1] = b * (im + a5 * re);
      }
    }
}

/* direct: a = 2 kappa mu f, b = 1;  inverse: a = -2 kappa mu f, b = 1/(1+a^2); dagger flips a */
static void FN(twist_kind)(REAL *out, const REAL *in, int dagger, REAL kappa, REAL mu, int flavor, long n, int inverse)
{
  REAL a, b;
  if (!inverse) { a = (REAL)2.0 * kappa * mu * flavor; b = 1; }
  else { a = (REAL)-2.0 * kappa * mu * flavor; b = (REAL)1.0 / ((REAL)1.0 + a * a); }
  if (dagger) a *= (REAL)-1.0;
  FN(orc_twist)(out, in, a, b, n);
}

void FN(orc_twist_gamma5)(REAL *out, const REAL *in, int dagger, double kappa, double mu, int flavor, long n, int inverse)
{ FN(twist_kind)(out, in, dagger, (REAL)kappa, (REAL)mu, flavor, n, inverse); }

/* y = x + a*y (reference xpay, blas_reference.cpp) */
static void FN(xpay_)(const REAL *x, REAL a, REAL *y, long n)
{
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; i++) y[i] = x[i] + a * y[i];
}

static int FN(symmetric_)(int matpc) { return matpc == ORC_MATPC_EVEN_EVEN || matpc == ORC_MATPC_ODD_ODD; }

/* The single-parity "dslash" of the twisted-mass tests:  A^-1 D  (or D^dag A^-dag for symmetric dagger) */
void FN(orc_tm_dslash)(REAL *out, REAL *const *gauge, const REAL *in, double kappa, double mu, int flavor,
                       int parity, int matpc, int dagger)
{
  const long Vh = orc_lat.Vh;
  if (dagger && FN(symmetric_)(matpc)) {
    REAL *tmp = (REAL *)malloc(sizeof(REAL) * Vh * 24);
    FN(twist_kind)(tmp, in, dagger, (REAL)kappa, (REAL)mu, flavor, Vh, 1);
    FN(orc_wil_dslash)(out, gauge, tmp, parity, dagger);
    free(tmp);
  } else {
    FN(orc_wil_dslash)(out, gauge, in, parity, dagger);
    FN(twist_kind)(out, out, dagger, (REAL)kappa, (REAL)mu, flavor, Vh, 1);
  }
}

/* full operator on [even | odd]:  M = (1 + i a g5) - kappa D */
void FN(orc_tm_mat)(REAL *out, REAL *const *gauge, const REAL *in, double kappa, double mu, int flavor, int dagger)
{
  const long Vh = orc_lat.Vh, V = 2 * Vh;
  REAL *tmp = (REAL *)malloc(sizeof(REAL) * V * 24);
  FN(orc_wil_dslash)(out + Vh * 24, gauge, in, 1, dagger);
  FN(orc_wil_dslash)(out, gauge, in + Vh * 24, 0, dagger);
  FN(twist_kind)(tmp, in, dagger, (REAL)kappa, (REAL)mu, flavor, V, 0);
  FN(xpay_)(tmp, (REAL)-kappa, out, V * 24);
  free(tmp);
}

void FN(orc_wil_mat)(REAL *out, REAL *const *gauge, const REAL *in, double kappa, int dagger)
{
  const long Vh = orc_lat.Vh, V = 2 * Vh;
  FN(orc_wil_dslash)(out + Vh * 24, gauge, in, 1, dagger);
  FN(orc_wil_dslash)(out, gauge, in + Vh * 24, 0, dagger);
  FN(xpay_)(in, (REAL)-kappa, out, V * 24);
}

void FN(orc_wil_matpc)(REAL *out, REAL *const *gauge, const REAL *in, double kappa, int matpc, int dagger)
{
  const long Vh = orc_lat.Vh;
  const int p_out = (matpc == ORC_MATPC_EVEN_EVEN || matpc == ORC_MATPC_EVEN_EVEN_ASYM) ? 0 : 1;
  REAL *tmp = (REAL *)malloc(sizeof(REAL) * Vh * 24);
  FN(orc_wil_dslash)(tmp, gauge, in, 1 - p_out, dagger);
  FN(orc_wil_dslash)(out, gauge, tmp, p_out, dagger);
  FN(xpay_)(in, (REAL)(-kappa * kappa), out, Vh * 24);
  free(tmp);
}

/* even-odd preconditioned operator, all four matpc types, with/without dagger */
void FN(orc_tm_matpc)(REAL *out, REAL *const *gauge, const REAL *in, double kappa, double mu, int flavor,
                      int matpc, int dagger)
{
  const long Vh = orc_lat.Vh;
  const REAL k = (REAL)kappa, m = (REAL)mu;
  const int p_out = (matpc == ORC_MATPC_EVEN_EVEN || matpc == ORC_MATPC_EVEN_EVEN_ASYM) ? 0 : 1;
  REAL *tmp = (REAL *)malloc(sizeof(REAL) * Vh * 24);
  const double kappa2 = -kappa * kappa;
  if (!FN(symmetric_)(matpc)) {
    /* A - kappa^2 D A^-1 D */
    FN(orc_wil_dslash)(tmp, gauge, in, 1 - p_out, dagger);
    FN(twist_kind)(tmp, tmp, dagger, k, m, flavor, Vh, 1);
    FN(orc_wil_dslash)(out, gauge, tmp, p_out, dagger);
    FN(twist_kind)(tmp, in, dagger, k, m, flavor, Vh, 0);
    FN(xpay_)(tmp, (REAL)kappa2, out, Vh * 24);
  } else if (!dagger) {
    /* 1 - kappa^2 A^-1 D A^-1 D */
    FN(orc_wil_dslash)(tmp, gauge, in, 1 - p_out, dagger);
    FN(twist_kind)(tmp, tmp, dagger, k, m, flavor, Vh, 1);
    FN(orc_wil_dslash)(out, gauge, tmp, p_out, dagger);
    FN(twist_kind)(out, out, dagger, k, m, flavor, Vh, 1);
    FN(xpay_)(in, (REAL)kappa2, out, Vh * 24);
  } else {
    /* 1 - kappa^2 D^dag A^-dag D^dag A^-dag */
    REAL *tin = (REAL *)malloc(sizeof(REAL) * Vh * 24);
    FN(twist_kind)(tin, in, dagger, k, m, flavor, Vh, 1);
    FN(orc_wil_dslash)(tmp, gauge, tin, 1 - p_out, dagger);
    FN(twist_kind)(tmp, tmp, dagger, k, m, flavor, Vh, 1);
    FN(orc_wil_dslash)(out, gauge, tmp, p_out, dagger);
    /* the reference un-twists its input in place before the xpay; use the pristine input */
    FN(twist_kind)(tin, tin, dagger, k, m, flavor, Vh, 0);
    FN(xpay_)(tin, (REAL)kappa2, out, Vh * 24);
    free(tin);
  }
  free(tmp);
}


/* ---------------------------------------------------------------------------------------------
 * Non-degenerate twisted-mass doublet: /root/reference/tests/wilson_dslash_reference.cpp
 *   ndegTwistGamma5 :412-445  (1 + i a gamma5 tau3 + b tau1) on the flavour pair, a = +-2 kappa mu, b = -+2 kappa epsilon,
 *                             inverse: times d = 1 / (1 + a^2 - b^2); dagger flips a
 *   tm_ndeg_dslash :461-473, tm_ndeg_matpc :476-541, tm_ndeg_mat :544-587
 * A doublet parity field is [flavour 1 | flavour 2], each Vh x 24 reals; a full doublet field is [even doublet | odd doublet].
 * Same operation order as the reference (fp64 agrees to the bit); unlike the reference the inputs are never modified.
 * ------------------------------------------------------------------------------------------- */
void FN(orc_ndeg_twist)(REAL *out1, REAL *out2, const REAL *in1, const REAL *in2, int dagger, double kappa_, double mu_, double eps_, long n, int inverse)
{
  const REAL kappa = (REAL)kappa_, mu = (REAL)mu_, epsilon = (REAL)eps_;
  REAL a, b, d;
  if (!inverse) { a = (REAL)2.0 * kappa * mu; b = (REAL)-2.0 * kappa * epsilon; d = (REAL)1.0; }
  else { a = (REAL)-2.0 * kappa * mu; b = (REAL)2.0 * kappa * epsilon; d = (REAL)1.0 / ((REAL)1.0 + a * a - b * b); }
  if (dagger) a *= (REAL)-1.0;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; i++) {
    REAL t1[24], t2[24];
    for (int s = 0; s < 4; s++)
      for (int c = 0; c < 3; c++) {
        const REAL a5 = ((s / 2) ? (REAL)-1.0 : (REAL)1.0) * a;
        const long k = i * 24 + s * 6 + c * 2;
        t1[s * 6 + c * 2 + 0] = d * (in1[k + 0] - a5 * in1[k + 1] + b * in2[k + 0]);
        t1[s * 6 + c * 2 + 1] = d * (in1[k + 1] + a5 * in1[k + 0] + b * in2[k + 1]);
        t2[s * 6 + c * 2 + 0] = d * (in2[k + 0] + a5 * in2[k + 1] + b * in1[k + 0]);
        t2[s * 6 + c * 2 + 1] = d * (in2[k + 1] - a5 * in2[k + 0] + b * in1[k + 1]);
      }
    for (int j = 0; j < 24; j++) { out1[i * 24 + j] = t1[j]; out2[i * 24 + j] = t2[j]; }
  }
}

/* out, in: doublet parity fields */
void FN(orc_tm_ndeg_dslash)(REAL *out, REAL *const *gauge, const REAL *in, double kappa, double mu, double eps, int parity, int matpc, int dagger)
{
  const long Vh = orc_lat.Vh, F = Vh * 24;
  if (dagger && FN(symmetric_)(matpc)) {
    REAL *tmp = (REAL *)malloc(sizeof(REAL) * 2 * F);
    FN(orc_ndeg_twist)(tmp, tmp + F, in, in + F, dagger, kappa, mu, eps, Vh, 1);
    FN(orc_wil_dslash)(out, gauge, tmp, parity, dagger);
    FN(orc_wil_dslash)(out + F, gauge, tmp + F, parity, dagger);
    free(tmp);
  } else {
    FN(orc_wil_dslash)(out, gauge, in, parity, dagger);
    FN(orc_wil_dslash)(out + F, gauge, in + F, parity, dagger);
    FN(orc_ndeg_twist)(out, out + F, out, out + F, dagger, kappa, mu, eps, Vh, 1);
  }
}

void FN(orc_tm_ndeg_matpc)(REAL *out, REAL *const *gauge, const REAL *in, double kappa, double mu, double eps, int matpc, int dagger)
{
  const long Vh = orc_lat.Vh, F = Vh * 24;
  const int p_out = (matpc == ORC_MATPC_EVEN_EVEN || matpc == ORC_MATPC_EVEN_EVEN_ASYM) ? 0 : 1, q = 1 - p_out;
  REAL *tmp = (REAL *)malloc(sizeof(REAL) * 2 * F);
  REAL *xin = (REAL *)malloc(sizeof(REAL) * 2 * F);   /* the vector the final xpay adds: in, or the twisted in for the asymmetric types */
  memcpy(xin, in, sizeof(REAL) * 2 * F);
  if (!FN(symmetric_)(matpc)) {
    /* A - kappa^2 D A^-1 D  (both daggers take this branch in the reference) */
    FN(orc_wil_dslash)(tmp, gauge, in, q, dagger);
    FN(orc_wil_dslash)(tmp + F, gauge, in + F, q, dagger);
    FN(orc_ndeg_twist)(tmp, tmp + F, tmp, tmp + F, dagger, kappa, mu, eps, Vh, 1);
    FN(orc_wil_dslash)(out, gauge, tmp, p_out, dagger);
    FN(orc_wil_dslash)(out + F, gauge, tmp + F, p_out, dagger);
    FN(orc_ndeg_twist)(xin, xin + F, in, in + F, dagger, kappa, mu, eps, Vh, 0);
  } else if (!dagger) {
    FN(orc_wil_dslash)(tmp, gauge, in, q, dagger);
    FN(orc_wil_dslash)(tmp + F, gauge, in + F, q, dagger);
    FN(orc_ndeg_twist)(tmp, tmp + F, tmp, tmp + F, dagger, kappa, mu, eps, Vh, 1);
    FN(orc_wil_dslash)(out, gauge, tmp, p_out, dagger);
    FN(orc_wil_dslash)(out + F, gauge, tmp + F, p_out, dagger);
    FN(orc_ndeg_twist)(out, out + F, out, out + F, dagger, kappa, mu, eps, Vh, 1);
  } else {
    FN(orc_ndeg_twist)(tmp, tmp + F, in, in + F, dagger, kappa, mu, eps, Vh, 1);
    FN(orc_wil_dslash)(out, gauge, tmp, q, dagger);
    FN(orc_wil_dslash)(out + F, gauge, tmp + F, q, dagger);
    FN(orc_ndeg_twist)(tmp, tmp + F, out, out + F, dagger, kappa, mu, eps, Vh, 1);
    FN(orc_wil_dslash)(out, gauge, tmp, p_out, dagger);
    FN(orc_wil_dslash)(out + F, gauge, tmp + F, p_out, dagger);
  }
  FN(xpay_)(xin, (REAL)(-kappa * kappa), out, 2 * F);
  free(tmp); free(xin);
}

/* out, in: full doublet fields [even doublet | odd doublet] */
void FN(orc_tm_ndeg_mat)(REAL *out, REAL *const *gauge, const REAL *in, double kappa, double mu, double eps, int dagger)
{
  const long Vh = orc_lat.Vh, F = Vh * 24;
  const REAL *ie = in, *io = in + 2 * F;
  REAL *oe = out, *oo = out + 2 * F;
  REAL *tmp = (REAL *)malloc(sizeof(REAL) * 4 * F);
  FN(orc_wil_dslash)(oo, gauge, ie, 1, dagger);
  FN(orc_wil_dslash)(oo + F, gauge, ie + F, 1, dagger);
  FN(orc_wil_dslash)(oe, gauge, io, 0, dagger);
  FN(orc_wil_dslash)(oe + F, gauge, io + F, 0, dagger);
  FN(orc_ndeg_twist)(tmp, tmp + F, ie, ie + F, dagger, kappa, mu, eps, Vh, 0);
  FN(orc_ndeg_twist)(tmp + 2 * F, tmp + 3 * F, io, io + F, dagger, kappa, mu, eps, Vh, 0);
  FN(xpay_)(tmp + 2 * F, (REAL)(-kappa), oo, 2 * F);
  FN(xpay_)(tmp, (REAL)(-kappa), oe, 2 * F);
  free(tmp);
}

/* ---------------------------------------------------------------------------------------------
 * Twisted-clover (and Wilson-clover) host path: /root/reference/tests/clover_reference.cpp
 *   apply_clover / cloverReference   :19-79   packed clover: per site and chirality 6 real diagonal
 *                                             entries, then the 15 complex strictly-lower-triangular
 *                                             entries column by column (QUDA_PACKED_CLOVER_ORDER)
 *   applyTwist                       :160-192 out = tmpH + i a gamma5 in
 *   twistCloverGamma5                :203-232 direct: C + i a g5;  inverse: cInv (C - i a0 g5), cInv = (C^2 + a0^2)^-1
 *   tmc_dslash :234-255, tmc_mat :257-282, tmc_matpc :284-341
 * Same operation order as the reference (fp64 agrees to the bit); inputs are never modified.
 * ------------------------------------------------------------------------------------------- */
void FN(orc_apply_clover)(REAL *out, const REAL *clover, const REAL *in, int parity)
{
  const long Vh = orc_lat.Vh;
  const int N = 6, chiralBlock = 36;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < Vh; i++) {
    const REAL *In = in + i * 24;
    REAL *Out = out + i * 24;
    for (int chi = 0; chi < 2; chi++) {
      const REAL *D = clover + ((parity * Vh + i) * 2 + chi) * chiralBlock;
      const REAL *L = D + N;
      for (int col = 0; col < N; col++) {
        const int Col = chi * N + col;
        REAL re = 0, im = 0;
        for (int row = 0; row < N; row++) {
          const int Row = chi * N + row;
          const REAL br = In[2 * Row], bi = In[2 * Row + 1];
          if (row == col) {
            re += D[row] * br; im += D[row] * bi;
          } else if (col < row) {
            const int k = N * (N - 1) / 2 - (N - col) * (N - col - 1) / 2 + row - col - 1;
            const REAL ar = L[2 * k], ai = -L[2 * k + 1];  /* conj(L[k]) */
            re += ar * br - ai * bi; im += ar * bi + ai * br;
          } else {
            const int k = N * (N - 1) / 2 - (N - row) * (N - row - 1) / 2 + col - row - 1;
            const REAL ar = L[2 * k], ai = L[2 * k + 1];
            re += ar * br - ai * bi; im += ar * bi + ai * br;
          }
        }
        Out[2 * Col] = re; Out[2 * Col + 1] = im;
      }
    }
  }
}

static void FN(apply_twist_)(REAL *out, const REAL *in, const REAL *tmpH, double a)
{
  const long Vh = orc_lat.Vh;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < Vh; i++)
    for (int s = 0; s < 4; s++) {
      const REAL a5 = (REAL)(((s / 2) ? -1.0 : +1.0) * a);
      for (int c = 0; c < 3; c++) {
        const long o = i * 24 + s * 6 + c * 2;
        const REAL xr = in[o], xi = in[o + 1];
        out[o] = tmpH[o] - a5 * xi;
        out[o + 1] = tmpH[o + 1] + a5 * xr;
      }
    }
}

void FN(orc_twist_clover_gamma5)(REAL *out, const REAL *in, const REAL *clover, const REAL *cinv, int dagger, double kappa, double mu,
                                 int flavor, int parity, int inverse)
{
  const long Vh = orc_lat.Vh;
  REAL *tmp1 = (REAL *)malloc(sizeof(REAL) * Vh * 24), *tmp2 = (REAL *)malloc(sizeof(REAL) * Vh * 24);
  if (!inverse) {
    double a = 2.0 * kappa * mu * flavor;
    if (dagger) a *= -1.0;
    FN(orc_apply_clover)(tmp1, clover, in, parity);
    FN(apply_twist_)(tmp2, in, tmp1, a);
    memcpy(out, tmp2, sizeof(REAL) * Vh * 24);
  } else {
    double a = -2.0 * kappa * mu * flavor;
    if (dagger) a *= -1.0;
    FN(orc_apply_clover)(tmp1, clover, in, parity);
    FN(apply_twist_)(tmp2, in, tmp1, a);
    FN(orc_apply_clover)(tmp1, cinv, tmp2, parity);
    memcpy(out, tmp1, sizeof(REAL) * Vh * 24);
  }
  free(tmp1); free(tmp2);
}

void FN(orc_tmc_dslash)(REAL *out, REAL *const *gauge, const REAL *in, const REAL *clover, const REAL *cinv, double kappa, double mu, int flavor,
                        int parity, int matpc, int dagger)
{
  const long Vh = orc_lat.Vh;
  REAL *tmp1 = (REAL *)malloc(sizeof(REAL) * Vh * 24), *tmp2 = (REAL *)malloc(sizeof(REAL) * Vh * 24);
  if (dagger) {
    FN(orc_twist_clover_gamma5)(tmp1, in, clover, cinv, dagger, kappa, mu, flavor, 1 - parity, 1);
    if (!FN(symmetric_)(matpc)) {
      FN(orc_wil_dslash)(tmp2, gauge, tmp1, parity, dagger);
      FN(orc_twist_clover_gamma5)(out, tmp2, clover, cinv, dagger, kappa, mu, flavor, parity, 1);
    } else {
      FN(orc_wil_dslash)(out, gauge, tmp1, parity, dagger);
    }
  } else {
    FN(orc_wil_dslash)(tmp1, gauge, in, parity, dagger);
    FN(orc_twist_clover_gamma5)(out, tmp1, clover, cinv, dagger, kappa, mu, flavor, parity, 1);
  }
  free(tmp1); free(tmp2);
}

void FN(orc_tmc_mat)(REAL *out, REAL *const *gauge, const REAL *clover, const REAL *in, double kappa, double mu, int flavor, int dagger)
{
  const long Vh = orc_lat.Vh, V = 2 * Vh;
  REAL *tmp = (REAL *)malloc(sizeof(REAL) * V * 24);
  FN(orc_wil_dslash)(out + Vh * 24, gauge, in, 1, dagger);
  FN(orc_twist_clover_gamma5)(tmp + Vh * 24, in + Vh * 24, clover, NULL, dagger, kappa, mu, flavor, 1, 0);
  FN(orc_wil_dslash)(out, gauge, in + Vh * 24, 0, dagger);
  FN(orc_twist_clover_gamma5)(tmp, in, clover, NULL, dagger, kappa, mu, flavor, 0, 0);
  FN(xpay_)(tmp, (REAL)-kappa, out, V * 24);
  free(tmp);
}

void FN(orc_tmc_matpc)(REAL *out, REAL *const *gauge, const REAL *in, const REAL *clover, const REAL *cinv, double kappa, double mu, int flavor,
                       int matpc, int dagger)
{
  const long Vh = orc_lat.Vh;
  const double kappa2 = -kappa * kappa;
  const int p = (matpc == ORC_MATPC_EVEN_EVEN || matpc == ORC_MATPC_EVEN_EVEN_ASYM) ? 0 : 1, q = 1 - p;
  REAL *tmp1 = (REAL *)malloc(sizeof(REAL) * Vh * 24), *tmp2 = (REAL *)malloc(sizeof(REAL) * Vh * 24);
  if (FN(symmetric_)(matpc)) {
    if (!dagger) {
      FN(orc_wil_dslash)(out, gauge, in, q, dagger);
      FN(orc_twist_clover_gamma5)(tmp1, out, clover, cinv, dagger, kappa, mu, flavor, q, 1);
      FN(orc_wil_dslash)(tmp2, gauge, tmp1, p, dagger);
      FN(orc_twist_clover_gamma5)(out, tmp2, clover, cinv, dagger, kappa, mu, flavor, p, 1);
    } else {
      FN(orc_twist_clover_gamma5)(out, in, clover, cinv, dagger, kappa, mu, flavor, p, 1);
      FN(orc_wil_dslash)(tmp1, gauge, out, q, dagger);
      FN(orc_twist_clover_gamma5)(tmp2, tmp1, clover, cinv, dagger, kappa, mu, flavor, q, 1);
      FN(orc_wil_dslash)(out, gauge, tmp2, p, dagger);
    }
    FN(xpay_)(in, (REAL)kappa2, out, Vh * 24);
  } else {
    FN(orc_wil_dslash)(tmp1, gauge, in, q, dagger);
    FN(orc_twist_clover_gamma5)(tmp2, tmp1, clover, cinv, dagger, kappa, mu, flavor, q, 1);
    FN(orc_wil_dslash)(out, gauge, tmp2, p, dagger);
    FN(orc_twist_clover_gamma5)(tmp2, in, clover, cinv, dagger, kappa, mu, flavor, p, 0);
    FN(xpay_)(tmp2, (REAL)kappa2, out, Vh * 24);
  }
  free(tmp1); free(tmp2);
}

#undef FN
#undef CAT
#undef CAT_
