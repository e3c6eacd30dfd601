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

#undef FN
#undef CAT
#undef CAT_
