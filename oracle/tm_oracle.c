/* TEST INFRASTRUCTURE ONLY (oracle) -- plain-C CPU restatement of the reference's host
 * verification path for the twisted-mass Wilson Dslash / Mat / MatPC and of its deterministic
 * input generators.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this; the product library never does.
 *
 * Parity pin: tests/test_oracle.py checks every function here against (a) the known-answer
 * fingerprints of SURVEY.md Appendix B, (b) oracle/_ref/libtmref.so = the reference's own
 * unmodified sources compiled in place (bit-for-bit in fp64) and (c) the committed golden
 * fixtures under tests/golden/ produced by that library.
 *
 * Reference files followed:
 *   site indexing      /root/reference/tests/test_util.cpp:419-472
 *   gauge generator    /root/reference/tests/test_util.cpp:865-925 (random SU(3)), :682-704 (scaling, BC)
 *   spinor generator   /root/reference/lib/comm_common.cpp:73-87 (48-bit LCG), lib/color_spinor_util.cu:12-24
 *   operators          see tm_oracle_impl.h
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int X[4];
  long V, Vh;
} orc_lattice_t;

static orc_lattice_t orc_lat;

enum { ORC_MATPC_EVEN_EVEN = 0, ORC_MATPC_ODD_ODD = 1, ORC_MATPC_EVEN_EVEN_ASYM = 2, ORC_MATPC_ODD_ODD_ASYM = 3 };

void orc_set_dims(const int *X)
{
  orc_lat.V = 1;
  for (int d = 0; d < 4; d++) { orc_lat.X[d] = X[d]; orc_lat.V *= X[d]; }
  orc_lat.Vh = orc_lat.V / 2;
}

/* checkerboard index -> coordinates: idx = ((t Z + z) Y + y) X + x, cb = idx / 2 */
static inline void orc_coords(int *x, long cb, int parity)
{
  const int *X = orc_lat.X;
  long za = cb / (X[0] / 2);
  long zb = za / X[1];
  x[1] = (int)(za - zb * X[1]);
  x[3] = (int)(zb / X[2]);
  x[2] = (int)(zb - (long)x[3] * X[2]);
  x[0] = (int)(2 * cb + ((x[1] + x[2] + x[3] + parity) & 1) - za * X[0]);
}

static inline long orc_cb_index(const int *x)
{
  const int *X = orc_lat.X;
  return ((((long)x[3] * X[2] + x[2]) * X[1] + x[1]) * X[0] + x[0]) >> 1;
}

/* exported for index-parity tests */
void orc_site_coords(int *x, long cb, int parity) { orc_coords(x, cb, parity); }
long orc_site_index(const int *x) { return orc_cb_index(x); }

/* DeGrand-Rossi gamma matrices: one non-zero per row; gamma_mu[s][orc_gamma_col[mu][s]] = val */
static const int orc_gamma_col[4][4] = {{3, 2, 1, 0}, {3, 2, 1, 0}, {2, 3, 0, 1}, {2, 3, 0, 1}};
static const double orc_gamma_val[4][4][2] = {
  {{0, 1}, {0, 1}, {0, -1}, {0, -1}},  /* gamma_x */
  {{-1, 0}, {1, 0}, {1, 0}, {-1, 0}},  /* gamma_y */
  {{0, 1}, {0, -1}, {0, -1}, {0, 1}},  /* gamma_z */
  {{1, 0}, {1, 0}, {1, 0}, {1, 0}}};   /* gamma_t */

/* 48-bit LCG of the reference (clone of drand48), state passed in/out */
void orc_drand_fill(double *out, long n, uint64_t *state)
{
  const uint64_t m = 25214903917ULL, a = 11ULL, mask = 281474976710655ULL;
  uint64_t s = *state;
  for (long i = 0; i < n; i++) {
    s = (m * s + a) & mask;
    out[i] = 0.35527136788005009e-14 * (double)s;
  }
  *state = s;
}

void orc_drand_fill_f(float *out, long n, uint64_t *state)
{
  const uint64_t m = 25214903917ULL, a = 11ULL, mask = 281474976710655ULL;
  uint64_t s = *state;
  for (long i = 0; i < n; i++) {
    s = (m * s + a) & mask;
    out[i] = (float)(0.35527136788005009e-14 * (double)s);
  }
  *state = s;
}

/* ---- gauge generator ------------------------------------------------------------------ */
static void row_normalize(double *r)
{
  double s = 0.0;
  for (int i = 0; i < 3; i++) s += r[2 * i] * r[2 * i] + r[2 * i + 1] * r[2 * i + 1];
  const double n = sqrt(s);
  for (int i = 0; i < 3; i++) {
    /* complex / real as std::complex<double>::operator/=(double) does it */
    r[2 * i] /= n;
    r[2 * i + 1] /= n;
  }
}

static void row_orthogonalize(const double *a, double *b)
{
  double dr = 0.0, di = 0.0;
  for (int i = 0; i < 3; i++) {
    /* conj(a)*b */
    dr += a[2 * i] * b[2 * i] + a[2 * i + 1] * b[2 * i + 1];
    di += a[2 * i] * b[2 * i + 1] - a[2 * i + 1] * b[2 * i];
  }
  for (int i = 0; i < 3; i++) {
    const double pr = dr * a[2 * i] - di * a[2 * i + 1];
    const double pi = dr * a[2 * i + 1] + di * a[2 * i];
    b[2 * i] -= pr;
    b[2 * i + 1] -= pi;
  }
}

static inline void acc_conj_prod(double *a, const double *b, const double *c, int sign)
{
  a[0] += sign * (b[0] * c[0] - b[1] * c[1]);
  a[1] -= sign * (b[0] * c[1] + b[1] * c[0]);
}

static void complete_su3(double *m)
{
  double *w = m, *u = m + 6, *v = m + 12;
  row_normalize(u);
  row_orthogonalize(u, v);
  row_normalize(v);
  for (int n = 0; n < 6; n++) w[n] = 0.0;
  acc_conj_prod(w + 0, u + 2, v + 4, +1);
  acc_conj_prod(w + 0, u + 4, v + 2, -1);
  acc_conj_prod(w + 2, u + 4, v + 0, +1);
  acc_conj_prod(w + 2, u + 0, v + 4, -1);
  acc_conj_prod(w + 4, u + 0, v + 2, +1);
  acc_conj_prod(w + 4, u + 2, v + 0, -1);
}

/* type 0: unit links, 1: random SU(3) from glibc rand() (seeded here), QDP order
 * gauge[mu][(parity*Vh + cb)*18 + (row*3+col)*2 + reim]; then anisotropy and antiperiodic-T sign
 * on the last time slice.  fp64 only (the fp32 variant of the reference draws the same rand()
 * sequence but rounds per element; callers that need fp32 links convert the fp64 ones). */
void orc_construct_gauge(double *const *gauge, int type, int antiperiodic_t, double anisotropy, unsigned seed)
{
  const long Vh = orc_lat.Vh;
  const int *X = orc_lat.X;
  if (type == 0) {
    for (int mu = 0; mu < 4; mu++) {
      memset(gauge[mu], 0, sizeof(double) * 2 * Vh * 18);
      for (long i = 0; i < 2 * Vh; i++)
        for (int r = 0; r < 3; r++) gauge[mu][i * 18 + r * 8] = 1.0;
    }
  } else {
    srand(seed);
    for (int mu = 0; mu < 4; mu++)
      for (long i = 0; i < Vh; i++) {
        double *e = gauge[mu] + i * 18, *o = gauge[mu] + (Vh + i) * 18;
        for (int m = 1; m < 3; m++)
          for (int n = 0; n < 3; n++) {
            e[m * 6 + n * 2 + 0] = rand() / (double)RAND_MAX;
            e[m * 6 + n * 2 + 1] = rand() / (double)RAND_MAX;
            o[m * 6 + n * 2 + 0] = rand() / (double)RAND_MAX;
            o[m * 6 + n * 2 + 1] = rand() / (double)RAND_MAX;
          }
        complete_su3(e);
        complete_su3(o);
      }
  }
  for (int mu = 0; mu < 3; mu++)
    for (long i = 0; i < 2 * Vh * 18; i++) gauge[mu][i] /= anisotropy;
  if (antiperiodic_t) {
    const long first = (long)(X[0] / 2) * X[1] * X[2] * (X[3] - 1);
    for (long j = first; j < Vh; j++)
      for (int k = 0; k < 18; k++) {
        gauge[3][j * 18 + k] *= -1.0;
        gauge[3][(Vh + j) * 18 + k] *= -1.0;
      }
  }
}

/* weak-field SU(3): U = exp(i eps H) via project(1 + i eps H), H random hermitian from the LCG.
 * Not in the reference; used for MG solve tests where fully random links have no near-null space
 * (SURVEY.md section 8d).  Deterministic in (seed, eps). */
void orc_construct_weak_gauge(double *const *gauge, double eps, int antiperiodic_t, uint64_t seed)
{
  const long Vh = orc_lat.Vh;
  const int *X = orc_lat.X;
  uint64_t st = seed;
  for (int mu = 0; mu < 4; mu++)
    for (long i = 0; i < 2 * Vh; i++) {
      double r[12], *m = gauge[mu] + i * 18;
      orc_drand_fill(r, 12, &st);
      for (int k = 0; k < 12; k++) r[k] = eps * (2.0 * r[k] - 1.0);
      /* rows 1,2 of (1 + i eps H) with H hermitian built from r */
      double H[3][3][2];
      H[0][0][0] = r[0]; H[1][1][0] = r[1]; H[2][2][0] = -r[0] - r[1];
      H[0][0][1] = H[1][1][1] = H[2][2][1] = 0;
      H[0][1][0] = r[2]; H[0][1][1] = r[3]; H[1][0][0] = r[2]; H[1][0][1] = -r[3];
      H[0][2][0] = r[4]; H[0][2][1] = r[5]; H[2][0][0] = r[4]; H[2][0][1] = -r[5];
      H[1][2][0] = r[6]; H[1][2][1] = r[7]; H[2][1][0] = r[6]; H[2][1][1] = -r[7];
      for (int a = 1; a < 3; a++)
        for (int b = 0; b < 3; b++) {
          m[a * 6 + b * 2 + 0] = (a == b ? 1.0 : 0.0) - H[a][b][1];
          m[a * 6 + b * 2 + 1] = H[a][b][0];
        }
      complete_su3(m);
    }
  if (antiperiodic_t) {
    const long first = (long)(X[0] / 2) * X[1] * X[2] * (X[3] - 1);
    for (long j = first; j < Vh; j++)
      for (int k = 0; k < 18; k++) {
        gauge[3][j * 18 + k] *= -1.0;
        gauge[3][(Vh + j) * 18 + k] *= -1.0;
      }
  }
}

/* random clover term of the reference's tests (tests/test_util.cpp:1100-1114 constructCloverField): uniform entries in
 * (-norm, norm) from glibc rand(), `diag` added to the 2 x 6 diagonal entries; 72 reals per site, even sites first */
void orc_construct_clover(double *res, double norm, double diag, unsigned seed)
{
  const double c = 2.0 * norm / RAND_MAX;
  srand(seed);
  for (long i = 0; i < orc_lat.V; i++) {
    for (int j = 0; j < 72; j++) res[i * 72 + j] = c * rand() - norm;
    for (int j = 0; j < 6; j++) { res[i * 72 + j] += diag; res[i * 72 + j + 36] += diag; }
  }
}

#define REAL double
#define SUF _d
#include "tm_oracle_impl.h"
#undef REAL
#undef SUF

#define REAL float
#define SUF _f
#include "tm_oracle_impl.h"
#undef REAL
#undef SUF
