// TEST INFRASTRUCTURE ONLY (oracle).  Never linked into the product library.
//
// Thin C-ABI wrapper around the *reference's own, unmodified* CPU verification
// objects (tests/wilson_dslash_reference.cpp, tests/blas_reference.cpp,
// tests/test_util.cpp, tests/misc.cpp) which oracle/Makefile compiles from where
// they lie under /root/reference into oracle/_ref/libtmref.so.
//
// This file contains (a) the handful of symbols those objects expect from libquda
// (comms + logging hooks) as single-process stubs and (b) extern "C" entry points
// so tests / bench.py can drive the reference through ctypes.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <quda.h>
#include <test_util.h>
#include <wilson_dslash_reference.h>
#include <tune_key.h>
#include <comm_quda.h>

// ---- stubs for the libquda symbols the reference test objects pull in -------------
void comm_allreduce(double *) {}
void comm_allreduce_max(double *) {}
void comm_allreduce_int(int *) {}
int comm_dim_partitioned(int) { return 0; }
void comm_abort(int status) { exit(status); }
int comm_rank(void) { return 0; }
int comm_size(void) { return 1; }
int comm_dim(int) { return 1; }
int comm_coord(int) { return 0; }
void initCommsGridQuda(int, const int *, QudaCommsMap, void *) {}
static char prefix_[4] = "";
char *getOutputPrefix() { return prefix_; }
FILE *getOutputFile() { return stdout; }
static quda::TuneKey last_key_;
quda::TuneKey getLastTuneKey() { return last_key_; }


static QudaGaugeParam gparam_;

extern "C" {

// lattice setup: what tests/dslash_test.cpp:83-130 does on the host side
void tmref_setup(const int *X, int antiperiodic_t, double anisotropy)
{
  int dims[4] = {X[0], X[1], X[2], X[3]};
  setDims(dims);
  setSpinorSiteSize(24);
  memset(&gparam_, 0, sizeof(gparam_));
  for (int d = 0; d < 4; d++) gparam_.X[d] = X[d];
  gparam_.anisotropy = anisotropy;
  gparam_.type = QUDA_WILSON_LINKS;
  gparam_.gauge_order = QUDA_QDP_GAUGE_ORDER;
  gparam_.t_boundary = antiperiodic_t ? QUDA_ANTI_PERIODIC_T : QUDA_PERIODIC_T;
  gparam_.cpu_prec = QUDA_DOUBLE_PRECISION;
  gparam_.gauge_fix = QUDA_GAUGE_FIXED_NO;
}

// reference gauge generator (tests/test_util.cpp:1018): type 0 unit, 1 random SU(3)
void tmref_construct_gauge(void **gauge, int type, int prec_bytes, unsigned seed)
{
  srand(seed);
  gparam_.cpu_prec = (QudaPrecision)prec_bytes;
  construct_gauge_field(gauge, type, (QudaPrecision)prec_bytes, &gparam_);
}

void tmref_wil_dslash(void *out, void **gauge, void *in, int parity, int dagger, int prec_bytes)
{ wil_dslash(out, gauge, in, parity, dagger, (QudaPrecision)prec_bytes, gparam_); }

void tmref_tm_dslash(void *out, void **gauge, void *in, double kappa, double mu, int flavor,
                     int parity, int matpc, int dagger, int prec_bytes)
{ tm_dslash(out, gauge, in, kappa, mu, (QudaTwistFlavorType)flavor, parity, (QudaMatPCType)matpc,
            dagger, (QudaPrecision)prec_bytes, gparam_); }

void tmref_tm_matpc(void *out, void **gauge, void *in, double kappa, double mu, int flavor,
                    int matpc, int dagger, int prec_bytes)
{ tm_matpc(out, gauge, in, kappa, mu, (QudaTwistFlavorType)flavor, (QudaMatPCType)matpc, dagger,
           (QudaPrecision)prec_bytes, gparam_); }

void tmref_tm_mat(void *out, void **gauge, void *in, double kappa, double mu, int flavor,
                  int dagger, int prec_bytes)
{ tm_mat(out, gauge, in, kappa, mu, (QudaTwistFlavorType)flavor, dagger, (QudaPrecision)prec_bytes, gparam_); }

// non-degenerate doublet (wilson_dslash_reference.cpp:461-587).  Doublet parity field = [flavour 1 | flavour 2]; the reference
// twists its inputs in place for some variants, so callers hand in scratch copies.
void tmref_tm_ndeg_dslash(void *out, void **gauge, void *in, double kappa, double mu, double eps, int parity, int matpc, int dagger, int prec_bytes)
{
  const size_t F = (size_t)Vh * 24 * prec_bytes;
  tm_ndeg_dslash(out, (char *)out + F, gauge, in, (char *)in + F, kappa, mu, eps, parity, dagger, (QudaMatPCType)matpc, (QudaPrecision)prec_bytes, gparam_);
}
void tmref_tm_ndeg_matpc(void *out, void **gauge, void *in, double kappa, double mu, double eps, int matpc, int dagger, int prec_bytes)
{
  const size_t F = (size_t)Vh * 24 * prec_bytes;
  tm_ndeg_matpc(out, (char *)out + F, gauge, in, (char *)in + F, kappa, mu, eps, (QudaMatPCType)matpc, dagger, (QudaPrecision)prec_bytes, gparam_);
}
void tmref_tm_ndeg_mat(void *out, void **gauge, void *in, double kappa, double mu, double eps, int dagger, int prec_bytes)
{
  const size_t D = (size_t)Vh * 48 * prec_bytes;   // one parity of a doublet field
  tm_ndeg_mat(out, (char *)out + D, gauge, in, (char *)in + D, kappa, mu, eps, dagger, (QudaPrecision)prec_bytes, gparam_);
}

void tmref_wil_mat(void *out, void **gauge, void *in, double kappa, int dagger, int prec_bytes)
{ wil_mat(out, gauge, in, kappa, dagger, (QudaPrecision)prec_bytes, gparam_); }

void tmref_wil_matpc(void *out, void **gauge, void *in, double kappa, int matpc, int dagger, int prec_bytes)
{ wil_matpc(out, gauge, in, kappa, (QudaMatPCType)matpc, dagger, (QudaPrecision)prec_bytes, gparam_); }

// twisted-clover host path of the reference (tests/clover_reference.cpp) and its clover generator (tests/test_util.cpp:1116)
void tmref_construct_clover(void *clover, double norm, double diag, int prec_bytes, unsigned seed)
{ srand(seed); construct_clover_field(clover, norm, diag, (QudaPrecision)prec_bytes); }

void tmref_apply_clover(void *out, void *clover, void *in, int parity, int prec_bytes)
{ apply_clover(out, clover, in, parity, (QudaPrecision)prec_bytes); }

void tmref_tmc_dslash(void *out, void **gauge, void *in, void *clover, void *cinv, double kappa, double mu, int flavor,
                      int parity, int matpc, int dagger, int prec_bytes)
{ tmc_dslash(out, gauge, in, clover, cinv, kappa, mu, (QudaTwistFlavorType)flavor, parity, (QudaMatPCType)matpc, dagger,
             (QudaPrecision)prec_bytes, gparam_); }

void tmref_tmc_mat(void *out, void **gauge, void *clover, void *in, double kappa, double mu, int flavor, int dagger, int prec_bytes)
{ tmc_mat(out, gauge, clover, in, kappa, mu, (QudaTwistFlavorType)flavor, dagger, (QudaPrecision)prec_bytes, gparam_); }

void tmref_tmc_matpc(void *out, void **gauge, void *in, void *clover, void *cinv, double kappa, double mu, int flavor,
                     int matpc, int dagger, int prec_bytes)
{ tmc_matpc(out, gauge, in, clover, cinv, kappa, mu, (QudaTwistFlavorType)flavor, (QudaMatPCType)matpc, dagger,
            (QudaPrecision)prec_bytes, gparam_); }

int tmref_volume(void) { return V; }

} // extern "C"
