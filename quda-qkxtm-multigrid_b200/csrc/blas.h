// BLAS-1 and reductions on resident fields (replaces /root/reference/lib/blas_quda.cu,
// lib/reduce_quda.cu, include/blas_quda.h:33-144 for the GCR / MR / BiCGStab / multigrid path).
//
// Fields are flat arrays of complex numbers (every 16-byte plane holds whole complex values), so
// all kernels are layout-agnostic grid-stride loops with 128-bit accesses.  Reductions accumulate
// in double (QudaSumFloat of the reference, lib/reduce_quda.cu:30-33), finish on the device
// (one partial per CTA + a last-block pass) and are summed across ranks with NCCL.
// Same names / argument order / return conventions as the reference's quda::blas namespace.
#pragma once
#include <complex>
#include <vector>
#include "field.h"

namespace qb {
namespace blas {

typedef std::complex<double> Complex;
struct double3_ { double x, y, z; };

extern unsigned long long flops, bytes;

void init();
void end();

void zero(SpinorField &a);
void copy(SpinorField &dst, const SpinorField &src);  // converts precision if needed

void ax(double a, SpinorField &x);                                         // x *= a
void axpy(double a, const SpinorField &x, SpinorField &y);                 // y += a x
void xpy(const SpinorField &x, SpinorField &y);                            // y += x
void xpay(const SpinorField &x, double a, SpinorField &y);                 // y = x + a y
void mxpy(const SpinorField &x, SpinorField &y);                           // y -= x
void axpby(double a, const SpinorField &x, double b, SpinorField &y);      // y = a x + b y
void caxpy(Complex a, const SpinorField &x, SpinorField &y);               // y += a x
void caxpby(Complex a, const SpinorField &x, Complex b, SpinorField &y);   // y = a x + b y
void cxpaypbz(const SpinorField &x, Complex a, const SpinorField &y, Complex b, SpinorField &z);  // z = x + a y + b z
void caxpbypz(Complex a, const SpinorField &x, Complex b, const SpinorField &y, SpinorField &z);  // z += a x + b y
void caxpbypzYmbw(Complex a, const SpinorField &x, Complex b, SpinorField &y, SpinorField &z, const SpinorField &w);  // z += a x + b y; y -= b w
void cabxpyAx(double a, Complex b, SpinorField &x, SpinorField &y);        // y += a b x; x *= a
void caxpyXmaz(Complex a, SpinorField &x, SpinorField &y, const SpinorField &z);  // y += a x; x -= a z
void mrFirstStep(Complex a, const SpinorField &b, const SpinorField &Ab, SpinorField &x, SpinorField &r, bool accumulate);  // x (+)= a b; r = b - a Ab
void cax(Complex a, const SpinorField &x, SpinorField &y);                  // y = a x

double norm2(const SpinorField &x);
double reDotProduct(const SpinorField &x, const SpinorField &y);
Complex cDotProduct(const SpinorField &x, const SpinorField &y);           // sum conj(x) y
double3_ cDotProductNormA(const SpinorField &x, const SpinorField &y);     // (Re, Im, |x|^2)
double3_ cDotProductNormB(const SpinorField &x, const SpinorField &y);     // (Re, Im, |y|^2)
double axpyNorm(double a, const SpinorField &x, SpinorField &y);           // y += a x; |y|^2
double xmyNorm(const SpinorField &x, SpinorField &y);                      // y = x - y; |y|^2
double caxpyNorm(Complex a, const SpinorField &x, SpinorField &y);         // y += a x; |y|^2
// x += a p + w r; r -= w t; (Re <r0, r>, Im <r0, r>, |r|^2) of the new r
double3_ bicgstabUpdate(Complex a, const SpinorField &p, Complex w, SpinorField &r, const SpinorField &t, SpinorField &x, const SpinorField &r0);
double cabxpyAxNorm(double a, Complex b, SpinorField &x, SpinorField &y);  // y += a b x; x *= a; |y|^2 (what lib/reduce_quda.cu:490-497 computes; its comment says x)
Complex caxpyDotzy(Complex a, const SpinorField &x, SpinorField &y, const SpinorField &z);  // y += a x; (z, y)
double caxpyXmazNormX(Complex a, SpinorField &x, SpinorField &y, const SpinorField &z);     // y += a x; x -= a z; |x|^2
Complex xpaycDotzy(const SpinorField &x, double a, SpinorField &y, const SpinorField &z);   // y = x + a y; (z, y)

// block variants (one pass over memory): result[i] = (x_i, y);  y += sum_i a_i x_i
void cDotProduct(Complex *result, const std::vector<SpinorField *> &x, const SpinorField &y);
void caxpy(const Complex *a, const std::vector<SpinorField *> &x, SpinorField &y);

// when false, reductions are not summed over ranks (Schwarz-style local smoothers,
// lib/inv_mr_quda.cpp:39,128-133)
void set_global_reduction(bool on);

}  // namespace blas
}  // namespace qb
