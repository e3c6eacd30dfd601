"""GPU parity tests of the BLAS-1 / reduction kernels used by GCR / MR / BiCGStab, in the style of the reference's
tests/blas_test.cu (same operation on the host, relative error of norms <= 1e-11 double / 1e-5 single,
blas_test.cu:453, :952-962).  Oracle: numpy complex128 restatement of the functor bodies of
lib/blas_quda.cu:109-692 and lib/reduce_quda.cu:166-849."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

A, B = 0.37 - 1.21j, -0.58 + 0.44j


def run(L, name, prec, x, y, z, w):
    dt = np.float64 if prec == 8 else np.float32
    arrs = []
    for v in (x, y, z, w):
        f = np.empty(2 * v.size, dtype=dt)
        f[0::2] = v.real; f[1::2] = v.imag
        arrs.append(f)
    coef = (C.c_double * 4)(A.real, A.imag, B.real, B.imag)
    res = (C.c_double * 8)()
    n = L.blasQudaB200(name.encode(), x.size, prec, coef, *[a.ctypes.data_as(C.c_void_p) for a in arrs], res)
    back = [a[0::2].astype(np.float64) + 1j * a[1::2].astype(np.float64) for a in arrs]
    return back, list(res)[:n]


def host(name, x, y, z, w):
    a, b, ar, br = A, B, A.real, B.real
    r = None
    if name == "ax": x = ar * x
    elif name == "axpy": y = y + ar * x
    elif name == "xpy": y = y + x
    elif name == "xpay": y = x + ar * y
    elif name == "mxpy": y = y - x
    elif name == "axpby": y = ar * x + br * y
    elif name == "caxpy": y = y + a * x
    elif name == "caxpby": y = a * x + b * y
    elif name == "cxpaypbz": z = x + a * y + b * z
    elif name == "caxpbypz": z = z + a * x + b * y
    elif name == "caxpbypzYmbw": z = z + a * x + b * y; y = y - b * w
    elif name == "cabxpyAx": y = y + ar * b * x; x = ar * x
    elif name == "caxpyXmaz": y = y + a * x; x = x - a * z
    elif name == "norm2": r = [np.vdot(x, x).real]
    elif name == "reDotProduct": r = [np.vdot(x, y).real]
    elif name == "cDotProduct": d = np.vdot(x, y); r = [d.real, d.imag]
    elif name == "cDotProductNormA": d = np.vdot(x, y); r = [d.real, d.imag, np.vdot(x, x).real]
    elif name == "cDotProductNormB": d = np.vdot(x, y); r = [d.real, d.imag, np.vdot(y, y).real]
    elif name == "axpyNorm": y = y + ar * x; r = [np.vdot(y, y).real]
    elif name == "xmyNorm": y = x - y; r = [np.vdot(y, y).real]
    elif name == "caxpyNorm": y = y + a * x; r = [np.vdot(y, y).real]
    elif name == "cabxpyAxNorm": y = y + ar * b * x; x = ar * x; r = [np.vdot(y, y).real]
    elif name == "caxpyDotzy": y = y + a * x; d = np.vdot(z, y); r = [d.real, d.imag]
    elif name == "caxpyXmazNormX": y = y + a * x; x = x - a * z; r = [np.vdot(x, x).real]
    elif name == "xpaycDotzy": y = x + ar * y; d = np.vdot(z, y); r = [d.real, d.imag]
    elif name == "bicgstabUpdate": w = w + a * x + b * y; y = y - b * z; d = np.vdot(x, y); r = [d.real, d.imag, np.vdot(y, y).real]
    elif name == "block_cDotProduct":
        r = []
        for v in (x, y, z):
            d = np.vdot(v, w); r += [d.real, d.imag]
    elif name == "block_caxpy": w = w + a * x + b * y + np.conj(a) * z
    else: raise KeyError(name)
    return [x, y, z, w], r


OPS = ["ax", "axpy", "xpy", "xpay", "mxpy", "axpby", "caxpy", "caxpby", "cxpaypbz", "caxpbypz", "caxpbypzYmbw", "cabxpyAx",
       "caxpyXmaz", "norm2", "reDotProduct", "cDotProduct", "cDotProductNormA", "cDotProductNormB", "axpyNorm", "xmyNorm",
       "caxpyNorm", "cabxpyAxNorm", "caxpyDotzy", "caxpyXmazNormX", "xpaycDotzy", "bicgstabUpdate", "block_cDotProduct", "block_caxpy"]


@pytest.mark.parametrize("prec", [8, 4])
@pytest.mark.parametrize("n", [12, 12 * 1000, 12 * 70001])
def test_blas_against_host(quda, prec, n):
    L = quda.lib()
    rng = np.random.default_rng(n + prec)
    vecs = [rng.standard_normal(n) + 1j * rng.standard_normal(n) for _ in range(4)]
    if prec == 4:
        vecs = [v.astype(np.complex64).astype(np.complex128) for v in vecs]
    tol = 1e-11 if prec == 8 else 1e-5
    for name in OPS:
        got, res = run(L, name, prec, *vecs)
        want, wres = host(name, *vecs)
        for g, h in zip(got, want):
            assert abs(np.linalg.norm(g) - np.linalg.norm(h)) <= tol * np.linalg.norm(h), name
            assert np.linalg.norm(g - h) <= 10 * tol * np.linalg.norm(h), name
        if wres is not None:
            scale = np.linalg.norm(vecs[0]) * np.linalg.norm(vecs[1]) if n > 12 else 50.0
            assert len(res) == len(wres), name
            for a, b in zip(res, wres):
                assert abs(a - b) <= tol * max(abs(b), scale), (name, a, b)


def test_reductions_are_deterministic(quda):
    L = quda.lib()
    rng = np.random.default_rng(3)
    n = 12 * 50000
    vecs = [rng.standard_normal(n) + 1j * rng.standard_normal(n) for _ in range(4)]
    first = run(L, "cDotProductNormA", 4, *vecs)[1]
    for _ in range(5):
        assert run(L, "cDotProductNormA", 4, *vecs)[1] == first


@pytest.mark.parametrize("prec", [8, 4])
def test_blas_against_the_reference_host_blas(quda, prec):
    """the five operations the reference's own host BLAS provides (tests/blas_reference.cpp: ax, axpy, xpay, mxpy, norm_2, compiled
    unmodified into oracle/_ref/libtmref.so -- what its solvers' host verification uses), same inputs, element-wise"""
    import os
    from tests import oracle_util as ou
    path = os.path.join(ou.ORACLE_DIR, "_ref", "libtmref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libtmref.so is absent")
    R = C.CDLL(path)
    vp_, dbl, it = C.c_void_p, C.c_double, C.c_int
    R.norm_2.restype = dbl
    R.norm_2.argtypes = [vp_, it, it]
    R.ax.argtypes = [dbl, vp_, it, it]
    R.axpy.argtypes = [dbl, vp_, vp_, it, it]
    R.xpay.argtypes = [vp_, dbl, vp_, it, it]
    R.mxpy.argtypes = [vp_, vp_, it, it]
    L = quda.lib()
    n = 12 * 4099
    dt = np.float64 if prec == 8 else np.float32
    rng = np.random.default_rng(prec)
    vecs = [(rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex128 if prec == 8 else np.complex64).astype(np.complex128) for _ in range(4)]

    def reals(v):
        f = np.empty(2 * v.size, dtype=dt)
        f[0::2] = v.real; f[1::2] = v.imag
        return f

    def ptr(a):
        return a.ctypes.data_as(C.c_void_p)

    tol = 1e-14 if prec == 8 else 1e-6
    a = A.real
    for name in ("ax", "axpy", "xpay", "mxpy", "norm2"):
        got, res = run(L, name, prec, *vecs)
        x, y = reals(vecs[0]), reals(vecs[1])
        if name == "ax": R.ax(a, ptr(x), x.size, prec)
        elif name == "axpy": R.axpy(a, ptr(x), ptr(y), x.size, prec)
        elif name == "xpay": R.xpay(ptr(x), a, ptr(y), x.size, prec)
        elif name == "mxpy": R.mxpy(ptr(x), ptr(y), x.size, prec)
        else:
            ref = R.norm_2(ptr(x), x.size, prec)
            assert abs(res[0] - ref) <= (1e-13 if prec == 8 else 1e-6) * ref, (res[0], ref)
            continue
        for g, h in zip(got[:2], (x, y)):
            hc = h[0::2].astype(np.float64) + 1j * h[1::2].astype(np.float64)
            assert np.abs(g - hc).max() <= tol * np.abs(hc).max(), name


def run_half(L, name, x, y, z, w):
    """prec = 2: fp32 host arrays in the internal plane order of a parity field ([plane of 2 complex][site]); the library converts to
    int16 + norm and back"""
    arrs = []
    for v in (x, y, z, w):
        f = np.empty(2 * v.size, dtype=np.float32)
        f[0::2] = v.real; f[1::2] = v.imag
        arrs.append(f)
    coef = (C.c_double * 4)(A.real, A.imag, B.real, B.imag)
    res = (C.c_double * 8)()
    n = L.blasQudaB200(name.encode(), x.size, 2, coef, *[a.ctypes.data_as(C.c_void_p) for a in arrs], res)
    back = [a[0::2].astype(np.float64) + 1j * a[1::2].astype(np.float64) for a in arrs]
    return back, list(res)[:n]


@pytest.mark.parametrize("n", [12 * 64, 12 * 40000])
def test_half_precision_blas(quda, n):
    """int16 + norm solver vectors (the reference's half precision: short4 + norm, fp32 arithmetic, lib/blas_core.h:12-52,
    lib/io_spinor.h:282-320): every BLAS / reduction operation against the host on the same fp32 inputs.  Each site is quantised to
    16 bits relative to its largest component on the way in and on the way out: tolerance 1e-3 of the vector norm (north_star's half bar);
    measured 2e-5 .. 6e-5."""
    L = quda.lib()
    rng = np.random.default_rng(n)
    vecs = [(rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64).astype(np.complex128) for _ in range(4)]
    worst = 0.0
    for name in OPS:
        got, res = run_half(L, name, *vecs)
        want, wres = host(name, *vecs)
        for g, h in zip(got, want):
            err = np.linalg.norm(g - h) / np.linalg.norm(h)
            worst = max(worst, err)
            assert err <= 1e-3, (name, err)
        if wres is not None:
            scale = np.linalg.norm(vecs[0]) * np.linalg.norm(vecs[1])
            assert len(res) == len(wres), name
            for a, b in zip(res, wres):
                assert abs(a - b) <= 1e-3 * max(abs(b), scale), (name, a, b)
    print(f"half-precision BLAS, n = {n}: worst relative L2 error of an output vector {worst:.2e}")
