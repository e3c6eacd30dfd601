"""Test-side helpers: load the CPU oracles (oracle/liboracle.so, oracle/_ref/libtmref.so) and wrap
them with numpy signatures.  TEST INFRASTRUCTURE ONLY - the product never imports this."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _ptrs(g):
    return (C.c_void_p * 4)(*[a.ctypes.data for a in g])


class Oracle:
    """numpy face of oracle/tm_oracle.c (our C restatement of the reference's host verify path)."""

    def __init__(self, lib):
        self.L = lib
        L = lib
        L.orc_set_dims.argtypes = [C.POINTER(C.c_int)]
        L.orc_drand_fill.argtypes = [C.c_void_p, C.c_long, C.POINTER(C.c_uint64)]
        L.orc_construct_gauge.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_uint]
        L.orc_construct_weak_gauge.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_uint64]
        L.orc_site_coords.argtypes = [C.POINTER(C.c_int), C.c_long, C.c_int]
        L.orc_site_index.argtypes = [C.POINTER(C.c_int)]
        L.orc_site_index.restype = C.c_long
        for s in ("_d", "_f"):
            getattr(L, "orc_wil_dslash" + s).argtypes = [C.c_void_p] * 3 + [C.c_int] * 2
            getattr(L, "orc_twist_gamma5" + s).argtypes = [C.c_void_p] * 2 + [C.c_int, C.c_double, C.c_double, C.c_int, C.c_long, C.c_int]
            getattr(L, "orc_tm_dslash" + s).argtypes = [C.c_void_p] * 3 + [C.c_double] * 2 + [C.c_int] * 4
            getattr(L, "orc_tm_matpc" + s).argtypes = [C.c_void_p] * 3 + [C.c_double] * 2 + [C.c_int] * 3
            getattr(L, "orc_tm_mat" + s).argtypes = [C.c_void_p] * 3 + [C.c_double] * 2 + [C.c_int] * 2
            getattr(L, "orc_tm_ndeg_dslash" + s).argtypes = [C.c_void_p] * 3 + [C.c_double] * 3 + [C.c_int] * 3
            getattr(L, "orc_tm_ndeg_matpc" + s).argtypes = [C.c_void_p] * 3 + [C.c_double] * 3 + [C.c_int] * 2
            getattr(L, "orc_tm_ndeg_mat" + s).argtypes = [C.c_void_p] * 3 + [C.c_double] * 3 + [C.c_int]
            getattr(L, "orc_wil_mat" + s).argtypes = [C.c_void_p] * 3 + [C.c_double, C.c_int]
            getattr(L, "orc_wil_matpc" + s).argtypes = [C.c_void_p] * 3 + [C.c_double, C.c_int, C.c_int]
        L.orc_construct_clover.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_uint]
        for s in ("_d", "_f"):
            getattr(L, "orc_apply_clover" + s).argtypes = [C.c_void_p] * 3 + [C.c_int]
            getattr(L, "orc_twist_clover_gamma5" + s).argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int]
            getattr(L, "orc_tmc_dslash" + s).argtypes = [C.c_void_p] * 5 + [C.c_double] * 2 + [C.c_int] * 4
            getattr(L, "orc_tmc_mat" + s).argtypes = [C.c_void_p] * 4 + [C.c_double] * 2 + [C.c_int] * 2
            getattr(L, "orc_tmc_matpc" + s).argtypes = [C.c_void_p] * 5 + [C.c_double] * 2 + [C.c_int] * 3
        self.dims = None

    def set_dims(self, X):
        self.dims = tuple(int(x) for x in X)
        self.V = int(np.prod(self.dims))
        self.Vh = self.V // 2
        self.L.orc_set_dims((C.c_int * 4)(*self.dims))

    def drand(self, n, seed=137):
        out = np.empty(n, dtype=np.float64)
        st = C.c_uint64(seed)
        self.L.orc_drand_fill(_ptr(out), n, C.byref(st))
        return out

    def gauge(self, kind=1, antiperiodic=True, anisotropy=1.0, seed=137):
        g = [np.zeros(self.V * 18, dtype=np.float64) for _ in range(4)]
        self.L.orc_construct_gauge(_ptrs(g), kind, int(antiperiodic), anisotropy, seed)
        return g

    def weak_gauge(self, eps=0.2, antiperiodic=False, seed=4711):
        g = [np.zeros(self.V * 18, dtype=np.float64) for _ in range(4)]
        self.L.orc_construct_weak_gauge(_ptrs(g), eps, int(antiperiodic), seed)
        return g

    @staticmethod
    def _suffix(a):
        return "_d" if a.dtype == np.float64 else "_f"

    def _cast_gauge(self, g, dtype):
        return g if g[0].dtype == dtype else [a.astype(dtype) for a in g]

    def wil_dslash(self, g, inp, parity, dagger):
        g = self._cast_gauge(g, inp.dtype)
        out = np.zeros(self.Vh * 24, dtype=inp.dtype)
        getattr(self.L, "orc_wil_dslash" + self._suffix(inp))(_ptr(out), _ptrs(g), _ptr(inp), parity, dagger)
        return out

    def tm_dslash(self, g, inp, kappa, mu, flavor, parity, matpc, dagger):
        g = self._cast_gauge(g, inp.dtype)
        out = np.zeros(self.Vh * 24, dtype=inp.dtype)
        getattr(self.L, "orc_tm_dslash" + self._suffix(inp))(_ptr(out), _ptrs(g), _ptr(inp), kappa, mu, flavor, parity, matpc, dagger)
        return out

    def tm_matpc(self, g, inp, kappa, mu, flavor, matpc, dagger):
        g = self._cast_gauge(g, inp.dtype)
        out = np.zeros(self.Vh * 24, dtype=inp.dtype)
        getattr(self.L, "orc_tm_matpc" + self._suffix(inp))(_ptr(out), _ptrs(g), _ptr(inp), kappa, mu, flavor, matpc, dagger)
        return out

    def tm_mat(self, g, inp, kappa, mu, flavor, dagger):
        g = self._cast_gauge(g, inp.dtype)
        out = np.zeros(self.V * 24, dtype=inp.dtype)
        getattr(self.L, "orc_tm_mat" + self._suffix(inp))(_ptr(out), _ptrs(g), _ptr(inp), kappa, mu, flavor, dagger)
        return out

    # non-degenerate doublet: parity field = [flavour 1 | flavour 2] (2 * Vh * 24 reals), full field = [even doublet | odd doublet]
    def tm_ndeg_dslash(self, g, inp, kappa, mu, eps, parity, matpc, dagger):
        g = self._cast_gauge(g, inp.dtype)
        out = np.zeros(2 * self.Vh * 24, dtype=inp.dtype)
        getattr(self.L, "orc_tm_ndeg_dslash" + self._suffix(inp))(_ptr(out), _ptrs(g), _ptr(inp), kappa, mu, eps, parity, matpc, dagger)
        return out

    def tm_ndeg_matpc(self, g, inp, kappa, mu, eps, matpc, dagger):
        g = self._cast_gauge(g, inp.dtype)
        out = np.zeros(2 * self.Vh * 24, dtype=inp.dtype)
        getattr(self.L, "orc_tm_ndeg_matpc" + self._suffix(inp))(_ptr(out), _ptrs(g), _ptr(inp), kappa, mu, eps, matpc, dagger)
        return out

    def tm_ndeg_mat(self, g, inp, kappa, mu, eps, dagger):
        g = self._cast_gauge(g, inp.dtype)
        out = np.zeros(2 * self.V * 24, dtype=inp.dtype)
        getattr(self.L, "orc_tm_ndeg_mat" + self._suffix(inp))(_ptr(out), _ptrs(g), _ptr(inp), kappa, mu, eps, dagger)
        return out

    def wil_mat(self, g, inp, kappa, dagger):
        g = self._cast_gauge(g, inp.dtype)
        out = np.zeros(self.V * 24, dtype=inp.dtype)
        getattr(self.L, "orc_wil_mat" + self._suffix(inp))(_ptr(out), _ptrs(g), _ptr(inp), kappa, dagger)
        return out

    def wil_matpc(self, g, inp, kappa, matpc, dagger):
        g = self._cast_gauge(g, inp.dtype)
        out = np.zeros(self.Vh * 24, dtype=inp.dtype)
        getattr(self.L, "orc_wil_matpc" + self._suffix(inp))(_ptr(out), _ptrs(g), _ptr(inp), kappa, matpc, dagger)
        return out

    # ---- twisted-clover host path (tests/clover_reference.cpp) ----
    def clover(self, norm=0.1, diag=1.0, seed=4242):
        """Random packed clover term of the reference's tests: [V][2 chiralities][6 diag + 15 complex lower-triangular]."""
        c = np.zeros(self.V * 72, dtype=np.float64)
        self.L.orc_construct_clover(_ptr(c), norm, diag, seed)
        return c

    @staticmethod
    def clover_unpack(c):
        """packed [..., 36] -> Hermitian [..., 6, 6] complex (L[k] holds M[row][col], row > col, column by column)."""
        c = np.asarray(c, dtype=np.float64).reshape(-1, 36)
        M = np.zeros((c.shape[0], 6, 6), dtype=np.complex128)
        for i in range(6):
            M[:, i, i] = c[:, i]
        k = 0
        for col in range(6):
            for row in range(col + 1, 6):
                v = c[:, 6 + 2 * k] + 1j * c[:, 6 + 2 * k + 1]
                M[:, row, col] = v
                M[:, col, row] = np.conj(v)
                k += 1
        return M

    @staticmethod
    def clover_pack(M):
        M = np.asarray(M)
        out = np.zeros((M.shape[0], 36), dtype=np.float64)
        for i in range(6):
            out[:, i] = M[:, i, i].real
        k = 0
        for col in range(6):
            for row in range(col + 1, 6):
                out[:, 6 + 2 * k] = M[:, row, col].real
                out[:, 6 + 2 * k + 1] = M[:, row, col].imag
                k += 1
        return out.ravel()

    def clover_inverse(self, c, kappa, mu):
        """(C^2 + (2 kappa mu)^2)^-1 in packed order: the field the reference's loadCloverQuda hands back for twisted clover
        (lib/clover_invert.cu:56-90, mu2 = 4 kappa^2 mu^2, interface_quda.cpp:790)."""
        M = self.clover_unpack(c)
        a2 = (2.0 * kappa * mu) ** 2
        return self.clover_pack(np.linalg.inv(M @ M + a2 * np.eye(6)))

    def _cl(self, c, dtype):
        return c if c is None or c.dtype == dtype else c.astype(dtype)

    def apply_clover(self, c, inp, parity):
        out = np.zeros_like(inp)
        getattr(self.L, "orc_apply_clover" + self._suffix(inp))(_ptr(out), _ptr(self._cl(c, inp.dtype)), _ptr(inp), parity)
        return out

    def tmc_dslash(self, g, inp, c, cinv, kappa, mu, flavor, parity, matpc, dagger):
        g = self._cast_gauge(g, inp.dtype)
        out = np.zeros(self.Vh * 24, dtype=inp.dtype)
        getattr(self.L, "orc_tmc_dslash" + self._suffix(inp))(_ptr(out), _ptrs(g), _ptr(inp), _ptr(self._cl(c, inp.dtype)), _ptr(self._cl(cinv, inp.dtype)),
                                                              kappa, mu, flavor, parity, matpc, dagger)
        return out

    def tmc_mat(self, g, inp, c, kappa, mu, flavor, dagger):
        g = self._cast_gauge(g, inp.dtype)
        out = np.zeros(self.V * 24, dtype=inp.dtype)
        getattr(self.L, "orc_tmc_mat" + self._suffix(inp))(_ptr(out), _ptrs(g), _ptr(self._cl(c, inp.dtype)), _ptr(inp), kappa, mu, flavor, dagger)
        return out

    def tmc_matpc(self, g, inp, c, cinv, kappa, mu, flavor, matpc, dagger):
        g = self._cast_gauge(g, inp.dtype)
        out = np.zeros(self.Vh * 24, dtype=inp.dtype)
        getattr(self.L, "orc_tmc_matpc" + self._suffix(inp))(_ptr(out), _ptrs(g), _ptr(inp), _ptr(self._cl(c, inp.dtype)), _ptr(self._cl(cinv, inp.dtype)),
                                                             kappa, mu, flavor, matpc, dagger)
        return out

    def twist(self, inp, kappa, mu, flavor, dagger, inverse):
        out = np.zeros_like(inp)
        getattr(self.L, "orc_twist_gamma5" + self._suffix(inp))(_ptr(out), _ptr(inp), dagger, kappa, mu, flavor, inp.size // 24, int(inverse))
        return out


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "liboracle.so"])


def load_oracle():
    path = os.path.join(ORACLE_DIR, "liboracle.so")
    if not os.path.exists(path):
        build_oracle()
    return Oracle(C.CDLL(path))


class Ref:
    """numpy face of oracle/_ref/libtmref.so = the reference's own unmodified CPU sources."""

    def __init__(self, lib):
        self.L = lib
        L = lib
        L.tmref_setup.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_double]
        L.tmref_construct_gauge.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint]
        L.tmref_wil_dslash.argtypes = [C.c_void_p] * 3 + [C.c_int] * 3
        L.tmref_tm_dslash.argtypes = [C.c_void_p] * 3 + [C.c_double] * 2 + [C.c_int] * 5
        L.tmref_tm_matpc.argtypes = [C.c_void_p] * 3 + [C.c_double] * 2 + [C.c_int] * 4
        L.tmref_tm_mat.argtypes = [C.c_void_p] * 3 + [C.c_double] * 2 + [C.c_int] * 3
        if hasattr(L, "tmref_tm_ndeg_dslash"):
            L.tmref_tm_ndeg_dslash.argtypes = [C.c_void_p] * 3 + [C.c_double] * 3 + [C.c_int] * 4
            L.tmref_tm_ndeg_matpc.argtypes = [C.c_void_p] * 3 + [C.c_double] * 3 + [C.c_int] * 3
            L.tmref_tm_ndeg_mat.argtypes = [C.c_void_p] * 3 + [C.c_double] * 3 + [C.c_int] * 2
        if hasattr(L, "tmref_tmc_dslash"):
            L.tmref_construct_clover.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_uint]
            L.tmref_apply_clover.argtypes = [C.c_void_p] * 3 + [C.c_int] * 2
            L.tmref_tmc_dslash.argtypes = [C.c_void_p] * 5 + [C.c_double] * 2 + [C.c_int] * 5
            L.tmref_tmc_mat.argtypes = [C.c_void_p] * 4 + [C.c_double] * 2 + [C.c_int] * 3
            L.tmref_tmc_matpc.argtypes = [C.c_void_p] * 5 + [C.c_double] * 2 + [C.c_int] * 4

    def setup(self, X, antiperiodic=True, anisotropy=1.0):
        self.dims = tuple(int(x) for x in X)
        self.V = int(np.prod(self.dims))
        self.Vh = self.V // 2
        self.L.tmref_setup((C.c_int * 4)(*self.dims), int(antiperiodic), anisotropy)

    def gauge(self, kind=1, seed=137, dtype=np.float64):
        g = [np.zeros(self.V * 18, dtype=dtype) for _ in range(4)]
        self.L.tmref_construct_gauge(_ptrs(g), kind, g[0].itemsize, seed)
        return g

    def tm_dslash(self, g, inp, kappa, mu, flavor, parity, matpc, dagger):
        out = np.zeros(self.Vh * 24, dtype=inp.dtype)
        work = inp.copy()  # the reference twists its input in place for some variants
        self.L.tmref_tm_dslash(_ptr(out), _ptrs(g), _ptr(work), kappa, mu, flavor, parity, matpc, dagger, inp.itemsize)
        return out

    def tm_matpc(self, g, inp, kappa, mu, flavor, matpc, dagger):
        out = np.zeros(self.Vh * 24, dtype=inp.dtype)
        work = inp.copy()
        self.L.tmref_tm_matpc(_ptr(out), _ptrs(g), _ptr(work), kappa, mu, flavor, matpc, dagger, inp.itemsize)
        return out

    def tm_mat(self, g, inp, kappa, mu, flavor, dagger):
        out = np.zeros(self.V * 24, dtype=inp.dtype)
        self.L.tmref_tm_mat(_ptr(out), _ptrs(g), _ptr(inp.copy()), kappa, mu, flavor, dagger, inp.itemsize)
        return out

    def tm_ndeg_dslash(self, g, inp, kappa, mu, eps, parity, matpc, dagger):
        out = np.zeros(2 * self.Vh * 24, dtype=inp.dtype)
        self.L.tmref_tm_ndeg_dslash(_ptr(out), _ptrs(g), _ptr(inp.copy()), kappa, mu, eps, parity, matpc, dagger, inp.itemsize)
        return out

    def tm_ndeg_matpc(self, g, inp, kappa, mu, eps, matpc, dagger):
        out = np.zeros(2 * self.Vh * 24, dtype=inp.dtype)
        self.L.tmref_tm_ndeg_matpc(_ptr(out), _ptrs(g), _ptr(inp.copy()), kappa, mu, eps, matpc, dagger, inp.itemsize)
        return out

    def tm_ndeg_mat(self, g, inp, kappa, mu, eps, dagger):
        out = np.zeros(2 * self.V * 24, dtype=inp.dtype)
        self.L.tmref_tm_ndeg_mat(_ptr(out), _ptrs(g), _ptr(inp.copy()), kappa, mu, eps, dagger, inp.itemsize)
        return out

    def clover(self, norm=0.1, diag=1.0, seed=4242):
        c = np.zeros(self.V * 72, dtype=np.float64)
        self.L.tmref_construct_clover(_ptr(c), norm, diag, 8, seed)
        return c

    def tmc_dslash(self, g, inp, c, cinv, kappa, mu, flavor, parity, matpc, dagger):
        out = np.zeros(self.Vh * 24, dtype=inp.dtype)
        self.L.tmref_tmc_dslash(_ptr(out), _ptrs(g), _ptr(inp.copy()), _ptr(c.copy()), _ptr(cinv.copy()), kappa, mu, flavor, parity, matpc, dagger, inp.itemsize)
        return out

    def tmc_mat(self, g, inp, c, kappa, mu, flavor, dagger):
        out = np.zeros(self.V * 24, dtype=inp.dtype)
        self.L.tmref_tmc_mat(_ptr(out), _ptrs(g), _ptr(c.copy()), _ptr(inp.copy()), kappa, mu, flavor, dagger, inp.itemsize)
        return out

    def tmc_matpc(self, g, inp, c, cinv, kappa, mu, flavor, matpc, dagger):
        out = np.zeros(self.Vh * 24, dtype=inp.dtype)
        self.L.tmref_tmc_matpc(_ptr(out), _ptrs(g), _ptr(inp.copy()), _ptr(c.copy()), _ptr(cinv.copy()), kappa, mu, flavor, matpc, dagger, inp.itemsize)
        return out

    def wil_dslash(self, g, inp, parity, dagger):
        out = np.zeros(self.Vh * 24, dtype=inp.dtype)
        self.L.tmref_wil_dslash(_ptr(out), _ptrs(g), _ptr(inp.copy()), parity, dagger, inp.itemsize)
        return out


def load_ref():
    """Returns None when neither the prebuilt library nor the reference tree is available."""
    path = os.path.join(ORACLE_DIR, "_ref", "libtmref.so")
    if not os.path.exists(path):
        if os.path.isdir("/root/reference/tests"):
            subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "_ref/libtmref.so"])
        else:
            return None
    return Ref(C.CDLL(path))


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    if not (np.iscomplexobj(a) or np.iscomplexobj(b)):
        a = a.astype(np.float64)
        b = b.astype(np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))
