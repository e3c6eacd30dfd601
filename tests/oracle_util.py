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


# QudaDiracType values of the reference's include/enum_quda.h (the `dirac` argument of CoarseOp / calculateY)
def _ref_enum(name):
    """value of an enumerator of the reference's include/enum_quda.h, read from the table written when libmgref.so was built"""
    return MGREF_ENUMS[name]


MGREF_ENUMS = {}


class MgRef:
    """numpy face of oracle/_ref/libmgref.so = the reference's own multigrid HOST code (lib/transfer.cpp, transfer_util.cu,
    prolongator.cu, restrictor.cu, coarse_op.cu(h), coarsecoarse_op.cu, dslash_coarse.cu), driven by oracle/mg_ref_shim.cpp.
    Host field order everywhere: [parity][x_cb][spin][colour][re, im] fp32, DeGrand-Rossi basis (the `generic` order of the
    mg*QudaB200 test hooks)."""

    def __init__(self, lib):
        self.L = L = lib
        vp, ip = C.c_void_p, C.POINTER(C.c_int)
        L.mgref_transfer_new.argtypes = [vp, C.c_int, ip, C.c_int, C.c_int, ip, C.c_int]
        L.mgref_transfer_new.restype = vp
        L.mgref_transfer_free.argtypes = [vp]
        L.mgref_transfer_V.argtypes = [vp, vp]
        L.mgref_P.argtypes = [vp, vp, vp, C.c_int]
        L.mgref_R.argtypes = [vp, vp, vp, C.c_int]
        L.mgref_coarse_op.argtypes = [vp, vp, C.c_double, C.c_double, C.c_int, C.c_int]
        L.mgref_coarse_op.restype = vp
        L.mgref_coarse_coarse_op.argtypes = [vp, vp, C.c_double, C.c_int, C.c_int]
        L.mgref_coarse_coarse_op.restype = vp
        L.mgref_coarse_free.argtypes = [vp]
        L.mgref_coarse_dims.argtypes = [vp, ip]
        L.mgref_coarse_links.argtypes = [vp, C.c_int, vp]
        L.mgref_apply_coarse.argtypes = [vp, vp, vp, vp, C.c_double] + [C.c_int] * 5
        L.mgref_enum.argtypes = [C.c_char_p]
        L.mgref_enum.restype = C.c_int
        for n in ("QUDA_WILSON_DIRAC", "QUDA_TWISTED_MASS_DIRAC", "QUDA_TWISTED_MASSPC_DIRAC", "QUDA_COARSE_DIRAC", "QUDA_COARSEPC_DIRAC",
                  "QUDA_MATPC_EVEN_EVEN", "QUDA_MATPC_ODD_ODD", "QUDA_MATPC_INVALID"):
            MGREF_ENUMS[n] = L.mgref_enum(n.encode())

    class Transfer:
        def __init__(self, ref, B, X, nspin, ncolor, geo_bs, spin_bs):
            """B: list of null vectors (full host fields, fp32)"""
            self.ref, self.X, self.nspin, self.ncolor, self.nvec, self.spin_bs = ref, tuple(X), nspin, ncolor, len(B), spin_bs
            Bs = np.ascontiguousarray(np.stack([np.asarray(b, dtype=np.float32) for b in B]))
            bs = (C.c_int * 4)(*geo_bs)
            self.h = ref.L.mgref_transfer_new(_ptr(Bs), self.nvec, (C.c_int * 4)(*X), nspin, ncolor, bs, spin_bs)
            self.geo_bs = tuple(bs)
            self.Xc = tuple(X[d] // self.geo_bs[d] for d in range(4))
            self.nf = int(np.prod(X)) * nspin * ncolor          # complex components of a fine field
            self.nc = int(np.prod(self.Xc)) * (nspin // spin_bs) * self.nvec

        def V(self):
            out = np.zeros(2 * self.nf * self.nvec, dtype=np.float32)
            self.ref.L.mgref_transfer_V(self.h, _ptr(out))
            return out

        def P(self, coarse, parity=-1):
            out = np.zeros(2 * self.nf // (2 if parity >= 0 else 1), dtype=np.float32)
            coarse = np.ascontiguousarray(coarse, dtype=np.float32)
            self.ref.L.mgref_P(self.h, _ptr(out), _ptr(coarse), parity)
            return out

        def R(self, fine, parity=-1):
            out = np.zeros(2 * self.nc, dtype=np.float32)
            fine = np.ascontiguousarray(fine, dtype=np.float32)
            self.ref.L.mgref_R(self.h, _ptr(out), _ptr(fine), parity)
            return out

        def free(self):
            self.ref.L.mgref_transfer_free(self.h)

    class Coarse:
        def __init__(self, ref, h):
            self.ref, self.h = ref, h
            info = (C.c_int * 5)()
            ref.L.mgref_coarse_dims(h, info)
            self.Xc, self.N = tuple(info[0:4]), info[4]
            self.V = int(np.prod(self.Xc))

        def links(self, which):
            """which: 'Y' | 'X' | 'Xinv' | 'Yhat' -> complex array [dir][parity*Vh + x_cb][row][col] (dir axis only for Y / Yhat)"""
            w = {"Y": 0, "X": 1, "Xinv": 2, "Yhat": 3}[which]
            geo = 8 if w in (0, 3) else 1
            out = np.zeros(geo * self.V * self.N * self.N * 2, dtype=np.float32)
            self.ref.L.mgref_coarse_links(self.h, w, _ptr(out))
            z = (out[0::2] + 1j * out[1::2]).reshape(geo, self.V, self.N, self.N)
            return z if geo == 8 else z[0]

        def apply(self, inA, kappa, inB=None, parity=-1, dslash=True, clover=True, yhat=False, xinv=False):
            """ApplyCoarse (lib/dslash_coarse.cu:806-814) on host fields"""
            inA = np.ascontiguousarray(inA, dtype=np.float32)
            inB = inA if inB is None else np.ascontiguousarray(inB, dtype=np.float32)
            out = np.zeros(self.V * self.N * 2 // (2 if parity >= 0 else 1), dtype=np.float32)
            self.ref.L.mgref_apply_coarse(self.h, _ptr(out), _ptr(inA), _ptr(inB), kappa, parity, int(dslash), int(clover), int(yhat), int(xinv))
            return out

        def free(self):
            self.ref.L.mgref_coarse_free(self.h)

    def transfer(self, B, X, nspin, ncolor, geo_bs, spin_bs):
        return MgRef.Transfer(self, B, X, nspin, ncolor, geo_bs, spin_bs)

    def coarse_op(self, T, gauge32, kappa, mu_arg, dirac, matpc="QUDA_MATPC_INVALID"):
        """CoarseOp's host worker calculateY (lib/coarse_op.cu:152-213, coarse_op.cuh:1309-1497); gauge32: 4 fp32 QDP arrays"""
        g = [np.ascontiguousarray(a, dtype=np.float32) for a in gauge32]
        self._keep = g
        return MgRef.Coarse(self, self.L.mgref_coarse_op(T.h, _ptrs(g), kappa, mu_arg, _ref_enum(dirac), _ref_enum(matpc)))

    def coarse_coarse_op(self, T, fine, kappa, pc=False, matpc="QUDA_MATPC_EVEN_EVEN"):
        return MgRef.Coarse(self, self.L.mgref_coarse_coarse_op(T.h, fine.h, kappa, int(pc), _ref_enum(matpc)))


def load_mgref():
    path = os.path.join(ORACLE_DIR, "_ref", "libmgref.so")
    if not os.path.exists(path):
        if os.path.isdir("/root/reference/lib"):
            subprocess.check_call(["make", "-s", "-j", "8", "-C", ORACLE_DIR, "_ref/libmgref.so"])
        else:
            return None
    return MgRef(C.CDLL(path))
