"""Generates tests/golden/tm_ndeg_ref_4448.npz from the reference's own CPU code (tm_ndeg_dslash / tm_ndeg_matpc / tm_ndeg_mat of
tests/wilson_dslash_reference.cpp:461-587, compiled unmodified into oracle/_ref/libtmref.so).  Runs ONLY in the build container.
Inputs: reference gauge generator with srand(137), LCG spinor seeded 137 (a full doublet field), lattice 4x4x4x8, kappa = 0.1,
mu = 0.01, epsilon = 0.03, antiperiodic T, fp64."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import oracle_util  # noqa: E402

X = (4, 4, 4, 8)
KAPPA, MU, EPS = 0.1, 0.01, 0.03


def main():
    ref = oracle_util.load_ref()
    assert ref is not None, "reference tree not available"
    orc = oracle_util.load_oracle()
    ref.setup(X, antiperiodic=True)
    orc.set_dims(X)
    g = ref.gauge(kind=1, seed=137)
    sp = orc.drand(4 * ref.Vh * 24, seed=137)
    even = sp[: 2 * ref.Vh * 24].copy()
    out = {"X": np.array(X), "kappa": KAPPA, "mu": MU, "epsilon": EPS}
    out["dslash_p0_m0_d0"] = ref.tm_ndeg_dslash(g, even, KAPPA, MU, EPS, 0, 0, 0)
    out["dslash_p1_m0_d1"] = ref.tm_ndeg_dslash(g, even, KAPPA, MU, EPS, 1, 0, 1)
    out["matpc_m0_d0"] = ref.tm_ndeg_matpc(g, even, KAPPA, MU, EPS, 0, 0)
    out["matpc_m2_d1"] = ref.tm_ndeg_matpc(g, even, KAPPA, MU, EPS, 2, 1)
    out["mat_d0"] = ref.tm_ndeg_mat(g, sp, KAPPA, MU, EPS, 0)
    out["mat_d1"] = ref.tm_ndeg_mat(g, sp, KAPPA, MU, EPS, 1)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tm_ndeg_ref_4448.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
