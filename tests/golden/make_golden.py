"""Generates tests/golden/tm_ref_4448.npz from the reference's own CPU code.

Runs ONLY in the build container (needs /root/reference to build oracle/_ref/libtmref.so).
The vectors are outputs of the reference's unmodified tests/wilson_dslash_reference.cpp on inputs
from the reference's generators (tests/test_util.cpp:879-925 gauge with srand(137);
lib/comm_common.cpp:73-87 LCG spinor seeded 137), lattice 4x4x4x8, kappa=0.1, mu=0.01,
antiperiodic T, fp64.  Commit the .npz; the GPU box has no /root/reference.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import oracle_util  # noqa: E402

X = (4, 4, 4, 8)
KAPPA, MU = 0.1, 0.01


def main():
    ref = oracle_util.load_ref()
    assert ref is not None, "reference tree not available"
    orc = oracle_util.load_oracle()
    ref.setup(X, antiperiodic=True)
    orc.set_dims(X)
    g = ref.gauge(kind=1, seed=137)
    sp = orc.drand(2 * ref.Vh * 24, seed=137)
    even = sp[: ref.Vh * 24].copy()
    out = {"X": np.array(X), "kappa": KAPPA, "mu": MU,
           "gauge_sum": np.array([sum(float(a.sum()) for a in g)]),
           "gauge_head": np.concatenate([a[:18] for a in g]),
           "spinor_head": sp[:24].copy()}
    for flavor in (1, -1):
        for parity in (0, 1):
            for matpc in (0, 2):
                for dag in (0, 1):
                    out[f"dslash_f{flavor}_p{parity}_m{matpc}_d{dag}"] = ref.tm_dslash(g, even, KAPPA, MU, flavor, parity, matpc, dag)
    for matpc in (0, 1, 2, 3):
        for dag in (0, 1):
            out[f"matpc_m{matpc}_d{dag}"] = ref.tm_matpc(g, even, KAPPA, MU, 1, matpc, dag)
    for dag in (0, 1):
        out[f"mat_d{dag}"] = ref.tm_mat(g, sp, KAPPA, MU, 1, dag)
        out[f"wil_dslash_d{dag}"] = ref.wil_dslash(g, even, 0, dag)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tm_ref_4448.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
