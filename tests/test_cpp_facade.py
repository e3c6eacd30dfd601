"""The `namespace quda` C++ facade (include/quda_cpp.h, SURVEY 8(b) second boundary): a C++ program written like the QKXTM code inside
the reference library is compiled with plain g++ against the header and libquda_b200.so, run on the GPU, and its solution checked
against the oracle's host operator."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "facade_solve.cpp")
PKG = os.path.join(ROOT, "quda-qkxtm-multigrid_b200")


def build_exe(out_dir):
    exe = os.path.join(out_dir, "facade_solve")
    cmd = ["g++", "-std=c++11", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe, "-L", PKG, "-l:libquda_b200.so",
           "-Wl,-rpath," + PKG, "-Wl,-rpath,/usr/local/cuda/lib64", "-Wl,--allow-shlib-undefined"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def test_facade_header_compiles_and_links_without_cuda(tmp_path):
    """CPU check: the header is plain C++11 (no CUDA / torch types) and every facade symbol resolves against the library."""
    exe = build_exe(str(tmp_path))
    assert os.path.exists(exe)
    r = subprocess.run(["nm", "-D", "--defined-only", os.path.join(PKG, "libquda_b200.so")], capture_output=True, text=True)
    for sym in ("createDirac", "ColorSpinorField6Create", "Solver6create", "Dirac7prepare", "Dirac11reconstruct", "massRescale", "blas5norm2"):
        assert sym in r.stdout, sym


@pytest.mark.gpu
@pytest.mark.parametrize("pc_solve", [1, 0])
def test_facade_solve_matches_oracle(oracle, tmp_path, pc_solve):
    X, kappa, mu = (8, 8, 8, 8), 0.12, 0.02
    oracle.set_dims(X)
    g = oracle.weak_gauge(eps=0.3, antiperiodic=False, seed=99)
    b = np.random.default_rng(5).standard_normal(oracle.V * 24)
    np.concatenate([np.ascontiguousarray(a, dtype=np.float64).ravel() for a in g]).tofile(tmp_path / "gauge.bin")
    b.tofile(tmp_path / "src.bin")
    exe = build_exe(str(tmp_path))
    r = subprocess.run([exe, *map(str, X), str(kappa), str(mu), str(pc_solve), str(tmp_path / "gauge.bin"), str(tmp_path / "src.bin"), str(tmp_path / "sol.bin")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    print(r.stdout)
    x = np.fromfile(tmp_path / "sol.bin")
    res = np.linalg.norm(b - oracle.tm_mat(g, x, kappa, mu, 1, 0)) / np.linalg.norm(b)
    line = [l for l in r.stdout.splitlines() if l.startswith("FACADE iter")][0].split()
    norms = [l for l in r.stdout.splitlines() if l.startswith("FACADE norm2")][0].split()
    assert abs(float(norms[3]) - float(norms[5])) < 1e-10 * float(norms[5])
    assert res < 5e-9 and float(line[4]) < 5e-9 and float(line[6]) < 5e-9 and int(line[2]) > 0, (res, line)
