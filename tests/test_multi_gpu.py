"""Multi-rank tests.
 * CPU (gloo, world_size 2): the rank <-> coordinate map, the local <-> global site maps and the launcher
   plumbing (no CUDA call is made; libquda_b200.so only does rank bookkeeping).
 * GPU (needs >= 2 devices, skipped otherwise): tests/multi_gpu_dslash.py under torchrun -- NCCL halo exchange,
   distributed reductions, a distributed GCR solve, all checked against the global CPU oracle.
 * GPU (1 device): face-index map of the pack kernel against a restatement of the reference's
   indexFromFaceIndex (lib/dslash_index.cuh:13-96): bit-exact.
"""
import ctypes as C
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dist_util = importlib.import_module("quda-qkxtm-multigrid_b200.dist")

GLOO_CHILD = r"""
import sys, os, ctypes as C, importlib
sys.path.insert(0, %(root)r)
import numpy as np
import torch, torch.distributed as dist
import quda_b200 as q
du = importlib.import_module("quda-qkxtm-multigrid_b200.dist")
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
L = q.lib()
grid = tuple(int(x) for x in os.environ["QB_GRID"].split(","))
L.commsBootstrapQudaB200(rank, world, None)
L.initCommsGridQuda(4, (C.c_int*4)(*grid), None, None)
info = (C.c_int*10)(); L.commRankInfoQudaB200(info)
assert info[0] == rank and info[1] == world and tuple(info[6:10]) == grid
coords = tuple(info[2:6])
assert coords == du.rank_coords(rank, grid), (coords, du.rank_coords(rank, grid))
# every global site is owned by exactly one rank
Xl = (4, 2, 4, 6)
idx, Xg = du.local_to_global_index(Xl, grid, coords)
mine = torch.zeros(int(np.prod(Xg)), dtype=torch.int32); mine[torch.from_numpy(idx)] = 1
dist.all_reduce(mine)
assert int(mine.min()) == 1 and int(mine.max()) == 1
# a field sliced per rank and re-assembled is the global field
g = torch.arange(int(np.prod(Xg)) * 3, dtype=torch.float64).reshape(-1, 3)
loc = g[torch.from_numpy(idx)]
parts = [torch.zeros_like(loc) for _ in range(world)]
dist.all_gather(parts, loc)
back = torch.zeros_like(g)
for r in range(world):
    ridx, _ = du.local_to_global_index(Xl, grid, du.rank_coords(r, grid))
    back[torch.from_numpy(ridx)] = parts[r]
assert torch.equal(back, g)
if rank == 0: print("GLOO_OK")
dist.destroy_process_group()
"""


@pytest.mark.parametrize("grid", ["1,1,1,2", "1,1,2,1", "2,1,1,1"])
def test_rank_grid_and_site_maps_gloo_world2(grid):
    env = dict(os.environ, QB_GRID=grid, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29500 + (hash(grid) % 400)), "-c", GLOO_CHILD % {"root": ROOT}]
    # torchrun cannot take -c; write the child to a temp file instead
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as f:
        f.write(GLOO_CHILD % {"root": ROOT})
        path = f.name
    cmd = cmd[:-2] + [path]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=300)
    os.unlink(path)
    assert r.returncode == 0 and "GLOO_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_rank_coords_t_fastest():
    # rank = ((c0*g1 + c1)*g2 + c2)*g3 + c3   (lib/interface_quda.cpp:261-274)
    grid = (1, 2, 2, 4)
    seen = set()
    for r in range(16):
        c = dist_util.rank_coords(r, grid)
        assert ((c[0] * grid[1] + c[1]) * grid[2] + c[2]) * grid[3] + c[3] == r
        seen.add(c)
    assert len(seen) == 16
    assert dist_util.rank_coords(1, grid) == (0, 0, 0, 1)
    assert dist_util.default_grid(8) == (1, 1, 2, 4) and dist_util.default_grid(4) == (1, 1, 1, 4)


def ref_index_from_face_index(face_idx, dim, face_num, parity, X):
    """Restatement of indexFromFaceIndex<4, QUDA_4D_PC, dim, nLayers=1, face_num> (lib/dslash_index.cuh:13-96)."""
    fX = [X[0], X[1], X[2], X[3]]
    fX[dim] = 1
    face_X, face_Y, face_Z = fX[0], fX[1], fX[2]
    face_XY, face_XYZ = face_X * face_Y, face_X * face_Y * face_Z
    face_parity = (parity + face_num * (X[dim] - 1)) & 1
    f = 2 * face_idx
    if not (face_X & 1):
        aux1 = f // face_X; aux2 = aux1 // face_Y; aux3 = aux2 // face_Z
        y = aux1 - aux2 * face_Y; z = aux2 - aux3 * face_Z; t = aux3
        f += (face_parity + t + z + y) & 1
    elif not (face_Y & 1):
        t = f // face_XYZ; z = (f // face_XY) % face_Z
        f += (face_parity + t + z) & 1
    elif not (face_Z & 1):
        t = f // face_XYZ
        f += (face_parity + t) & 1
    else:
        f += face_parity
    gap = X[dim] - 1
    idx = f
    if dim == 0:
        idx += (f // face_X + face_num) * gap
    elif dim == 1:
        idx += (f // face_XY + face_num) * gap * face_X
    elif dim == 2:
        idx += (f // face_XYZ + face_num) * gap * face_XY
    else:
        idx += face_num * gap * face_XYZ
    return idx >> 1


@pytest.mark.gpu
def test_face_index_map_bit_exact(quda, oracle):
    q, L = quda, quda.lib()
    for X in ((8, 4, 6, 8), (4, 4, 4, 4), (6, 2, 8, 4)):
        oracle.set_dims(X)
        g = oracle.gauge(kind=0)
        gp = q.gauge_param(X, cuda_prec=4, reconstruct=18)
        L.loadGaugeQuda((C.c_void_p * 4)(*[a.ctypes.data for a in g]), C.byref(gp))
        V = int(np.prod(X))
        for dim in range(4):
            fv = V // X[dim] // 2
            for face_num in (0, 1):
                for parity in (0, 1):
                    out = (C.c_int * fv)()
                    L.faceIndexMapQudaB200(dim, face_num, parity, out)
                    ref = [ref_index_from_face_index(f, dim, face_num, parity, X) for f in range(fv)]
                    assert list(out) == ref, (X, dim, face_num, parity)


@pytest.mark.gpu
@pytest.mark.parametrize("nranks,grid,local,peer", [(2, "1,1,1,2", "8,4,6,8", 1), (2, "1,1,2,1", "4,4,4,8", 1), (4, "1,1,2,2", "4,4,4,4", 1),
                                                    (2, "1,1,1,2", "16,16,16,32", 1),   # large enough for the slab-pipelined host path
                                                    (2, "1,1,1,2", "8,4,6,8", 0), (2, "1,1,2,1", "16,16,16,32", 0), (4, "1,1,2,2", "4,4,4,4", 0)])
def test_nccl_halo_exchange_vs_global_oracle(nranks, grid, local, peer):
    """peer = 1: faces stored by the pack kernel straight into the neighbour's ghost zone (CUDA IPC over NVLink, arrival flags);
    peer = 0: NCCL send / recv groups"""
    import torch
    if torch.cuda.device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    env = dict(os.environ, QB_GRID=grid, QB_LOCAL=local, QB_PEER_HALO=str(peer))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks), "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "multi_gpu_dslash.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MULTIGPU_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("nranks,grid,pc,peer", [(2, "1,1,1,2", 0, 1), (2, "1,1,2,1", 0, 1), (4, "1,1,2,2", 0, 1),
                                                 (2, "1,1,1,2", 1, 1), (2, "1,1,2,1", 1, 0), (2, "1,1,1,2", 0, 0)])
def test_distributed_multigrid_solve(nranks, grid, pc, peer):
    """pc = 1: hierarchy on the even-odd system + QUDA_DIRECT_PC_SOLVE; peer = 1: all-reduces fused into the reduction kernels over the
    NVLink peer mailboxes and fine + coarse halo faces stored straight into the neighbours' ghost zones (comm.h), 0: ncclAllReduce and
    NCCL send / recv groups on the compute stream"""
    import torch
    if torch.cuda.device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    env = dict(os.environ, QB_GRID=grid, QB_LOCAL="8,8,8,8", QB_MG_PC=str(pc), QB_PEER_REDUCE=str(peer), QB_PEER_HALO=str(peer))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks), "--master-addr", "127.0.0.1",
           "--master-port", "29633", os.path.join(ROOT, "tests", "multi_gpu_mg.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MULTIGPU_MG_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.gpu
@pytest.mark.parametrize("nranks,grid,peer,pc", [(2, "1,1,1,2", 1, 0), (2, "1,1,2,1", 1, 0), (4, "1,1,2,2", 1, 0), (2, "1,1,2,1", 0, 0), (2, "1,1,1,2", 1, 1)])
def test_distributed_block_multigrid(nranks, grid, peer, pc):
    """BASELINE config 5's "multi-RHS coarse grid" on a lattice partitioned over real ranks: batched coarse null-vector setup and the
    block multigrid behind invertMultiSrcQuda with the ghost zones of block fields filled by the neighbours' pack kernels over NVLink peer
    memory (peer = 1) or exchanged over NCCL (peer = 0); no fallback to one-at-a-time"""
    import torch
    if torch.cuda.device_count() < nranks:
        pytest.skip(f"needs {nranks} GPUs")
    # pc = 1: even-odd hierarchy and QUDA_DIRECT_PC_SOLVE (level-0 smoother of all sources in lock-step on batch fields, whose hops go
    # member by member through the overlapped halo path on a partitioned lattice)
    env = dict(os.environ, QB_GRID=grid, QB_LOCAL="8,8,8,8", QB_MG_MULTISRC="1", QB_PEER_HALO=str(peer), QB_MG_PC=str(pc))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks), "--master-addr", "127.0.0.1",
           "--master-port", "29655", os.path.join(ROOT, "tests", "multi_gpu_mg.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MULTIGPU_MG_MULTISRC_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    assert "null vectors from one batched BiCGStab on the multi-RHS tensor-core operator" in r.stdout, r.stdout[-3000:]
    assert "invertMultiSrcQuda: block of 3 sources" in r.stdout, r.stdout[-3000:]


@pytest.mark.gpu
def test_null_vectors_saved_on_two_ranks_load_on_one(quda, oracle, tmp_path):
    """the near-null vector container is independent of the rank layout: a 3-level setup saved by 2 ranks (T split) is loaded by this
    single-rank process (compute_null_vector = NO) and preconditions the same solve in the same number of iterations"""
    import ctypes as C
    import re
    import numpy as np
    import torch
    from tests.test_multigrid_gpu import host_residual, load_gauge, mg_inv_param, vp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    base = str(tmp_path / "nv2")
    env = dict(os.environ, QB_GRID="1,1,1,2", QB_LOCAL="8,8,8,8", QB_MG_VECFILE=base)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29677", os.path.join(ROOT, "tests", "multi_gpu_mg.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MULTIGPU_MG_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    it2 = int(re.search(r"mg_iters=(\d+)", r.stdout).group(1))
    assert os.path.exists(base + "_level_0") and os.path.exists(base + "_level_1")
    q, L = quda, quda.lib()
    X, kappa, mu = (8, 8, 8, 16), 0.1245, 0.005
    oracle.set_dims(X)
    g = oracle.weak_gauge(eps=0.25, antiperiodic=True, seed=4711)
    load_gauge(q, g, X, prec=8, sloppy=4, precond=4, recon=12, antiperiodic=True)
    ip = mg_inv_param(q, kappa, mu)
    mgp = q.multigrid_param(ip, n_level=3, geo_block=((2, 2, 2, 2), (2, 2, 2, 2)), n_vec=(8, 8), setup_maxiter=100, setup_tol=5e-6)
    mgp.compute_null_vector = q.QUDA_COMPUTE_NULL_VECTOR_NO
    mgp.vec_infile = base.encode()
    mg = L.newMultigridQuda(C.byref(mgp))
    b = oracle.drand(2 * oracle.Vh * 24, seed=11)
    x = np.zeros_like(b)
    p = mg_inv_param(q, kappa, mu)
    p.inv_type_precondition = q.QUDA_MG_INVERTER
    p.preconditioner = mg
    p.gcrNkrylov = 20; p.tol = 1e-8; p.maxiter = 2000; p.reliable_delta = 1e-4
    L.invertQuda(vp(x), vp(b), C.byref(p))
    res = host_residual(oracle, g, x, b, kappa, mu)
    L.destroyMultigridQuda(mg)
    assert res < 5e-8 and abs(p.iter - it2) <= 2, (res, p.iter, it2)
