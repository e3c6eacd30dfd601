"""Multi-GPU launch helpers: one process per GPU (torchrun), torch.distributed for the rendezvous,
NCCL inside libquda_b200.so for halos and reductions.  Replaces the MPI/QMP start-up of the reference's
tests (tests/test_util.cpp:44-67 initComms) -- the process grid semantics are the reference's
(initCommsGridQuda, rank lexicographic in the grid coordinates with t fastest)."""
import ctypes as C
import os

import numpy as np


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def default_grid(world):
    """T first, then Z (BASELINE.json north_star): (1,1,1,N) up to 4 ranks, (1,1,2,4) for 8."""
    if world <= 4:
        return (1, 1, 1, world)
    if world == 8:
        return (1, 1, 2, 4)
    z = 2
    return (1, 1, z, world // z)


def init_comms(L, grid=None, backend=None):
    """Initialise device, NCCL communicator and the process grid on every rank.  Returns (rank, world, dist or None)."""
    rank, world, local_rank = env_rank()
    L.initQudaDevice(local_rank)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        cuda = torch.cuda.is_available()
        if not dist.is_initialized():
            dist.init_process_group(backend or ("nccl" if cuda else "gloo"),
                                    **({"device_id": torch.device("cuda", local_rank)} if cuda else {}))
        uid = torch.zeros(128, dtype=torch.uint8)
        if cuda:
            if rank == 0:
                buf = (C.c_char * 128)()
                L.ncclUniqueIdQudaB200(buf)
                uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
            dev = torch.device("cuda", local_rank)
            uid = uid.to(dev)
            dist.broadcast(uid, 0)
            L.commsBootstrapQudaB200(rank, world, uid.cpu().numpy().tobytes())
        else:
            L.commsBootstrapQudaB200(rank, world, None)  # CPU: rank bookkeeping only
        grid = grid or default_grid(world)
        assert int(np.prod(grid)) == world, (grid, world)
        L.initCommsGridQuda(4, (C.c_int * 4)(*grid), None, None)
    return rank, world, dist


def rank_coords(rank, grid):
    """coords of `rank` in the reference's default map: rank = ((c0*g1 + c1)*g2 + c2)*g3 + c3."""
    c = [0, 0, 0, 0]
    r = rank
    for d in (3, 2, 1, 0):
        c[d] = r % grid[d]
        r //= grid[d]
    return tuple(c)


def cb_coords(cb, parity, X):
    cb = np.asarray(cb)
    za = cb // (X[0] // 2)
    zb = za // X[1]
    y = za - zb * X[1]
    t = zb // X[2]
    z = zb - t * X[2]
    x = 2 * cb + ((y + z + t + parity) & 1) - za * X[0]
    return x, y, z, t


def local_to_global_index(Xl, grid, coords):
    """For every local full index (parity*Vh + cb) the global full index of the same site."""
    Xg = tuple(Xl[d] * grid[d] for d in range(4))
    Vhl = int(np.prod(Xl)) // 2
    Vhg = int(np.prod(Xg)) // 2
    out = np.empty(2 * Vhl, dtype=np.int64)
    for par in (0, 1):
        x, y, z, t = cb_coords(np.arange(Vhl), par, Xl)
        gx, gy, gz, gt = x + coords[0] * Xl[0], y + coords[1] * Xl[1], z + coords[2] * Xl[2], t + coords[3] * Xl[3]
        lex = ((gt * Xg[2] + gz) * Xg[1] + gy) * Xg[0] + gx
        gpar = (gx + gy + gz + gt) & 1
        out[par * Vhl:(par + 1) * Vhl] = gpar * Vhg + (lex >> 1)
    return out, Xg


def slice_field(global_field, index, per_site):
    """Rows `index` of a [2*Vh_global][per_site] field."""
    return global_field.reshape(-1, per_site)[index].ravel().copy()
