"""Multi-rank parity check (launch with torchrun, one rank per GPU): the global lattice is generated
identically on every rank by the oracle, each rank loads its sub-lattice through the C ABI, applies
dslashQuda / MatQuda with NCCL halo exchange, and compares with its slice of the global oracle result.
Prints MULTIGPU_OK <worst rel L2> on rank 0."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib  # noqa: E402

import quda_b200 as q  # noqa: E402
from tests import oracle_util as ou  # noqa: E402

dist_util = importlib.import_module("quda-qkxtm-multigrid_b200.dist")


def main():
    grid = tuple(int(x) for x in os.environ.get("QB_GRID", "").split(",")) if os.environ.get("QB_GRID") else None
    Xl = tuple(int(x) for x in os.environ.get("QB_LOCAL", "8,4,6,8").split(","))
    L = q.lib()
    rank, world, dist = dist_util.init_comms(L, grid)
    grid = grid or dist_util.default_grid(world)
    L.initQudaMemory()
    info = (C.c_int * 10)()
    L.commRankInfoQudaB200(info)
    coords = tuple(info[2:6])
    assert coords == dist_util.rank_coords(rank, grid), (coords, dist_util.rank_coords(rank, grid))
    idx, Xg = dist_util.local_to_global_index(Xl, grid, coords)
    o = ou.load_oracle()
    o.set_dims(Xg)
    g = o.gauge(kind=1, antiperiodic=True, seed=137)
    sp = o.drand(2 * o.Vh * 24, seed=137)
    Vhl = int(np.prod(Xl)) // 2
    gl = [dist_util.slice_field(a, idx, 18) for a in g]
    spl = dist_util.slice_field(sp, idx, 24)
    worst = 0.0
    kappa, mu = 0.1, 0.01
    for prec, tol in ((8, 1e-13), (4, 1e-6), (2, 1e-3)):   # the north_star tolerances (bench parity gate measures 2.8e-16 / 8.5e-8 / 3.0e-5)
        gp = q.gauge_param(Xl, cuda_prec=prec, reconstruct=12)
        L.loadGaugeQuda((C.c_void_p * 4)(*[a.ctypes.data for a in gl]), C.byref(gp))
        for flavor, parity, matpc, dag in ((1, 0, 0, 0), (1, 1, 0, 1), (-1, 0, 2, 1)):
            p = q.invert_param(cuda_prec=prec, flavor=flavor, matpc=matpc, dagger=dag)
            inp = spl[(1 - parity) * Vhl * 24:(2 - parity) * Vhl * 24].copy()
            out = np.zeros(Vhl * 24)
            L.dslashQuda(out.ctypes.data_as(C.c_void_p), inp.ctypes.data_as(C.c_void_p), C.byref(p), parity)
            gin = sp[(1 - parity) * o.Vh * 24:(2 - parity) * o.Vh * 24].copy()
            ref = o.tm_dslash(g, gin, kappa, mu, flavor, parity, matpc, dag)
            full = np.zeros(2 * o.Vh * 24)
            full[parity * o.Vh * 24:(parity + 1) * o.Vh * 24] = ref
            ref_l = dist_util.slice_field(full, idx, 24)[parity * Vhl * 24:(parity + 1) * Vhl * 24]
            err = ou.rel_l2(out, ref_l)
            assert err <= tol, (rank, prec, flavor, parity, matpc, dag, err)
            worst = max(worst, err / tol)
        p = q.invert_param(cuda_prec=prec, solution_type=q.QUDA_MAT_SOLUTION)
        out = np.zeros(2 * Vhl * 24)
        L.MatQuda(out.ctypes.data_as(C.c_void_p), spl.ctypes.data_as(C.c_void_p), C.byref(p))
        ref_l = dist_util.slice_field(o.tm_mat(g, sp, kappa, mu, 1, 0), idx, 24)
        err = ou.rel_l2(out, ref_l)
        assert err <= tol, (rank, prec, "mat", err)
        worst = max(worst, err / tol)
    # a distributed solve: GCR on the even-odd system, global reductions over NCCL
    gp = q.gauge_param(Xl, cuda_prec=8, reconstruct=12, cuda_prec_sloppy=4)
    L.loadGaugeQuda((C.c_void_p * 4)(*[a.ctypes.data for a in gl]), C.byref(gp))
    p = q.invert_param(kappa=0.1, mu=0.1, cuda_prec=8, solution_type=q.QUDA_MAT_SOLUTION)
    p.cuda_prec_sloppy = 4; p.cuda_prec_precondition = 4
    p.solve_type = q.QUDA_DIRECT_PC_SOLVE; p.inv_type = q.QUDA_GCR_INVERTER; p.tol = 1e-9; p.maxiter = 500
    p.gcrNkrylov = 16; p.reliable_delta = 1e-4
    x = np.zeros(2 * Vhl * 24)
    L.invertQuda(x.ctypes.data_as(C.c_void_p), spl.ctypes.data_as(C.c_void_p), C.byref(p))
    # gather the solution on every rank and check the global residual with the oracle
    import torch
    xs = [torch.zeros(2 * Vhl * 24, dtype=torch.float64, device="cuda") for _ in range(world)] if world > 1 else None
    if world > 1:
        dist.all_gather(xs, torch.from_numpy(x).cuda())
        xg = np.zeros(2 * o.Vh * 24).reshape(-1, 24)
        for r in range(world):
            ridx, _ = dist_util.local_to_global_index(Xl, grid, dist_util.rank_coords(r, grid))
            xg[ridx] = xs[r].cpu().numpy().reshape(-1, 24)
        xg = xg.ravel()
    else:
        xg = np.zeros(2 * o.Vh * 24).reshape(-1, 24); xg[idx] = x.reshape(-1, 24); xg = xg.ravel()
    res = np.linalg.norm(sp - o.tm_mat(g, xg, 0.1, 0.1, 1, 0)) / np.linalg.norm(sp)
    assert res < 5e-9, res
    if rank == 0:
        print(f"MULTIGPU_OK ranks={world} grid={grid} local={Xl} worst_err/tol={worst:.3f} solve_res={res:.2e} iters={p.iter}", flush=True)
    L.endQuda()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
