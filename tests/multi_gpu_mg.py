"""Multi-rank multigrid check (torchrun, one rank per GPU): distributed 3-level MG-GCR solve on a lattice
partitioned over the ranks; the solution is gathered and its residual is computed with the global CPU oracle
(the host check of tests/multigrid_invert_test.cpp:529-577).  Prints MULTIGPU_MG_OK on rank 0."""
import ctypes as C
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quda_b200 as q  # noqa: E402
from tests import oracle_util as ou  # noqa: E402

dist_util = importlib.import_module("quda-qkxtm-multigrid_b200.dist")


def main():
    grid = tuple(int(x) for x in os.environ["QB_GRID"].split(",")) if os.environ.get("QB_GRID") else None
    Xl = tuple(int(x) for x in os.environ.get("QB_LOCAL", "8,8,8,8").split(","))
    L = q.lib()
    rank, world, dist = dist_util.init_comms(L, grid)
    grid = grid or dist_util.default_grid(world)
    L.initQudaMemory()
    coords = dist_util.rank_coords(rank, grid)
    idx, Xg = dist_util.local_to_global_index(Xl, grid, coords)
    o = ou.load_oracle()
    o.set_dims(Xg)
    kappa, mu = 0.1245, 0.005
    g = o.weak_gauge(eps=0.25, antiperiodic=True, seed=4711)
    gl = [dist_util.slice_field(a, idx, 18) for a in g]
    gp = q.gauge_param(Xl, cuda_prec=8, reconstruct=12, cuda_prec_sloppy=4, cuda_prec_precondition=4)
    L.loadGaugeQuda((C.c_void_p * 4)(*[a.ctypes.data for a in gl]), C.byref(gp))

    def inv_param():
        p = q.invert_param(kappa=kappa, mu=mu, cuda_prec=8, solution_type=q.QUDA_MAT_SOLUTION)
        p.cuda_prec_sloppy = 4; p.cuda_prec_precondition = 4
        p.solve_type = q.QUDA_DIRECT_SOLVE; p.inv_type = q.QUDA_GCR_INVERTER
        p.gcrNkrylov = 20; p.tol = 1e-8; p.maxiter = 2000; p.reliable_delta = 1e-4
        return p

    pc = os.environ.get("QB_MG_PC") == "1"   # hierarchy coarsened on the even-odd system + even-odd outer solve (the reference's default)
    ip = inv_param()
    if os.environ.get("QB_MG_MULTISRC") == "1":
        ip.verbosity = q.QUDA_SUMMARIZE
    mgp = q.multigrid_param(ip, n_level=3, geo_block=((2, 2, 2, 2), (2, 2, 2, 2)), n_vec=(8, 8), setup_maxiter=100, setup_tol=5e-6,
                            solve_type=q.QUDA_DIRECT_PC_SOLVE if pc else q.QUDA_DIRECT_SOLVE)
    if os.environ.get("QB_MG_VECFILE"):
        mgp.vec_outfile = os.environ["QB_MG_VECFILE"].encode()
    mg = L.newMultigridQuda(C.byref(mgp))
    for lvl in (0, 1):
        dev = (C.c_double * 3)()
        L.mgVerifyQudaB200(mg, lvl, dev)
        assert dev[0] < 5e-6 and dev[1] < 1e-4 and dev[2] < 5e-5, (rank, lvl, list(dev))
    bg = o.drand(2 * o.Vh * 24, seed=11)
    bl = dist_util.slice_field(bg, idx, 24)
    x = np.zeros_like(bl)
    p = inv_param()
    p.inv_type_precondition = q.QUDA_MG_INVERTER
    p.preconditioner = mg
    if pc:
        p.solve_type = q.QUDA_DIRECT_PC_SOLVE
    L.invertQuda(x.ctypes.data_as(C.c_void_p), bl.ctypes.data_as(C.c_void_p), C.byref(p))
    p0 = inv_param()
    if pc:
        p0.solve_type = q.QUDA_DIRECT_PC_SOLVE
    x0 = np.zeros_like(bl)
    L.invertQuda(x0.ctypes.data_as(C.c_void_p), bl.ctypes.data_as(C.c_void_p), C.byref(p0))
    import torch
    if world > 1:
        xs = [torch.zeros(x.size, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(xs, torch.from_numpy(x).cuda())
        xg = np.zeros(2 * o.Vh * 24).reshape(-1, 24)
        for r in range(world):
            ridx, _ = dist_util.local_to_global_index(Xl, grid, dist_util.rank_coords(r, grid))
            xg[ridx] = xs[r].cpu().numpy().reshape(-1, 24)
        xg = xg.ravel()
    else:
        xg = np.zeros(2 * o.Vh * 24).reshape(-1, 24); xg[idx] = x.reshape(-1, 24); xg = xg.ravel()
    res = np.linalg.norm(bg - o.tm_mat(g, xg, kappa, mu, 1, 0)) / np.linalg.norm(bg)
    assert res < 5e-8, res
    if world > 1:
        # every rank must have seen bit-identical global sums (the all-reduce adds the ranks' contributions in a fixed order)
        tr = [torch.zeros(2, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(tr, torch.tensor([p.true_res, float(p.iter)], dtype=torch.float64, device="cuda"))
        assert all(bool((t == tr[0]).all()) for t in tr), [t.tolist() for t in tr]
    assert p.iter < p0.iter / 2, (p.iter, p0.iter)
    if os.environ.get("QB_MG_MULTISRC") == "1":
        # block multigrid on the partitioned lattice (BASELINE config 5: multi-RHS coarse grid): ghost zones of block fields in the
        # tensor-core coarse operator, global block reductions; every solution checked with the global host operator
        nsrc = 3
        rngs = np.random.default_rng(5)
        bgs = [o.drand(2 * o.Vh * 24, seed=21 + k) for k in range(nsrc)]
        bls = [dist_util.slice_field(bgk, idx, 24) for bgk in bgs]
        xls = [np.zeros_like(bl) for _ in range(nsrc)]
        pm = inv_param()
        pm.inv_type_precondition = q.QUDA_MG_INVERTER
        pm.preconditioner = mg
        pm.num_src = nsrc
        pm.verbosity = q.QUDA_SUMMARIZE
        if pc:
            pm.solve_type = q.QUDA_DIRECT_PC_SOLVE
        L.invertMultiSrcQuda((C.c_void_p * nsrc)(*[a.ctypes.data for a in xls]), (C.c_void_p * nsrc)(*[a.ctypes.data for a in bls]), C.byref(pm))
        for k in range(nsrc):
            if world > 1:
                xs = [torch.zeros(xls[k].size, dtype=torch.float64, device="cuda") for _ in range(world)]
                dist.all_gather(xs, torch.from_numpy(xls[k]).cuda())
                xg = np.zeros(2 * o.Vh * 24).reshape(-1, 24)
                for r in range(world):
                    ridx, _ = dist_util.local_to_global_index(Xl, grid, dist_util.rank_coords(r, grid))
                    xg[ridx] = xs[r].cpu().numpy().reshape(-1, 24)
                xg = xg.ravel()
            else:
                xg = np.zeros(2 * o.Vh * 24).reshape(-1, 24); xg[idx] = xls[k].reshape(-1, 24); xg = xg.ravel()
            resk = np.linalg.norm(bgs[k] - o.tm_mat(g, xg, kappa, mu, 1, 0)) / np.linalg.norm(bgs[k])
            assert resk < 5e-8, (k, resk)
        if rank == 0:
            print(f"MULTIGPU_MG_MULTISRC_OK sources={nsrc} lockstep_iters={pm.iter} worst_true_res={pm.true_res:.2e}", flush=True)
    if rank == 0:
        print(f"MULTIGPU_MG_OK ranks={world} grid={grid} local={Xl} pc={int(pc)} peer_reduce={os.environ.get('QB_PEER_REDUCE', '1')} mg_iters={p.iter} plain_iters={p0.iter} host_res={res:.2e} true_res={p.true_res:.2e}", flush=True)
    L.destroyMultigridQuda(mg)
    L.endQuda()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
