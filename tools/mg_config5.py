"""BASELINE config 5 style run under torchrun: 3-level MG-GCR twisted-mass solve on a lattice partitioned over the
ranks (T first, then Z).  Every rank draws its own weak-field SU(3) links (site-local, so any set of local
fields is a valid global field once ghost links are exchanged); point source on rank 0; physical-point-like mu.
    QB_LOCAL=64,64,64,16 python -m torch.distributed.run --nproc-per-node 8 ... tools/mg_config5.py
Prints one JSON line on rank 0."""
import ctypes as C
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import quda_b200 as q  # noqa: E402
from tests import oracle_util as ou  # noqa: E402

du = importlib.import_module("quda-qkxtm-multigrid_b200.dist")


def main():
    grid = tuple(int(x) for x in os.environ["QB_GRID"].split(",")) if os.environ.get("QB_GRID") else None
    Xl = tuple(int(x) for x in os.environ.get("QB_LOCAL", "32,32,32,16").split(","))
    nvec = int(os.environ.get("QB_NVEC", "24"))
    kappa, mu = float(os.environ.get("QB_KAPPA", "0.1248")), float(os.environ.get("QB_MU", "0.001"))
    L = q.lib()
    rank, world, dist = du.init_comms(L, grid)
    grid = grid or du.default_grid(world)
    L.initQudaMemory()
    coords = du.rank_coords(rank, grid)
    o = ou.load_oracle()
    o.set_dims(Xl)
    g = o.weak_gauge(eps=0.25, antiperiodic=False, seed=4711 + 31 * rank)
    gp = q.gauge_param(Xl, cuda_prec=8, reconstruct=12, cuda_prec_sloppy=4, cuda_prec_precondition=4, t_boundary=q.QUDA_PERIODIC_T)
    L.loadGaugeQuda((C.c_void_p * 4)(*[a.ctypes.data for a in g]), C.byref(gp))
    del g

    def inv_param():
        p = q.invert_param(kappa=kappa, mu=mu, cuda_prec=8, solution_type=q.QUDA_MAT_SOLUTION)
        p.cuda_prec_sloppy = 4; p.cuda_prec_precondition = 4
        p.solve_type = q.QUDA_DIRECT_SOLVE; p.inv_type = q.QUDA_GCR_INVERTER
        p.gcrNkrylov = 20; p.tol = 1e-9; p.maxiter = 10000; p.reliable_delta = 1e-4
        return p

    ip = inv_param()
    mgp = q.multigrid_param(ip, n_level=3, geo_block=((4, 4, 4, 4), (2, 2, 2, 2)), n_vec=(nvec, nvec), nu_pre=2, nu_post=2,
                            setup_maxiter=int(os.environ.get("QB_SETUP_ITER", "500")), setup_tol=5e-6, run_verify=False)
    t0 = time.perf_counter()
    mg = L.newMultigridQuda(C.byref(mgp))
    setup = time.perf_counter() - t0
    V = int(np.prod(Xl))
    b = np.zeros(V * 24)
    if rank == 0:
        b[0:24:2] = 1.0
    x = np.zeros_like(b)
    p = inv_param()
    p.inv_type_precondition = q.QUDA_MG_INVERTER
    p.preconditioner = mg
    L.invertQuda(x.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), C.byref(p))
    out = {"ranks": world, "grid": list(grid), "local": list(Xl), "global": [Xl[d] * grid[d] for d in range(4)], "n_vec": nvec, "kappa": kappa, "mu": mu,
           "setup_seconds": setup, "mg_gcr_solve_seconds": p.secs, "mg_gcr_iterations": p.iter, "mg_gcr_true_res": p.true_res}
    if os.environ.get("QB_PLAIN", "1") == "1":
        p0 = inv_param()
        x0 = np.zeros_like(b)
        L.invertQuda(x0.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), C.byref(p0))
        out.update({"plain_gcr_seconds": p0.secs, "plain_gcr_iterations": p0.iter, "plain_gcr_true_res": p0.true_res})
    if rank == 0:
        print("CONFIG5 " + json.dumps(out), flush=True)
    L.destroyMultigridQuda(mg)
    L.endQuda()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
