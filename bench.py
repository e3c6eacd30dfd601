#!/usr/bin/env python
"""Benchmark of the hot path: even-odd twisted-mass Dslash (BASELINE.json metric, config[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one application of the even-odd preconditioned twisted-mass Dslash (A^-1 D, parity hop)
on the whole local lattice.  N=1 workload: 32^3 x 64, fp32, reconstruct-12 (BASELINE configs[1]); the
half-precision variant of the same config is reported in `extra`.  N>1: weak scaling, every rank
owns a 32^3 x 64 block of a lattice partitioned along T (one process per GPU, NCCL halo exchange).

`value`   : GFLOP/s (1368 flop/site, the reference's own flop model, lib/dslash_quda.cuh:496-536 +
            dslash_twisted_mass.cu:145) with fields resident in HBM, CUDA events on the library's stream.
`e2e`     : the same metric through dslashQuda() with pinned HOST buffers (H2D + D2H inside the timing).
`roofline`: compulsory bytes (576 B/site fp32 r12: 8 links x 48 B + in + out spinor) / kernel time vs
            the measured HBM peak (MEASURED_PEAKS.json).
`cpu_baseline`: the CPU oracle (OpenMP port of the reference's verify path) on the host cores.

N>1 additionally: `parity` = dslashQuda / MatQuda on a small partitioned lattice on all ranks against the
global CPU oracle before anything is timed (non-zero exit on mismatch); `extra.c3` = BASELINE configs[2], global 64^3x128
split T-only and T x Z (strong scaling, halo bytes and the NVLink rate of the halo path alone); `extra.c5_mg_gcr` (N=8) =
BASELINE configs[4], 3-level MG-GCR on global 64^3x128.  N=1 additionally: `extra.c4_48x96` = BASELINE configs[3].
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_SITE = 1368
BYTES_COMPULSORY = {(4, 12): 576, (2, 12): 296, (8, 12): 1152, (4, 18): 768, (4, 8): 448, (8, 18): 1536}
BYTES_REFMODEL = {(4, 12): 1152, (2, 12): 612, (8, 12): 2304}
LOCAL_X = (32, 32, 32, 64)
KAPPA, MU = 0.1, 0.01


def vp(a):
    return a.ctypes.data_as(C.c_void_p)


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device = device
        self.rows = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.splitlines()[0].split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit())
        reasons = []
        for idx, name in ((4, "hw_slowdown"), (5, "hw_thermal_slowdown"), (6, "sw_thermal_slowdown"), (7, "sw_power_cap")):
            if any(len(r) > idx and r[idx].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][2]) if self.rows[0][2].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def make_inputs(oracle, X, seed):
    oracle.set_dims(X)
    g = oracle.gauge(kind=1, antiperiodic=True, seed=seed)
    sp = oracle.drand(oracle.Vh * 24, seed=seed)
    return g, sp


def tiled_gauge(oracle, Xl, seed, tb=4, weak=False):
    """Random SU(3) links of a (X, Y, Z, tb) block from the oracle's generator, repeated along T to the local lattice, fp32, QDP
    even-odd order.  T is the slowest index of the checkerboard order and tb is even, so the repetition is a plain tile of each
    parity block; links are site-local, so any such field is a valid gauge field (used for the large timing-only lattices)."""
    Xb = (Xl[0], Xl[1], Xl[2], min(tb, Xl[3]))
    assert Xl[3] % Xb[3] == 0
    oracle.set_dims(Xb)
    g = oracle.weak_gauge(eps=0.25, antiperiodic=False, seed=seed) if weak else oracle.gauge(kind=1, antiperiodic=False, seed=seed)
    reps = Xl[3] // Xb[3]
    return [np.ascontiguousarray(np.tile(a.astype(np.float32).reshape(2, -1), (1, reps))).ravel() for a in g]


def parity_check(q, L, oracle, du, rank, world, grid, Xl=(8, 4, 6, 8)):
    """Multi-rank parity before timing: every rank loads its block of ONE global lattice (generated identically on every rank by the
    oracle), applies dslashQuda / MatQuda through the C ABI with the NCCL halo exchange, and compares with its slice of the global
    CPU result.  Returns {name: worst relative L2 over ranks}; raises on a tolerance violation (north_star: 1e-13 / 1e-6 / 1e-3)."""
    import torch
    import torch.distributed as dist
    L.initCommsGridQuda(4, (C.c_int * 4)(*grid), None, None)
    coords = du.rank_coords(rank, grid)
    idx, Xg = du.local_to_global_index(Xl, grid, coords)
    oracle.set_dims(Xg)
    g = oracle.gauge(kind=1, antiperiodic=True, seed=137)
    sp = oracle.drand(2 * oracle.Vh * 24, seed=137)
    Vhl, Vhg = int(np.prod(Xl)) // 2, oracle.Vh
    gl = [du.slice_field(a, idx, 18) for a in g]
    spl = du.slice_field(sp, idx, 24)
    out = {}
    for prec, name, tol in ((8, "fp64", 1e-13), (4, "fp32", 1e-6), (2, "half", 1e-3)):
        gp = q.gauge_param(Xl, cuda_prec=prec, reconstruct=12)
        L.loadGaugeQuda((C.c_void_p * 4)(*[a.ctypes.data for a in gl]), C.byref(gp))
        worst = 0.0
        for flavor, parity, matpc, dag in ((1, 0, 0, 0), (-1, 1, 0, 1)):
            p = q.invert_param(cuda_prec=prec, flavor=flavor, matpc=matpc, dagger=dag)
            inp = spl[(1 - parity) * Vhl * 24:(2 - parity) * Vhl * 24].copy()
            res = np.zeros(Vhl * 24)
            L.dslashQuda(vp(res), vp(inp), C.byref(p), parity)
            gin = sp[(1 - parity) * Vhg * 24:(2 - parity) * Vhg * 24].copy()
            full = np.zeros(2 * Vhg * 24)
            full[parity * Vhg * 24:(parity + 1) * Vhg * 24] = oracle.tm_dslash(g, gin, KAPPA, MU, flavor, parity, matpc, dag)
            ref_l = du.slice_field(full, idx, 24)[parity * Vhl * 24:(parity + 1) * Vhl * 24]
            worst = max(worst, ou_rel_l2(res, ref_l))
        p = q.invert_param(cuda_prec=prec, solution_type=q.QUDA_MAT_SOLUTION)
        res = np.zeros(2 * Vhl * 24)
        L.MatQuda(vp(res), vp(spl), C.byref(p))
        worst = max(worst, ou_rel_l2(res, du.slice_field(oracle.tm_mat(g, sp, KAPPA, MU, 1, 0), idx, 24)))
        t = torch.tensor([worst], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name] = float(t.item())
        if not out[name] <= tol:
            raise SystemExit(f"bench.py: multi-GPU parity FAILED on grid {grid}: {name} rel-L2 {out[name]:.3e} > {tol:g}")
    return {"grid": list(grid), "local": list(Xl), "global": list(Xg), "rel_l2": out,
            "checked": "dslashQuda (2 flavour/parity/dagger variants) + MatQuda per precision vs the global CPU oracle, max over ranks"}


def ou_rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def run_c3(q, L, oracle, rank, world, grid, steps, warmup, max_over_ranks, barrier, global_X=(64, 64, 64, 128)):
    """BASELINE configs[2]: global 64^3x128 over the ranks of `grid` (strong scaling), fp32 recon-12, resident fields."""
    Xl = tuple(global_X[d] // grid[d] for d in range(4))
    L.initCommsGridQuda(4, (C.c_int * 4)(*grid), None, None)
    g = tiled_gauge(oracle, Xl, seed=1000 + rank)
    gp = q.gauge_param(Xl, cpu_prec=4, cuda_prec=4, reconstruct=12, t_boundary=q.QUDA_PERIODIC_T)
    L.loadGaugeQuda((C.c_void_p * 4)(*[a.ctypes.data for a in g]), C.byref(gp))
    del g
    Vh = int(np.prod(Xl)) // 2
    sp = np.random.default_rng(7 + rank).standard_normal(Vh * 24, dtype=np.float32)
    p = q.invert_param(cuda_prec=4, cpu_prec=4)
    fin = L.newSpinorQudaB200(q.QUDA_PARITY_SITE_SUBSET, 4)
    fout = L.newSpinorQudaB200(q.QUDA_PARITY_SITE_SUBSET, 4)
    L.loadSpinorQudaB200(fin, vp(sp), C.byref(p))
    L.timeDslashQudaB200(fout, fin, C.byref(p), 0, warmup, None)
    barrier()
    ms = max_over_ranks(L.timeDslashQudaB200(fout, fin, C.byref(p), 0, steps, None))
    barrier()
    sent = C.c_double(0.0)
    L.timeHaloQudaB200(fout, fin, C.byref(p), 0, 3, C.byref(sent))
    barrier()
    halo_ms = max_over_ranks(L.timeHaloQudaB200(fout, fin, C.byref(p), 0, steps, C.byref(sent)))
    barrier()
    L.freeSpinorQudaB200(fin)
    L.freeSpinorQudaB200(fout)
    peaks, _ = measured_peaks()
    sites = Vh * world
    res = {"grid": list(grid), "local": list(Xl), "global": list(global_X), "ms_per_hop": ms,
           "gflops": FLOPS_PER_SITE * sites / (ms * 1e-3) / 1e9,
           "hbm_gbs_compulsory_per_gpu": 576 * Vh / (ms * 1e-3) / 1e9,
           "roofline_frac": 576 * Vh / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
           "effective_gbs_reference_model": 1152 * sites / (ms * 1e-3) / 1e9,
           "halo_bytes_sent_per_gpu_per_hop": sent.value,
           "halo_alone_ms": halo_ms,
           "nvlink_gbs_per_gpu_each_way_halo_alone": sent.value / (halo_ms * 1e-3) / 1e9 if halo_ms > 0 else None,
           "halo_note": "halo_alone = face pack kernel + ncclSend/Recv group of every partitioned face on the halo stream with no compute to hide "
                        "behind; inside a hop it overlaps the interior kernel"}
    return res


def run_c5(q, L, oracle, du, rank, world, grid, global_X=(64, 64, 64, 128)):
    """BASELINE configs[4]: 3-level MG-GCR twisted-mass solve on global 64^3x128, physical-point-like mu, 4^4 then 2^4 aggregates,
    24 vectors per level; every rank draws its own weak-field links (site-local: any set of local fields is a valid global field once
    ghost links are exchanged), point source on rank 0."""
    Xl = tuple(global_X[d] // grid[d] for d in range(4))
    L.initCommsGridQuda(4, (C.c_int * 4)(*grid), None, None)
    kappa, mu = 0.1248, 0.001
    oracle.set_dims(Xl)
    g = oracle.weak_gauge(eps=0.25, antiperiodic=False, seed=4711 + 31 * rank)
    gp = q.gauge_param(Xl, cuda_prec=8, reconstruct=12, cuda_prec_sloppy=4, cuda_prec_precondition=4, t_boundary=q.QUDA_PERIODIC_T)
    L.loadGaugeQuda((C.c_void_p * 4)(*[a.ctypes.data for a in g]), C.byref(gp))
    del g

    def inv_param():
        p = q.invert_param(kappa=kappa, mu=mu, cuda_prec=8, solution_type=q.QUDA_MAT_SOLUTION)
        p.cuda_prec_sloppy = 4; p.cuda_prec_precondition = 4
        p.solve_type = q.QUDA_DIRECT_SOLVE; p.inv_type = q.QUDA_GCR_INVERTER
        p.gcrNkrylov = 20; p.tol = 1e-9; p.maxiter = 10000; p.reliable_delta = 1e-4
        return p

    V = int(np.prod(Xl))
    b = np.zeros(V * 24)
    if rank == 0:
        b[0:24:2] = 1.0
    x = np.zeros_like(b)

    def one(pc):
        ip = inv_param()
        mgp = q.multigrid_param(ip, n_level=3, geo_block=((4, 4, 4, 4), (2, 2, 2, 2)), n_vec=(24, 24), nu_pre=2, nu_post=2, setup_maxiter=500,
                                setup_tol=5e-6, run_verify=False, solve_type=q.QUDA_DIRECT_PC_SOLVE if pc else q.QUDA_DIRECT_SOLVE)
        # untimed first call (lazy module loading, first cudaMallocs of the GB-sized arrays, IPC mapping of the halo arenas), as in run_mg_leg
        t0 = time.perf_counter()
        L.destroyMultigridQuda(L.newMultigridQuda(C.byref(mgp)))
        setup_first = time.perf_counter() - t0
        t0 = time.perf_counter()
        mg = L.newMultigridQuda(C.byref(mgp))
        setup = time.perf_counter() - t0

        def solve_param():
            p = inv_param()
            p.inv_type_precondition = q.QUDA_MG_INVERTER
            p.preconditioner = mg
            if pc:
                p.solve_type = q.QUDA_DIRECT_PC_SOLVE
            return p

        p = solve_param()
        L.invertQuda(vp(x), vp(b), C.byref(p))  # warm-up (allocations)
        p.iter = 0
        L.invertQuda(vp(x), vp(b), C.byref(p))
        # where the time goes: a third solve with the (stream-synchronising, hence slower) section timers of the cycle
        L.mgProfileEnableQudaB200(1)
        pp = solve_param()
        L.invertQuda(vp(x), vp(b), C.byref(pp))
        L.mgProfileEnableQudaB200(0)
        prof = {"solve_seconds_with_timers": pp.secs}
        for lvl in range(3):
            t6 = (C.c_double * 6)(); nc = C.c_long(0)
            L.mgProfileGetQudaB200(mg, lvl, t6, C.byref(nc))
            prof[f"level{lvl}"] = {"cycles": nc.value, "pre_smooth_or_coarsest_solve": t6[0], "residual": t6[1], "restrict": t6[2],
                                   "coarse_solve_incl_lower_levels": t6[3], "prolong": t6[4], "post_smooth": t6[5]}
        multi = None
        if os.environ.get("QB_BENCH_MULTI_SRC", "1") != "0":
            # the "multi-RHS coarse grid" of config 5: 12 spin-colour point sources through invertMultiSrcQuda; block path = lock-step
            # GCR, coarse levels on the multi-RHS tensor-core operator with ghost zones of block fields, against one source at a time
            nsrc = 12
            bs = []
            for k in range(nsrc):
                bk = np.zeros(V * 24)
                if rank == 0:
                    bk[2 * k] = 1.0
                bs.append(bk)
            xs = [np.zeros(V * 24) for _ in range(nsrc)]
            multi = {"sources": nsrc}
            for name, env in (("block", "1"), ("sequential", "0")):
                os.environ["QB_BLOCK_MG"] = env
                pm = solve_param()
                pm.num_src = nsrc
                ptr_x, ptr_b = (C.c_void_p * nsrc)(*[a.ctypes.data for a in xs]), (C.c_void_p * nsrc)(*[a.ctypes.data for a in bs])
                if name == "block":
                    L.invertMultiSrcQuda(ptr_x, ptr_b, C.byref(pm))  # warm-up (allocations)
                L.invertMultiSrcQuda(ptr_x, ptr_b, C.byref(pm))
                multi[name] = {"solve_seconds": pm.secs, "seconds_per_source": pm.secs / nsrc, "iterations": pm.iter, "worst_true_res": pm.true_res}
            os.environ.pop("QB_BLOCK_MG", None)
            multi["speedup_per_source"] = multi["sequential"]["solve_seconds"] / multi["block"]["solve_seconds"]
            del bs, xs
        L.destroyMultigridQuda(mg)
        out = {"setup_seconds": setup, "setup_seconds_first_call": setup_first, "solve_seconds": p.secs, "iterations": p.iter, "true_res": p.true_res, "profile": prof}
        if multi:
            out["multi_src_12_point_sources"] = multi
        return out

    res = {"peer_mailbox_allreduce_active": bool(L.commPeerReduceActiveQudaB200()), "grid": list(grid), "local": list(Xl), "global": list(global_X), "levels": 3,
           "blocks": [[4, 4, 4, 4], [2, 2, 2, 2]], "n_vec": [24, 24], "kappa": kappa, "mu": mu, "tol": 1e-9,
           "true_res_note": "relative L2 residual recomputed by the library with the fp64 operator whose multi-GPU parity is checked above"}
    res.update(one(False))
    res["even_odd"] = one(True)   # QUDA_DIRECT_PC_SOLVE outer solve, hierarchy coarsened on the even-odd system (the reference's default)
    return res


def run_c4(q, L, oracle, X=(48, 48, 48, 96)):
    """BASELINE configs[3]: 2-level MG on 48^3x96, 4^4 aggregates, 24 vectors -> 12^3x24 coarse lattice (41 472 sites, N = 48):
    coarse Dslash, prolongator and restrictor against the HBM roofline, and one MG-GCR solve.  Links: weak-field block tiled along T."""
    kappa, mu = 0.1248, 0.004
    g = tiled_gauge(oracle, X, seed=4711, weak=True)
    gp = q.gauge_param(X, cpu_prec=4, cuda_prec=8, reconstruct=12, cuda_prec_sloppy=4, cuda_prec_precondition=4, t_boundary=q.QUDA_PERIODIC_T)
    L.loadGaugeQuda((C.c_void_p * 4)(*[a.ctypes.data for a in g]), C.byref(gp))
    del g

    def inv_param():
        p = q.invert_param(kappa=kappa, mu=mu, cuda_prec=8, solution_type=q.QUDA_MAT_SOLUTION)
        p.cuda_prec_sloppy = 4; p.cuda_prec_precondition = 4
        p.solve_type = q.QUDA_DIRECT_SOLVE; p.inv_type = q.QUDA_GCR_INVERTER
        p.gcrNkrylov = 20; p.tol = 1e-9; p.maxiter = 5000; p.reliable_delta = 1e-4
        return p

    ip = inv_param()
    mgp = q.multigrid_param(ip, n_level=2, geo_block=((4, 4, 4, 4),), n_vec=(24,), nu_pre=2, nu_post=2, setup_maxiter=100, setup_tol=5e-6, run_verify=False)
    t0 = time.perf_counter()
    mg = L.newMultigridQuda(C.byref(mgp))
    setup = time.perf_counter() - t0
    V = int(np.prod(X))
    b = np.zeros(V * 24); b[0:24:2] = 1.0
    x = np.zeros_like(b)
    p = inv_param()
    p.inv_type_precondition = q.QUDA_MG_INVERTER
    p.preconditioner = mg
    L.invertQuda(vp(x), vp(b), C.byref(p))
    peaks, _ = measured_peaks()
    info = (C.c_int * 8)()
    L.mgLevelInfoQudaB200(mg, 0, info)
    sites, N = int(np.prod(info[0:4])), info[7]
    res = {"lattice": list(X), "coarse_lattice": list(info[0:4]), "n_vec": 24, "N": N, "kappa": kappa, "mu": mu, "setup_seconds": setup, "setup_maxiter": 100,
           "mg_gcr_2level": {"solve_seconds": p.secs, "iterations": p.iter, "true_res": p.true_res, "tol": 1e-9}}
    ms = L.mgTimeQudaB200(mg, 1, 0, 50)
    byts = sites * (9 * N * N * 8 + 10 * N * 8)
    res["coarse_dslash"] = {"sites": sites, "ms": ms, "gflops": sites * ((8 + 1) * 8 * N * N - 2 * N) / ms / 1e6, "hbm_gbs": byts / ms / 1e6,
                            "roofline_frac": byts / ms / 1e6 / peaks["hbm_gbs"], "bytes": byts}
    for what, name in ((2, "prolong"), (3, "restrict")):
        ms = L.mgTimeQudaB200(mg, 0, what, 20)
        byts = V * (8 * 12 * 24 + 96)
        res[name] = {"ms": ms, "hbm_gbs": byts / ms / 1e6, "roofline_frac": byts / ms / 1e6 / peaks["hbm_gbs"], "bytes": byts}
    L.destroyMultigridQuda(mg)
    res["cpu_reference_coarse_dslash"] = cpu_baseline_coarse(oracle)
    return res


def cpu_baseline_coarse(oracle):
    """CPU baseline of the coarse Dslash: the reference's OWN host path -- ApplyCoarse -> CPU coarseDslash,
    lib/dslash_coarse.cu:263-290, single-threaded as shipped (its OpenMP pragma is commented out, :279) -- from oracle/_ref/libmgref.so,
    on a bounded sample: N = 48 (n_vec 24) links built by the reference's calculateY from random vectors on 8^3x16 with 2^4 aggregates
    -> 4^3x8 = 512 coarse sites (the per-site cost does not depend on the lattice size).  Checker / baseline only, never the product."""
    from tests import oracle_util as ou
    ref = ou.load_mgref()
    if ref is None:
        return None
    Xs, kappa = (8, 8, 8, 16), 0.1248
    oracle.set_dims(Xs)
    g = oracle.weak_gauge(eps=0.25, antiperiodic=False, seed=11)
    rng = np.random.default_rng(7)
    B = [rng.standard_normal(oracle.V * 24).astype(np.float32) for _ in range(24)]
    t0 = time.perf_counter()
    T = ref.transfer(B, Xs, 4, 3, (2, 2, 2, 2), 2)
    co = ref.coarse_op(T, g, kappa, 2 * kappa * 0.004, "QUDA_TWISTED_MASS_DIRAC")
    build_s = time.perf_counter() - t0
    v = rng.standard_normal(co.V * co.N * 2).astype(np.float32)
    co.apply(v, kappa)
    reps, t0 = 0, time.perf_counter()
    while reps < 3 or time.perf_counter() - t0 < 3.0:
        co.apply(v, kappa)
        reps += 1
    sec = (time.perf_counter() - t0) / reps
    N, sites = co.N, co.V
    out = {"kind": "reference", "cores": 1, "sites": sites, "N": N, "seconds_per_application": sec,
           "gflops": sites * ((8 + 1) * 8 * N * N - 2 * N) / sec / 1e9, "us_per_site": sec / sites * 1e6,
           "links_build_seconds_reference_calculateY": build_s, "fine_sites_of_the_build": int(np.prod(Xs)),
           "sample": "lib/dslash_coarse.cu CPU coarseDslash from oracle/_ref/libmgref.so, 512 coarse sites, N = 48, 1 thread"}
    co.free(); T.free()
    return out


def run_mg_leg(q, L, oracle, X, precond=2, half_storage=False, full=True, pc=False, multi_src=None, sloppy=4):
    """Second half of the BASELINE metric: 3-level MG-GCR twisted-mass solve (seconds), plus the coarse-operator
    kernels against their HBM roofline.  32^3x64, 4^4 then 2^4 aggregates, 24 vectors per level, MR(2,2) smoother,
    K-cycle, fp64 outer GCR(20) / fp32 MG / int16 level-0 smoother, weak-field SU(3) gauge (periodic)."""
    kappa, mu = 0.1248, 0.004
    oracle.set_dims(X)
    g = oracle.weak_gauge(eps=0.25, antiperiodic=False, seed=4711)
    # sloppy = 2: int16 + norm Krylov vectors and int16 links in the outer GCR (the multigrid keeps fp32 vectors: its invert_param says so)
    gp = q.gauge_param(X, cuda_prec=8, reconstruct=12, cuda_prec_sloppy=4, cuda_prec_precondition=precond, t_boundary=q.QUDA_PERIODIC_T)
    if sloppy == 2:
        gp.cuda_prec_sloppy = 2
        gp.cuda_prec_precondition = 4 if precond == 4 else 2
    L.loadGaugeQuda((C.c_void_p * 4)(*[a.ctypes.data for a in g]), C.byref(gp))

    def inv_param():
        p = q.invert_param(kappa=kappa, mu=mu, cuda_prec=8, solution_type=q.QUDA_MAT_SOLUTION)
        p.cuda_prec_sloppy = 4
        p.cuda_prec_precondition = precond
        p.solve_type = q.QUDA_DIRECT_SOLVE
        p.inv_type = q.QUDA_GCR_INVERTER
        p.gcrNkrylov = 20
        p.tol = 1e-9
        p.maxiter = 5000
        p.reliable_delta = 1e-4
        p.verbosity = int(os.environ.get("QB_VERBOSITY", q.QUDA_SILENT))
        return p

    # pc: the reference test's default (tests/test_util.cpp:1600, multigrid_invert_test.cpp:252): outer GCR on the even-odd system,
    # every level injects single-parity fields into its coarse grid (preconditioned coarsening); the MG's own invert_param stays
    # QUDA_DIRECT_SOLVE as interface_quda.cpp:2179 demands
    ip = inv_param()
    mgp = q.multigrid_param(ip, n_level=3, geo_block=((4, 4, 4, 4), (2, 2, 2, 2)), n_vec=(24, 24), nu_pre=2, nu_post=2,
                            setup_maxiter=500, setup_tol=5e-6, run_verify=False, solve_type=q.QUDA_DIRECT_PC_SOLVE if pc else q.QUDA_DIRECT_SOLVE)
    # storage precision of the preconditioner's data (V of the transfer operators, coarse links of the single-RHS kernel): fp32, or
    # fp16 with fp32 arithmetic (what cuda_prec_precondition = half selects; here chosen independently of the level-0 smoother)
    half_storage = half_storage or os.environ.get("QB_BENCH_HALF_STORAGE") == "1"
    os.environ["QB_MG_HALF_STORAGE"] = "1" if half_storage else "0"
    setup_first_s = None
    if True:  # untimed first call of every leg (lazy module loading of the kernels this hierarchy uses, first cudaMallocs of the GB-sized transfer / link arrays)
        t0 = time.perf_counter()
        L.destroyMultigridQuda(L.newMultigridQuda(C.byref(mgp)))
        setup_first_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    mg = L.newMultigridQuda(C.byref(mgp))
    setup_s = time.perf_counter() - t0
    os.environ.pop("QB_MG_HALF_STORAGE", None)
    V = oracle.V
    b = np.zeros(V * 24)
    b[0:24:2] = 1.0  # point source on the first site (multigrid_invert_test.cpp:497-508)
    x = np.zeros_like(b)
    p = inv_param()
    p.inv_type_precondition = q.QUDA_MG_INVERTER
    p.preconditioner = mg
    p.cuda_prec_sloppy = sloppy
    if sloppy == 2:
        p.reliable_delta = 1e-2
    if pc:
        p.solve_type = q.QUDA_DIRECT_PC_SOLVE
    L.invertQuda(vp(x), vp(b), C.byref(p))  # warm-up (allocations)
    p.iter = 0
    t0 = time.perf_counter()
    L.invertQuda(vp(x), vp(b), C.byref(p))
    wall_s = time.perf_counter() - t0
    p0 = inv_param()
    if pc:
        p0.solve_type = q.QUDA_DIRECT_PC_SOLVE
    x0 = np.zeros_like(b)
    L.invertQuda(vp(x0), vp(b), C.byref(p0))
    # 12 spin-colour point sources of one propagator: invertMultiSrcQuda on the block path (all sources through the K-cycle in
    # lock-step, coarse levels on the multi-RHS tensor-core operator) against the same call with the block path switched off
    multi = None
    if multi_src is None:
        multi_src = full
    if multi_src and os.environ.get("QB_BENCH_MULTI_SRC", "1") != "0":
        nsrc = int(os.environ.get("QB_BENCH_NSRC", "12"))
        bs = []
        for k in range(nsrc):
            bk = np.zeros(V * 24); bk[2 * k] = 1.0
            bs.append(bk)
        xs = [np.zeros(V * 24) for _ in range(nsrc)]
        multi = {"sources": nsrc}
        for name, env in (("block", "1"), ("sequential", "0")):
            if name == "sequential" and os.environ.get("QB_BENCH_BLOCK_ONLY") == "1":
                continue
            os.environ["QB_BLOCK_MG"] = env
            pm = inv_param()
            pm.inv_type_precondition = q.QUDA_MG_INVERTER
            pm.preconditioner = mg
            pm.num_src = nsrc
            if pc:
                pm.solve_type = q.QUDA_DIRECT_PC_SOLVE
            ptr_x, ptr_b = (C.c_void_p * nsrc)(*[a.ctypes.data for a in xs]), (C.c_void_p * nsrc)(*[a.ctypes.data for a in bs])
            if name == "block":
                L.invertMultiSrcQuda(ptr_x, ptr_b, C.byref(pm))  # warm-up (allocations), as for the single solve above
            prof = name == "block" and os.environ.get("QB_BENCH_CUDA_PROFILER") == "1"   # ncu --profile-from-start off: this solve only
            if prof:
                import torch
                torch.cuda.cudart().cudaProfilerStart()
            t0 = time.perf_counter()
            L.invertMultiSrcQuda(ptr_x, ptr_b, C.byref(pm))
            wall = time.perf_counter() - t0
            if prof:
                torch.cuda.cudart().cudaProfilerStop()
            multi[name] = {"solve_seconds": pm.secs, "seconds_per_source": pm.secs / nsrc, "wall_seconds_incl_h2d_d2h": wall, "iterations": pm.iter, "worst_true_res": pm.true_res}
        os.environ.pop("QB_BLOCK_MG", None)
        if "sequential" in multi:
            multi["speedup_per_source"] = multi["sequential"]["solve_seconds"] / multi["block"]["solve_seconds"]
        del bs, xs
    res = {"lattice": list(X), "levels": 3, "outer_krylov_vectors": {4: "fp32", 2: "int16 + norm (cuda_prec_sloppy = half), reliable_delta 1e-2"}[sloppy], "outer_solve": "QUDA_DIRECT_PC_SOLVE, coarse_grid_solution_type MATPC on every level" if pc else "QUDA_DIRECT_SOLVE, coarse_grid_solution_type MAT",
           "smoother_precision_level0": {2: "int16", 4: "fp32"}[precond],
           "preconditioner_storage": "fp16 V and coarse links, fp32 arithmetic" if half_storage else "fp32", "blocks": [[4, 4, 4, 4], [2, 2, 2, 2]], "n_vec": [24, 24], "kappa": kappa, "mu": mu,
           "setup_seconds": setup_s, "setup_seconds_first_call": setup_first_s, "solve_seconds": p.secs, "solve_wall_seconds_incl_h2d_d2h": wall_s, "iterations": p.iter,
           "true_res": p.true_res, "tol": 1e-9, "plain_gcr_seconds": p0.secs, "plain_gcr_iterations": p0.iter, "plain_gcr_true_res": p0.true_res}
    if multi:
        res["multi_src_12_point_sources"] = multi
    if not full:
        L.destroyMultigridQuda(mg)
        return res
    peaks, _ = measured_peaks()
    for lvl in (1, 2):
        info = (C.c_int * 8)()
        L.mgLevelInfoQudaB200(mg, lvl - 1, info)
        sites = int(np.prod(info[0:4]))
        N = info[7]
        ms = L.mgTimeQudaB200(mg, lvl, 0, 50)
        byts = sites * (9 * N * N * 8 + 10 * N * 8)
        flops = sites * ((8 + 1) * 8 * N * N - 2 * N)
        res[f"coarse_dslash_level{lvl}"] = {"sites": sites, "N": N, "ms": ms, "gflops": flops / ms / 1e6, "hbm_gbs": byts / ms / 1e6,
                                            "roofline_frac": byts / ms / 1e6 / peaks["hbm_gbs"], "bytes": byts}
    # multi-RHS coarse Dslash of level 1 on the tensor cores (tcgen05 tf32, csrc/coarse_mrhs.cu): the links of a site are read
    # once for all right-hand sides; compulsory bytes = links + every vector read once and written once
    info = (C.c_int * 8)()
    L.mgLevelInfoQudaB200(mg, 0, info)
    sites, N = int(np.prod(info[0:4])), info[7]
    single_ms = res["coarse_dslash_level1"]["ms"]
    mrhs = {}
    for mode, name in ((1, "tf32"), (3, "split_tf32_fp32_accurate")):
        for nrhs in (12, 16, 32):
            if nrhs > L.mgMrhsMaxRhsQudaB200(mg, 1, mode):
                continue
            ms = L.mgTimeMrhsQudaB200(mg, 1, 0, nrhs, mode, 50)
            byts = sites * (9 * N * N * 8 + 2 * nrhs * N * 8)
            flops = sites * nrhs * ((8 + 1) * 8 * N * N - 2 * N)
            mrhs[f"{name}_nrhs{nrhs}"] = {"ms": ms, "ms_per_rhs": ms / nrhs, "speedup_per_rhs_vs_single": single_ms * nrhs / ms, "tflops": flops / ms / 1e9,
                                          "hbm_gbs": byts / ms / 1e6, "roofline_frac": byts / ms / 1e6 / peaks["hbm_gbs"]}
    res["coarse_dslash_level1_multi_rhs_tensor_core"] = mrhs
    for what, name in ((2, "prolong"), (3, "restrict")):
        ms = L.mgTimeQudaB200(mg, 0, what, 50)
        byts = V * (8 * 12 * 24 + 96)
        res[f"{name}_level0"] = {"ms": ms, "hbm_gbs": byts / ms / 1e6, "roofline_frac": byts / ms / 1e6 / peaks["hbm_gbs"]}
    L.destroyMultigridQuda(mg)
    return res


def cpu_baseline_port(oracle, g, sp, budget_s=12.0):
    """Oracle (OpenMP port) on the full 32^3x64 workload in fp32, bounded to ~budget_s seconds."""
    cores = os.cpu_count() or 1
    gf = [a.astype(np.float32) for a in g]
    ef = sp.astype(np.float32)
    oracle.set_dims(LOCAL_X)
    t0 = time.perf_counter()
    oracle.tm_dslash(gf, ef, KAPPA, MU, 1, 0, 0, 0)
    first = time.perf_counter() - t0
    reps = max(1, min(10, int(budget_s / max(first, 1e-3)) - 1))
    best = first
    for _ in range(reps):
        t0 = time.perf_counter()
        oracle.tm_dslash(gf, ef, KAPPA, MU, 1, 0, 0, 0)
        best = min(best, time.perf_counter() - t0)
    return {"value": FLOPS_PER_SITE * oracle.Vh / best / 1e9, "unit": "GFLOP/s", "cores": cores, "kind": "port",
            "sample": f"{reps + 1} x fp32 tm_dslash on the full 32^3x64 lattice (oracle/tm_oracle.c, OpenMP over sites), best of",
            "ms_per_step": best * 1e3}


_REF = {}


def _ref_init(seed_base):
    """Pool initializer: one process = one copy of the reference's own single-threaded CPU code with its own 32^3x64 fp32 fields
    (the same lattice, precision and parameters as the GPU arm's step)."""
    import multiprocessing as mp
    from tests import oracle_util as ou
    ident = mp.current_process()._identity
    seed = seed_base + (ident[0] if ident else 0)
    ref = ou.load_ref()
    orc = ou.load_oracle()
    ref.setup(LOCAL_X, antiperiodic=True)
    orc.set_dims(LOCAL_X)
    _REF["ref"] = ref
    _REF["g"] = [a.astype(np.float32) for a in orc.gauge(kind=1, antiperiodic=True, seed=seed)]
    _REF["sp"] = orc.drand(ref.Vh * 24, seed=seed).astype(np.float32)


def _ref_worker(reps):
    ref = _REF["ref"]
    t0 = time.perf_counter()
    for _ in range(reps):
        ref.tm_dslash(_REF["g"], _REF["sp"], KAPPA, MU, 1, 0, 0, 0)
    return (time.perf_counter() - t0) / reps, ref.Vh


def workload_config(prec=4, recon=12, world=1):
    """`config` of the bench line: the same dict for both arms."""
    dtype = {8: "f64", 4: "f32", 2: "i16-storage/f32-math"}[prec]
    X = LOCAL_X
    return {"workload": f"twisted-mass even-odd Dslash {X[0]}^3x{X[3]} per GPU, {dtype}, reconstruct-{recon} (BASELINE configs[1])",
            "global_lattice": [X[0], X[1], X[2], X[3] * world], "partition": [1, 1, 1, world], "kappa": KAPPA, "mu": MU,
            "l2_policy": "inputs larger than L2: gauge 403 MB + spinors 201 MB per hop vs 126 MB L2",
            "flops_per_site": FLOPS_PER_SITE, "bytes_per_site_compulsory": BYTES_COMPULSORY[(prec, recon)]}


def run_reference(args):
    """--impl reference: the reference's own CPU Dslash (oracle/_ref, unmodified sources, tests/wilson_dslash_reference.cpp) on all
    host cores, on the GPU arm's config: 32^3x64, fp32, kappa 0.1, mu 0.01, parity 0.  The code is single-threaded as shipped, so one
    independent copy with its own fields runs per core (a site-parallel stencil has no cross-core dependency) and one step = every
    copy applying tm_dslash once; the aggregate rate is reported."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from tests import oracle_util as ou
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    have_ref = ou.load_ref() is not None
    steps, warmup = args.steps, args.warmup
    rates = []
    vh = LOCAL_X[0] * LOCAL_X[1] * LOCAL_X[2] * LOCAL_X[3] // 2
    if have_ref:
        with mp.get_context("spawn").Pool(cores, initializer=_ref_init, initargs=(137,)) as pool:
            t_start = time.perf_counter()
            for s in range(warmup + steps):
                res = pool.map(_ref_worker, [1] * cores, chunksize=1)
                if s >= warmup:
                    rates.append(sum(FLOPS_PER_SITE * v / t for t, v in res) / 1e9)  # all copies run concurrently
                if time.perf_counter() - t_start > 150 and len(rates) >= 3:
                    break   # bounded sample: keep the arm within a few minutes on slow hosts
        gflops = float(np.mean(rates))
        kind = "reference"
        sample = (f"{cores} concurrent copies of the reference's tm_dslash (tests/wilson_dslash_reference.cpp, -O3, fp32), each on its own "
                  f"32^3x64 lattice, {len(rates)} timed steps of one hop per copy")
        ms = FLOPS_PER_SITE * vh * cores / gflops / 1e6
    else:
        orc = ou.load_oracle()
        g, sp = make_inputs(orc, LOCAL_X, 137)
        b = cpu_baseline_port(orc, g, sp, budget_s=20.0)
        gflops, kind, sample, ms = b["value"], "port", b["sample"], b["ms_per_step"]
    line = {"impl": "reference", "metric": "tm_dslash_gflops", "value": gflops, "unit": "GFLOP/s", "n_gpus": args.gpus,
            "steps": len(rates) if rates else steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(4, 12, args.gpus),
            "cpu_baseline": {"value": gflops, "unit": "GFLOP/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": gflops, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--prec", type=int, default=4, choices=[2, 4, 8])
    ap.add_argument("--recon", type=int, default=12, choices=[8, 12, 18])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the half-precision extra measurement")
    ap.add_argument("--no-mg", action="store_true", help="skip the 3-level MG-GCR solve leg (N=1 only)")
    ap.add_argument("--mg-precond", type=int, default=4, choices=[2, 4], help="precision of the level-0 smoother operator (2 = int16, 4 = fp32)")
    ap.add_argument("--no-parity", action="store_true", help="skip the multi-GPU parity check (N>1)")
    ap.add_argument("--no-c3", action="store_true", help="skip the 64^3x128 strong-scaling legs (N>1)")
    ap.add_argument("--no-c4", action="store_true", help="skip the 48^3x96 coarse-operator legs (N=1)")
    ap.add_argument("--c5", choices=["auto", "on", "off"], default="auto", help="3-level MG-GCR on global 64^3x128 (auto: N=8 only)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if args.gpus != 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    import importlib
    import torch
    import torch.distributed as dist
    import quda_b200 as q
    from tests import oracle_util as ou
    du = importlib.import_module("quda-qkxtm-multigrid_b200.dist")

    torch.cuda.set_device(local_rank)
    L = q.lib()
    L.initQudaDevice(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_char * 128)()
            L.ncclUniqueIdQudaB200(buf)
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        L.commsBootstrapQudaB200(rank, world, bytes(uid.cpu().numpy().tobytes()))
        L.initCommsGridQuda(4, (C.c_int * 4)(1, 1, 1, world), None, None)
    L.initQudaMemory()

    oracle = ou.load_oracle()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- multi-GPU parity gate: nothing is timed unless the partitioned operator reproduces the global CPU oracle ----
    parity_res = None
    if world > 1 and not args.no_parity:
        grids = [(1, 1, 1, world)]
        if world >= 4:
            grids.append(du.default_grid(world) if world == 8 else (1, 1, 2, world // 2))
        else:
            grids.append((1, 1, world, 1))
        parity_res = [parity_check(q, L, oracle, du, rank, world, gr) for gr in grids]
        L.initCommsGridQuda(4, (C.c_int * 4)(1, 1, 1, world), None, None)

    X = LOCAL_X
    g, sp = make_inputs(oracle, X, 137 + 17 * rank)
    Vh = oracle.Vh

    def run_config(prec, recon, steps, warmup, with_e2e):
        gp = q.gauge_param(X, cuda_prec=prec, reconstruct=recon)
        L.loadGaugeQuda((C.c_void_p * 4)(*[a.ctypes.data for a in g]), C.byref(gp))
        p = q.invert_param(cuda_prec=prec)
        fin = L.newSpinorQudaB200(q.QUDA_PARITY_SITE_SUBSET, prec)
        fout = L.newSpinorQudaB200(q.QUDA_PARITY_SITE_SUBSET, prec)
        L.loadSpinorQudaB200(fin, vp(sp), C.byref(p))
        L.timeDslashQudaB200(fout, fin, C.byref(p), 0, warmup, None)
        per = (C.c_float * steps)()
        n0 = L.kernelLaunchCountQudaB200()
        barrier()
        t0 = time.perf_counter()
        ms = L.timeDslashQudaB200(fout, fin, C.byref(p), 0, steps, per)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3 / steps
        launches = L.kernelLaunchCountQudaB200() - n0
        res = {"ms": ms, "wall_ms": wall_ms, "launches": launches, "per": list(per)}
        if with_e2e:
            # e2e: dslashQuda with pinned host buffers; H2D of the input and D2H of the result inside the timed region
            hin = torch.from_numpy(sp.astype(np.float32 if prec != 8 else np.float64)).pin_memory()
            hout = torch.empty_like(hin).pin_memory()
            pe = q.invert_param(cuda_prec=prec, cpu_prec=(8 if prec == 8 else 4))
            for _ in range(3):
                L.dslashQuda(C.c_void_p(hout.data_ptr()), C.c_void_p(hin.data_ptr()), C.byref(pe), 0)
            n = max(3, min(steps, 20))
            barrier()
            t0 = time.perf_counter()
            for _ in range(n):
                L.dslashQuda(C.c_void_p(hout.data_ptr()), C.c_void_p(hin.data_ptr()), C.byref(pe), 0)
            barrier()
            res["e2e_ms"] = (time.perf_counter() - t0) * 1e3 / n
            res["e2e_bytes"] = hin.numel() * hin.element_size()
            # what the PCIe link alone allows for these two buffers: H2D and D2H of the same sizes on two streams at once (no compute, no
            # dependency between them); dslashQuda cannot beat it and has to add the three slabs whose result can only leave after the
            # last input slab has arrived
            dbuf_in = torch.empty(hin.numel(), dtype=hin.dtype, device="cuda")
            dbuf_out = torch.empty(hin.numel(), dtype=hin.dtype, device="cuda")
            s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
            torch.cuda.synchronize()
            best = None
            for _ in range(6):
                t0 = time.perf_counter()
                with torch.cuda.stream(s1):
                    dbuf_in.copy_(hin, non_blocking=True)
                with torch.cuda.stream(s2):
                    hout.copy_(dbuf_out, non_blocking=True)
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) * 1e3
                best = dt if best is None else min(best, dt)
            res["pcie_ms"] = best
            del dbuf_in, dbuf_out
        L.freeSpinorQudaB200(fin)
        L.freeSpinorQudaB200(fout)
        return res

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    main_res = run_config(args.prec, args.recon, args.steps, args.warmup, True)
    extra = {}
    if not args.no_extra and args.prec == 4:
        half = run_config(2, 12, args.steps, args.warmup, False)
        extra["half_r12"] = half
    mg_res = None
    reduce_res = None
    if world == 1 and not args.no_mg and not args.no_extra:
        mg_res = run_mg_leg(q, L, oracle, X, args.mg_precond)
        mg_res_h16 = run_mg_leg(q, L, oracle, X, args.mg_precond, half_storage=True, full=False)
        mg_res_pc = run_mg_leg(q, L, oracle, X, args.mg_precond, full=False, pc=True, multi_src=True)
        mg_res_pc_h16 = run_mg_leg(q, L, oracle, X, args.mg_precond, half_storage=True, full=False, pc=True)
        # int16 outer Krylov space: needs the fp32 links resident for the multigrid, i.e. an fp32 level-0 smoother
        mg_res_pc_h16_hs = run_mg_leg(q, L, oracle, X, 4, half_storage=True, full=False, pc=True, sloppy=2) if args.mg_precond == 4 else None
    sampler.stop_flag = True

    # max over ranks of the device time
    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if world > 1:
        del g
    c3_res, c5_res, c4_res = None, None, None
    if world > 1 and not args.no_c3 and not args.no_extra:
        # BASELINE configs[2]: 64^3x128 over T, and over T x Z (Z alone on 2 GPUs)
        c3_res = {"t_only": run_c3(q, L, oracle, rank, world, (1, 1, 1, world), args.steps, args.warmup, max_over_ranks, barrier)}
        g2 = (1, 1, 2, world // 2)
        c3_res["z_only" if world == 2 else "t_and_z"] = run_c3(q, L, oracle, rank, world, g2, args.steps, args.warmup, max_over_ranks, barrier)
    if world > 1 and not args.no_extra:
        # global reductions: host-visible latency of one norm2 of a coarse-level sized vector (98 304 reals) with the all-reduce
        # fused into the reduction kernel over the NVLink peer mailboxes, and with ncclAllReduce on the compute stream
        barrier()
        active = L.commPeerReduceActiveQudaB200()
        us_peer = max_over_ranks(L.timeReduceQudaB200(98304, 200, 1))
        barrier()
        us_nccl = max_over_ranks(L.timeReduceQudaB200(98304, 200, 0))
        barrier()
        reduce_res = {"peer_mailbox_allreduce_active": bool(active), "norm2_us_fused_peer_allreduce": us_peer if active else None,
                      "norm2_us_nccl_allreduce": us_nccl, "vector_reals": 98304,
                      "note": "kernel launch + all-reduce over the ranks + the stream synchronisation that hands the sum to the host, mean of 200, max over ranks"}
    if world > 1 and (args.c5 == "on" or (args.c5 == "auto" and world == 8)) and not args.no_extra:
        c5_res = run_c5(q, L, oracle, du, rank, world, du.default_grid(world))
    if world == 1 and not args.no_c4 and not args.no_mg and not args.no_extra:
        c4_res = run_c4(q, L, oracle)

    ms = max_over_ranks(main_res["ms"])
    e2e_ms = max_over_ranks(main_res["e2e_ms"])
    half_ms = max_over_ranks(extra["half_r12"]["ms"]) if extra else None

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        sites = Vh * world
        gflops = FLOPS_PER_SITE * sites / (ms * 1e-3) / 1e9
        bpsite = BYTES_COMPULSORY[(args.prec, args.recon)]
        # the hop kernel is the only kernel in a step at N=1; at N>1 pack + interior + boundary launches share the step
        achieved = bpsite * Vh / (ms * 1e-3) / 1e9
        traffic = None
        prof = os.path.join(ROOT, "profiles", "dslash_traffic.json")
        if os.path.exists(prof):
            try:
                traffic = json.load(open(prof)).get(f"prec{args.prec}_recon{args.recon}")
            except Exception:
                traffic = None
        cpu = None
        if not args.no_cpu and world == 1:  # reported on rank 0 at N = 1 only (torchrun also pins OMP_NUM_THREADS=1)
            cpu = cpu_baseline_port(oracle, g, sp)
        dtype = {8: "f64", 4: "f32", 2: "i16-storage/f32-math"}[args.prec]
        line = {
            "metric": "tm_dslash_gflops", "value": gflops, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": workload_config(args.prec, args.recon, world),
            "effective_gbs_reference_model": BYTES_REFMODEL.get((args.prec, args.recon), 0) * sites / (ms * 1e-3) / 1e9,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                         "traffic": traffic, "traffic_source": "profiles/dslash_traffic.json (dram__bytes of one ncu --set full capture of this kernel; a constant, not measured in this run)",
                         "peak_source": peak_kind, "kernel": "dslash_kernel<StoreS,12,TWIST_IN=false,HAS_X=false,GHOST=false>" if args.prec == 4 else "dslash_kernel"},
            "e2e": {"value": FLOPS_PER_SITE * sites / (e2e_ms * 1e-3) / 1e9, "unit": "GFLOP/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": main_res["e2e_bytes"], "d2h_bytes_per_step": main_res["e2e_bytes"],
                    "pcie_floor_ms": main_res.get("pcie_ms"),
                    "rank0_cpu_affinity": len(os.sched_getaffinity(0)),
                    "affinity_note": "N > 1: each rank is bound to the cpus of its GPU's NUMA node (initQudaDevice / commsBootstrap, as the reference's setNumaAffinity) before the pinned host buffers are allocated",
                    "pcie_floor_note": "H2D + D2H of the same two pinned buffers on two streams at once, no compute: the link-bound lower limit of this call"},
            "gpu_launches": main_res["launches"],
            "clocks": sampler.summary(),
            "cpu_baseline": cpu,
            "extra": {},
        }
        if parity_res:
            line["parity"] = {"status": "green", "parity_rel_l2": {k: max(r["rel_l2"][k] for r in parity_res) for k in ("fp64", "fp32", "half")},
                              "tolerance": {"fp64": 1e-13, "fp32": 1e-6, "half": 1e-3}, "cases": parity_res}
        if c3_res:
            line["extra"]["c3_64x128"] = c3_res
        if reduce_res:
            line["extra"]["global_reduction"] = reduce_res
        if c5_res:
            line["extra"]["c5_mg_gcr"] = c5_res
        if c4_res:
            line["extra"]["c4_48x96"] = c4_res
        if mg_res:
            line["extra"]["mg_gcr_3level"] = mg_res
            line["extra"]["mg_gcr_3level_fp16_preconditioner_storage"] = mg_res_h16
            line["extra"]["mg_gcr_3level_even_odd"] = mg_res_pc
            line["extra"]["mg_gcr_3level_even_odd_fp16_preconditioner_storage"] = mg_res_pc_h16
            if mg_res_pc_h16_hs:
                line["extra"]["mg_gcr_3level_even_odd_fp16_storage_half_sloppy"] = mg_res_pc_h16_hs
        if half_ms:
            line["extra"]["half_r12"] = {"ms_per_step": half_ms, "gflops": FLOPS_PER_SITE * sites / (half_ms * 1e-3) / 1e9,
                                         "hbm_gbs_compulsory": 296 * Vh / (half_ms * 1e-3) / 1e9,
                                         "roofline_frac": 296 * Vh / (half_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]}
        print(json.dumps(line))
    L.endQuda()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
