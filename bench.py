#!/usr/bin/env python
"""bench.py -- surrogate-evaluated DAMH chain steps/s (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg2|cfg5]

Workload (default cfg3, BASELINE.json configs[2], the configuration the metric and its 1e9 steps/s
target are quoted on; it fits one GPU): per GPU 65536 lockstep DAMH chains, d=2, m=1, the
simple.json Gaussian problem, cubic RBF surrogate with 4096 centers, analytic toy G on the device,
surrogate_is_updated=true: every 64 lockstep steps the new snapshots are compacted, all-gathered
(NCCL), at most 256 of them enter the collector's refit (greedy thinning back to 4096 centers,
saddle-point system build, dense direct solve: null-space projection + Cholesky, `--solver direct` selects
the LU with partial pivoting that mirrors the reference's solver_type 'direct'), the new coefficients
are broadcast and every chain switches to them.  Weak scaling: every GPU carries 65536 chains.
    cfg2: 4096 chains, frozen polynomial surrogate (degree 5)        (BASELINE.json configs[1])
    cfg5: 131072 chains/GPU, d=4, m=3, frozen cubic RBF 500 centers, frozen 2000-center RBF as G
One bench "step" = one block of 64 lockstep steps of all chains (+ the refit round in cfg3).

One JSON line on stdout (rank 0).  `value` = chain steps/s over all GPUs with state resident in
HBM; `e2e` = the same through host buffers (pinned H2D of the chain samples, D2H of samples +
counters every step); `roofline` = the fused chain kernel against the measured FP64 FMA peak;
`cpu_baseline` = the numpy oracle (reference restatement) on the host cores, frozen surrogate.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

SIMPLE = dict(prior_mean=[5.0, 3.0], prior_std=[[4, -2], [-2, 4]], noise_std=2e-4, no_observations=1, observations=[-1e-3])
DARCY = dict(prior_mean=0.0, prior_std=1.0, noise_std=[0.2, 0.1, 0.1], no_observations=3,
             observations=[10.24925199, -6.07103574, -4.17821625])

WORKLOADS = {
    "cfg3": dict(d=2, m=1, problem=SIMPLE, chains=65536, surrogate="rbf", centers=4096, kernel_type=1, updated=True,
                 block=64, snapshots_per_refit=256, proposal_std=1.0, truth="toy", solver_type="nullspace",
                 name="illustrative-shape DAMH: 65536 chains/GPU, cubic RBF 4096 centers, surrogate_is_updated, refit every 64 steps"),
    "cfg2": dict(d=2, m=1, problem=SIMPLE, chains=4096, surrogate="poly", degree=5, updated=False, block=64,
                 proposal_std=1.0, truth="toy", name="simple.json-shape DAMH: 4096 chains, frozen poly degree 5"),
    "cfg5": dict(d=4, m=3, problem=DARCY, chains=131072, surrogate="rbf", centers=500, kernel_type=1, updated=False,
                 block=64, proposal_std=0.2, truth="rbf", truth_centers=2000,
                 name="Darcy-shape surrogate-only DAMH: 131072 chains/GPU, frozen cubic RBF 500 centers, 2000-center RBF as G"),
}


# ----------------------------------------------------------------------------------------------------
# synthetic inputs (host side, deterministic)
# ----------------------------------------------------------------------------------------------------
def toy_G(X):
    """Closed form of the toy model for building synthetic training sets (not on the timed path)."""
    return (-0.0375 * np.exp(-X[:, 0]) - 0.0125 * np.exp(-X[:, 1])).reshape(-1, 1)


def smooth_map(X, m):
    rs = np.random.RandomState(77)
    W = rs.standard_normal((X.shape[1], m)) * 0.6
    shift = np.array(DARCY["observations"][:m]) - np.array([0.3, -0.2, 0.1][:m])
    return shift + 2.0 * np.sin(X @ W)


def training_set(w, n, seed):
    rs = np.random.RandomState(seed)
    d = w["d"]
    if w["problem"] is SIMPLE:
        Lc = np.linalg.cholesky(np.array(SIMPLE["prior_std"], dtype=float))
        X = rs.standard_normal((n, d)) @ Lc.T + np.array(SIMPLE["prior_mean"])
        return X, toy_G(X)
    X = rs.standard_normal((n, d))
    return X, smooth_map(X, w["m"])


# ----------------------------------------------------------------------------------------------------
# CPU baseline: the numpy oracle's DAMH chain (restatement of classes_SAMPLER.py / surrogate_*.py)
# ----------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    wname, SOL, truth_SOL, seconds, seed = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    from oracle import chain as OC
    from oracle import problem as OPR
    from oracle import solvers as OS
    from oracle import surrogate_poly as OP
    from oracle import surrogate_rbf as ORB
    w = WORKLOADS[wname]
    d, m = w["d"], w["m"]
    if w["surrogate"] == "rbf":
        apply_o = ORB.RbfApply(d, m, kernel_type=w["kernel_type"])
    else:
        apply_o = OP.PolyApply(d, m)

    class Frozen:
        def send_parameters(self, p):
            self.p = p.copy()

        def recv_observations(self):
            return apply_o.apply(SOL, self.p)

        def send_to_data_collector(self, s):
            pass

    solver = OS.LinElast2Exp() if w["truth"] == "toy" else OS.SurrogateAsSolver(ORB.RbfApply(d, m, kernel_type=1), truth_SOL)
    prob = OPR.GaussProblem(d, **w["problem"])
    prop = OPR.RandomWalkProposal(d, proposal_std=w["proposal_std"], seed=seed + 1)
    ch = OC.Chain(prob, prop, OS.LocalSolverProxy(solver), kind="DAMH", seed=seed + 2, surrogate=Frozen(),
                  surrogate_is_updated=False)
    ch.prepare()
    t0 = time.perf_counter()
    steps = 0
    while time.perf_counter() - t0 < seconds:
        for _ in range(50):
            ch.step(steps)
            steps += 1
        ch.trace.clear(); ch.raw_rows.clear(); ch.data_rows.clear(); ch.damh_rows.clear()
    return steps, time.perf_counter() - t0


def cpu_baseline(wname, SOL, truth_SOL, seconds, cores=None):
    cores = cores or len(os.sched_getaffinity(0))
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [(wname, SOL, truth_SOL, seconds, 1000 * r) for r in range(cores)])
        wall = time.perf_counter() - t0
    steps = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    return dict(value=steps / busy, unit="chain steps/s", cores=cores, kind="port",
                sample="%d processes x %.0f s of oracle DAMH steps (frozen surrogate, in-process proxies, no MPI hop), %d steps"
                       % (cores, seconds, steps), wall_s=wall)


# ----------------------------------------------------------------------------------------------------
def clocks_monitor_start(path):
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,timestamp")
    try:
        return subprocess.Popen(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200"],
                                stdout=open(path, "w"), stderr=subprocess.DEVNULL)
    except OSError:
        return None


def clocks_summary(path, gpu_index, t0=None, t1=None):
    """Median SM clock and throttle reasons of the samples taken between wall-clock times t0 and t1 (the timed
    region; the monitor itself runs from before the warm-up so that it is sampling by then)."""
    import datetime
    sm, mx, reasons = [], 0.0, set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    try:
        for line in open(path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9 or f[0] != str(gpu_index):
                continue
            if t0 is not None and len(f) >= 10:
                try:
                    ts = datetime.datetime.strptime(f[9], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    if ts < t0 - 0.1 or ts > t1 + 0.1:
                        continue
                except ValueError:
                    pass
            try:
                sm.append(float(f[1])); mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
    except OSError:
        pass
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
            "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
def build_models(w, rank_seed=0):
    """Initial surrogate (GPU refit of a synthetic training set) and, for cfg5, the frozen truth."""
    from surrdamh_b200 import surrogate_poly as SP
    from surrdamh_b200 import surrogate_rbf as SR
    from surrdamh_b200.device import DeviceModel
    d, m = w["d"], w["m"]
    updater = None
    if w["surrogate"] == "rbf":
        X, Y = training_set(w, w["centers"], 1)
        updater = SR.Surrogate_update(d, m, kernel_type=w["kernel_type"], solver_type=w.get("solver_type", "direct"),
                                      no_keep=w["centers"])
        updater.add_arrays(X, Y)
        sur, _ = updater.update_device()
    else:
        X, Y = training_set(w, 2000, 1)
        updater = SP.Surrogate_update(d, m, max_degree=w["degree"])
        updater.add_arrays(X, Y)
        sur, _ = updater.update_device()
    truth = DeviceModel.toy()
    truth_SOL = None
    if w["truth"] == "rbf":
        Xt, Yt = training_set(w, w["truth_centers"], 2)
        tu = SR.Surrogate_update(d, m, kernel_type=1, solver_type="direct")
        tu.add_arrays(Xt, Yt)
        truth, _ = tu.update_device()
        truth_SOL = truth.sol_numpy()
    return updater, sur, truth, truth_SOL


def run_ours(args):
    import torch
    import torch.distributed as dist

    from surrdamh_b200 import _lib as L
    from surrdamh_b200.chains import ChainBlock, Proposal, Trace, compact_snapshots
    from surrdamh_b200.collector import Collector
    from surrdamh_b200.device import DeviceProblem

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    L.require_cuda()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if not os.environ.get("SDB_KEEP_NCCL_DEBUG"):
            os.environ["NCCL_DEBUG"] = "WARN"      # NCCL's version banner goes to stdout, which carries the JSON line
        dist.init_process_group("nccl", device_id=dev)
    w = dict(WORKLOADS[args.workload])
    if args.chains:
        w["chains"] = args.chains
    if args.centers and w["surrogate"] == "rbf":
        w["centers"] = args.centers
    if args.solver and w["surrogate"] == "rbf":
        w["solver_type"] = args.solver
    d, m, C_, K = w["d"], w["m"], w["chains"], w["block"]
    updater, sur, truth, truth_SOL = build_models(w)
    SOL0 = sur.sol_numpy()

    # ---- CPU baseline (rank 0, N == 1 only), before the timed GPU region -------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.profile:
        cpu = cpu_baseline(args.workload, SOL0, truth_SOL, args.cpu_seconds)

    prob = DeviceProblem(d, **w["problem"], device=dev)
    blk = ChainBlock(prob, C_, chain0=rank * C_, seed=2024, device=dev)
    prop = Proposal(d, w["proposal_std"], device=dev)
    col = None
    if w["updated"]:
        col = Collector(updater if (rank == 0 or world == 1) else None, "rbf", d, m, kernel_type=w["kernel_type"], rank=rank,
                        world=world, snapshots_per_refit=w["snapshots_per_refit"], refit_mode="broadcast", device=dev)
        col.model = sur
    overlap = bool(w["updated"]) and args.overlap
    traces = [Trace(K, C_, d, m, dev, snapshots=True, full=False) for _ in range(2 if overlap else 1)] if w["updated"] else [None]
    acol = None
    if overlap:
        from surrdamh_b200.collector import AsyncCollector
        acol = AsyncCollector(col, dev)
    flush = torch.empty((256 << 20) // 8, dtype=torch.float64, device=dev)      # > 126 MB of L2
    model = {"sur": sur, "swapped": False, "b": 0, "drained": 0}
    kern_ms, refit_ms = [], []
    ev_pool = []

    def one_block(timed, host_io=None):
        """One bench step.  host_io = (pinned_in, pinned_out_u, pinned_out_cnt): the e2e variant.
        With the asynchronous collector the snapshots of block b are refitted on the side stream while block b+1
        runs, and the surrogate of round b is switched in at the start of block b+2."""
        b = model["b"]
        model["b"] += 1
        main = torch.cuda.current_stream()
        if acol is not None and b >= 2:
            R, new = acol.result(b - 2)
            main.wait_event(R)
            acol.forget(b - 2)
            if new is not None:
                model["sur"], model["swapped"] = new, True
        trace = traces[b % len(traces)]
        flush.fill_(1.0)
        if host_io is not None:
            blk.u.copy_(host_io[0], non_blocking=True)                      # H2D: chain samples from pinned host memory
            blk.prepare("DAMH", model["sur"], truth)                         # G(u), posteriors of the uploaded samples
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        blk.run("DAMH", K, prop, model["sur"], truth, trace=trace, refresh_pre=model["swapped"], coresident=acol is not None)
        e1.record()
        if timed:
            ev_pool.append((e0, e1, model["swapped"]))
        model["swapped"] = False
        if acol is not None:
            acol.submit(b, trace, K, w["snapshots_per_refit"], e1, blk.step0, timed)
        elif col is not None:
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record()
            par, obs, cnt = compact_snapshots(trace, K, w["snapshots_per_refit"])
            if col.exchange_and_refit(par, obs, cnt, blk.step0):
                model["sur"], model["swapped"] = col.model, True
            r1.record()
            if timed:
                refit_ms.append((r0, r1))
        if host_io is not None:
            host_io[1].copy_(blk.u, non_blocking=True)                       # D2H: samples + counters
            host_io[2].copy_(blk.counters, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def drain():
        """Make the sampling stream wait for every collector round issued so far (their surrogates stay queued)."""
        if acol is None:
            return
        main = torch.cuda.current_stream()
        for b in range(max(model["drained"], model["b"] - 2), model["b"]):
            R, _ = acol.result(b)
            main.wait_event(R)
        model["drained"] = model["b"]

    def barrier():
        drain()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed_region(n_steps, host_io=None):
        barrier()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(n_steps):
            one_block(True, host_io)
        drain()             # the last collector rounds belong to the timed region
        t1.record()
        barrier()
        ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    clk_path = os.path.join(tempfile.gettempdir(), "sdb_clocks_%d.csv" % os.getpid())
    mon = clocks_monitor_start(clk_path) if (rank == 0 and not args.profile) else None
    blk.prepare("DAMH", sur, truth)
    for _ in range(max(args.warmup, 3)):
        one_block(False)
    if args.profile:
        # one bench step between cudaProfilerStart/Stop (ncu --profile-from-start off); nothing is reported
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.start()
        for _ in range(args.profile):
            one_block(False)
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.stop()
        print(json.dumps({"profile_blocks": args.profile, "note": "no measurement taken in profile mode"}))
        return
    # ---- FP64 peak probe (roofline denominator), live ---------------------------------------------------
    info = L.device_info()
    pb, iters = info["sm_count"] * 8, 20000
    pout = torch.empty((pb * 256,), dtype=torch.float64, device=dev)
    L.check(L.lib().sdb_fp64_probe(iters, pb, L.ptr(pout), L.stream_handle()))
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    L.check(L.lib().sdb_fp64_probe(iters, pb, L.ptr(pout), L.stream_handle()))
    p1.record()
    p1.synchronize()
    fp64_peak_tf = pb * 256 * 8 * 2.0 * iters / (p0.elapsed_time(p1) * 1e-3) / 1e12
    pout.zero_()
    L.check(L.lib().sdb_fp64_probe3(iters, pb, L.ptr(pout), L.stream_handle()))
    pout.zero_()
    p0.record()
    L.check(L.lib().sdb_fp64_probe3(iters, pb, L.ptr(pout), L.stream_handle()))
    p1.record()
    p1.synchronize()
    fp64_peak3_tf = pb * 256 * 8 * 2.0 * iters / (p0.elapsed_time(p1) * 1e-3) / 1e12

    # ---- timed: device-resident --------------------------------------------------------------------------
    L.lib().sdb_launch_count(1)
    wall0 = time.time()
    ms_total = timed_region(args.steps)
    wall1 = time.time()
    launches = int(L.lib().sdb_launch_count(1))
    if mon is not None:
        mon.terminate()
        mon.wait()
    kern = [(a.elapsed_time(b), sw) for a, b, sw in ev_pool]
    if acol is not None:
        refit_ms = list(acol.round_ms)
        acol.round_ms.clear()
    rf = [a.elapsed_time(b) for a, b in refit_ms]
    ev_pool.clear(); refit_ms.clear()

    # ---- timed: end to end through host buffers -------------------------------------------------------------
    pin_in = torch.empty((d, C_), dtype=torch.float64).pin_memory()
    pin_in.copy_(blk.u.cpu())
    pin_u = torch.empty((d, C_), dtype=torch.float64).pin_memory()
    pin_c = torch.empty((4, C_), dtype=torch.int32).pin_memory()
    io = (pin_in, pin_u, pin_c)
    one_block(False, io)
    e2e_steps = max(2, min(args.steps, 8))
    ms_e2e = timed_region(e2e_steps, io)
    ev_pool.clear(); refit_ms.clear()
    if acol is not None:
        acol.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    total_steps = float(C_) * K * args.steps * world
    value = total_steps / (ms_total * 1e-3)
    e2e_value = float(C_) * K * e2e_steps * world / (ms_e2e * 1e-3)
    # ---- roofline of the dominant kernel (fused chain kernel), algorithmic FLOPs per SURVEY.md 8(d) ---------------
    n_c = w.get("centers", 0)
    if w["surrogate"] == "rbf":
        per_point = n_c * (3 * d + 3 + 2 * m) + m * (2 * d + 1)
    else:
        P = int(sur.alpha.shape[0])
        per_point = d * w["degree"] + d * (w["degree"] + 1) ** 2 + P * (d + 2 * m)
    if w["truth"] == "rbf":
        # the true model is evaluated only after a stage-1 pass; counted at the observed rate below
        pass
    avg_ms = float(np.mean([k for k, _ in kern]))
    frac_sw = float(np.mean([1.0 if sw else 0.0 for _, sw in kern]))
    flops_launch = float(C_) * (K + frac_sw) * per_point
    achieved_tf = flops_launch / (avg_ms * 1e-3) / 1e12
    out = {
        "metric": "surrogate-evaluated DAMH chain steps/s", "value": value, "unit": "chain steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s: %s" % (args.workload, w["name"]), "chains_per_gpu": C_, "lockstep_steps_per_bench_step": K,
                   "no_parameters": d, "no_observations": m, "centers": n_c or None,
                   "l2": "256 MiB flush write before every bench step, inside the timed region",
                   "refit": ("every %d steps: compact + all-gather + thin + dense solve (%s, N=%d) + broadcast; %.2f ms avg per round; %s"
                             % (K, w.get("solver_type", "direct"), n_c + d + 1, float(np.mean(rf)),
                                "asynchronous collector: round b runs on a side stream during block b+1, its surrogate is used from block b+2"
                                if overlap else "synchronous between blocks")) if rf else "frozen surrogate"},
        "e2e": {"value": e2e_value, "unit": "chain steps/s", "h2d_bytes_per_step": int(pin_in.numel() * 8),
                "d2h_bytes_per_step": int(pin_u.numel() * 8 + pin_c.numel() * 4), "steps": e2e_steps,
                "ms_per_step": ms_e2e / e2e_steps},
        "gpu_launches": launches,
        "roofline": {"bound": "fp64-pipe", "achieved": achieved_tf, "peak": fp64_peak_tf, "unit": "TFLOP/s",
                     "frac": achieved_tf / fp64_peak_tf,
                     "traffic": 123539712 if args.workload == "cfg3" and not args.chains and not args.centers else None,
                     "traffic_note": "DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the ncu --set full capture "
                                     "in profiles/r01_chain_kernel_ncu.md; algorithmic bytes = C*K*25 B trace + state = 1.05e8",
                     "kernel": "chain_kernel<%d,%d>" % (d, m),
                     "kernel_ms": avg_ms, "kernel_share_of_step": avg_ms * len(kern) / ms_total,
                     "flops_per_launch": flops_launch, "peak_three_register_fma": fp64_peak3_tf,
                     "peak_source": "sdb_fp64_probe (8 independent DFMA chains/thread), measured live in this run; "
                                    "MEASURED_PEAKS.json has no FP64 entry"},
        "roofline_hbm": {"bound": "hbm", "frac": (float(C_) * K * (25 if w["updated"] else 0) + 2.0 * C_ * (8 * (d + m + 2) + 16))
                         / (avg_ms * 1e-3) / 1e9 / 6548.2, "achieved": (float(C_) * K * (25 if w["updated"] else 0) + 2.0 * C_ * (8 * (d + m + 2) + 16))
                         / (avg_ms * 1e-3) / 1e9, "peak": 6548.2, "unit": "GB/s",
                         "note": "the same kernel against the measured HBM copy bandwidth of MEASURED_PEAKS.json: algorithmic bytes "
                                 "(flags + snapshot trace, state read + written once per launch) over the launch time -- far from "
                                 "the HBM roofline, which is why the FP64 pipe is the bound reported above"},
        "clocks": clocks_summary(clk_path, local, wall0, wall1),
    }
    if cpu is not None:
        out["cpu_baseline"] = cpu
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    """The reference arm: the CPU implementation of the same path (numpy oracle = restatement of the
    reference's classes; the reference itself is pure Python and cannot travel to the GPU box), all host
    cores, each bench step a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    # the surrogate the CPU chains evaluate: same synthetic training set, fitted by the oracle on the host
    from oracle import surrogate_poly as OP
    from oracle import surrogate_rbf as ORB
    d, m = w["d"], w["m"]
    truth_SOL = None
    if w["surrogate"] == "rbf":
        X, Y = training_set(w, w["centers"], 1)
        up = ORB.RbfUpdate(d, m, kernel_type=w["kernel_type"], solver_type="direct")
    else:
        X, Y = training_set(w, 2000, 1)
        up = OP.PolyUpdate(d, m, max_degree=w["degree"])
    up.add_arrays(X, Y)
    SOL, _ = up.update()
    if w["truth"] == "rbf":
        Xt, Yt = training_set(w, w["truth_centers"], 2)
        tu = ORB.RbfUpdate(d, m, kernel_type=1, solver_type="direct")
        tu.add_arrays(Xt, Yt)
        truth_SOL, _ = tu.update()
    cores = len(os.sched_getaffinity(0))
    per_step = max(1.0, min(args.cpu_seconds, 120.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_baseline(args.workload, SOL, truth_SOL, min(per_step, 1.0), cores)
    steps, busy = 0, 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = cpu_baseline(args.workload, SOL, truth_SOL, per_step, cores)
        steps += r["value"] * per_step
        busy += per_step
    wall = time.perf_counter() - t0
    value = steps / busy
    sample = "%d processes x %d samples of %.1f s oracle DAMH steps (frozen surrogate, in-process proxies, no MPI hop)" % (
        cores, args.steps, per_step)
    out = {"impl": "reference", "metric": "surrogate-evaluated DAMH chain steps/s", "value": value, "unit": "chain steps/s",
           "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * wall / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": "%s: %s" % (args.workload, w["name"]), "note": "CPU arm: one chain per host core"},
           "cpu_baseline": {"value": value, "unit": "chain steps/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": "chain steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--chains", type=int, default=0)
    ap.add_argument("--centers", type=int, default=0)
    ap.add_argument("--solver", default="", choices=["", "direct", "nullspace"], help="dense solver of the RBF refit")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--overlap", action="store_true",
                    help="asynchronous collector: refit on a side stream with a lag of one block (measured slower on one GPU: "
                         "the refit's latency-bound kernels starve beside the FP64-bound chain kernel; DESIGN.md section 4.3)")
    ap.add_argument("--profile", type=int, default=0, help="run this many bench steps inside cudaProfilerStart/Stop and exit")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        import torch
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        # a non-default stream: kernels, events and NCCL calls all go to it; the legacy default stream
        # could not be captured into the refit's CUDA graph
        with torch.cuda.stream(torch.cuda.Stream(device=local)):
            run_ours(args)


if __name__ == "__main__":
    main()
