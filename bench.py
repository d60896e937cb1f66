#!/usr/bin/env python
"""bench.py -- surrogate-evaluated DAMH chain steps/s (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg3|cfg2|cfg5]

Workload (default cfg3, BASELINE.json configs[2], the configuration the metric and its 1e9 steps/s
target are quoted on; it fits one GPU): 65536 lockstep DAMH chains, d=2, m=1, the simple.json
Gaussian problem, cubic RBF surrogate with 4096 centers, analytic toy G on the device,
surrogate_is_updated=true: every 64 lockstep steps each rank picks its share of 256 new snapshots
(evenly spaced over the block), ONE NCCL all-gather makes the set common, the collector refits
(greedy thinning back to 4096 centers, saddle-point system build, dense direct solve: null-space
projection + Cholesky; `--solver direct` selects the LU with partial pivoting that mirrors the
reference's solver_type 'direct'), ONE broadcast of the packed state hands the new surrogate to
every rank.  The collector is asynchronous, as the reference's separate COLLECTOR process is: round b runs on a
side stream beside block b+1 and its surrogate is used from block b+2 (a deterministic lag; `--collector sync`
refits between blocks).  The headline is WEAK scaling (65536 chains per GPU); with N > 1 the same line carries
the STRONG-scaling record (65536 chains split evenly over the N GPUs) under "strong".
    cfg2: 4096 chains, frozen polynomial surrogate (degree 5)        (BASELINE.json configs[1])
    cfg5: 131072 chains/GPU, d=4, m=3, frozen cubic RBF 500 centers, frozen 2000-center RBF as G
One bench "step" = one block of 64 lockstep steps of all chains (+ the collector round in cfg3).

One JSON line on stdout (rank 0).  `value` = chain steps/s over all GPUs with state resident in
HBM; `e2e` = the same through host buffers: pinned H2D of the chain samples at the start of every
block, D2H of what a user gets back per block (final samples, counters, the per-chain running
moments of the chain, the accept / reject trace); `roofline` = the fused chain kernel against the
measured FP64 FMA peak; `roofline_refit` = the collector round's dense solve (n^3/3 flop) against
the same peak; `cpu_baseline` = the REFERENCE's own Algorithm_DAMH.run (unmodified modules from
oracle/_ref, CSV output on) on the host cores -- the numpy oracle when that copy is absent;
`other_workloads` (N = 1) = short cfg2 / cfg5 records.  NCCL's INFO log goes to stderr.
"""
import os as _os

# more hardware launch queues than the default 8: the collector adds ~8 streams (thin / solve / assemble lanes, the
# refit's side streams, one NCCL stream per communicator) and two streams sharing a queue serialise -- a KB-sized
# all-gather then waits behind a 14 ms chain kernel.  Must be set before the CUDA context exists.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

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
                 name="illustrative-shape DAMH: 65536 chains, cubic RBF 4096 centers, surrogate_is_updated, refit every 64 steps"),
    "cfg2": dict(d=2, m=1, problem=SIMPLE, chains=4096, surrogate="poly", degree=5, updated=False, block=64,
                 proposal_std=1.0, truth="toy", name="simple.json-shape DAMH: 4096 chains, frozen poly degree 5"),
    "cfg5": dict(d=4, m=3, problem=DARCY, chains=131072, surrogate="rbf", centers=500, kernel_type=1, updated=False,
                 block=64, proposal_std=0.2, truth="rbf", truth_centers=2000,
                 name="Darcy-shape surrogate-only DAMH: 131072 chains/GPU, frozen cubic RBF 500 centers, 2000-center RBF as G"),
}
METRIC = "surrogate-evaluated DAMH chain steps/s"


def static_config(wname, w):
    """The workload description both arms print verbatim (nothing measured in it)."""
    return {"workload": "%s: %s" % (wname, w["name"]), "chains_per_gpu": w["chains"], "lockstep_steps_per_bench_step": w["block"],
            "no_parameters": w["d"], "no_observations": w["m"], "surrogate": w["surrogate"], "centers": w.get("centers"),
            "poly_degree": w.get("degree"), "surrogate_is_updated": bool(w["updated"]),
            "refit_every": w["block"] if w["updated"] else None,
            "snapshots_per_refit": w.get("snapshots_per_refit") if w["updated"] else None,
            "collector": "asynchronous with a deterministic lag (collector.lag_blocks; the reference's collector is a separate process)" if w["updated"] else None,
            "proposal_std": w["proposal_std"], "true_model": w["truth"],
            "l2": "GPU arm: 256 MiB flush write before every bench step, inside the timed region"}


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
# CPU arm: the reference's own classes (oracle/_ref via oracle.ref_import) or, without them, the numpy oracle
# ----------------------------------------------------------------------------------------------------
def reference_available():
    from oracle import ref_import
    return ref_import.available()


class _Snap:
    def __init__(self, sample, G_sample, weight=1):
        self.sample, self.G_sample, self.weight = sample, G_sample, weight


def _ref_worker(args):
    """One host process = one reference sampler: Algorithm_DAMH.run (classes_SAMPLER.py:165-220) with in-process proxies
    (no MPI hop: favours the reference), CSV output on as process_SAMPLER.py:66-85 configures it."""
    wname, SOL, truth_SOL, seconds, seed = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    import contextlib
    import io
    import warnings
    from oracle import ref_import as R
    ns = R.load()
    w = WORKLOADS[wname]
    d, m = w["d"], w["m"]
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if w["surrogate"] == "rbf":
            apply_o = ns.rbf.Surrogate_apply(d, m, kernel_type=w["kernel_type"])
        else:
            apply_o = ns.poly.Surrogate_apply(d, m)

        class SurrogateProxy:               # the local-collector proxy of classes_communication.py:83-157, frozen between swaps
            queue = []

            def send_parameters(self, p):
                self.p = p.copy()

            def recv_observations(self):
                return apply_o.apply(SOL, self.p)

            def send_to_data_collector(self, s):
                self.queue.append(s)
                if len(self.queue) > 4096:
                    self.queue.clear()

        if w["truth"] == "toy":
            solver = ns.solvers.Solver_linela2exp_local()
        else:
            t_apply = ns.rbf.Surrogate_apply(d, m, kernel_type=1)

            class solver:                   # noqa: N801  (the frozen truth surrogate behind the solver interface)
                @staticmethod
                def set_parameters(p):
                    solver.p = np.array(p)

                @staticmethod
                def get_observations():
                    return t_apply.apply(truth_SOL, solver.p)[0, :]

        class SolverProxy:
            def send_parameters(self, p):
                solver.set_parameters(p)

            def recv_observations(self):
                return solver.get_observations()

        cwd = os.getcwd()
        tmp = tempfile.mkdtemp(prefix="sdb_ref_")
        os.chdir(tmp)
        try:
            prob = ns.cS.Problem_Gauss(d, saved_samples_name="bench", **w["problem"])
            prop = ns.cS.Proposal_GaussRandomWalk(d, proposal_std=w["proposal_std"], seed=seed + 1)
            alg = ns.cS.Algorithm_DAMH(prob, prop, SolverProxy(), Surrogate=SurrogateProxy(), surrogate_is_updated=bool(w["updated"]),
                                       max_samples=10 ** 9, time_limit=seconds, save_raw_data=True, name="alg000DAMH_rank%d" % seed,
                                       seed=seed + 2)
            t0 = time.perf_counter()
            alg.run()
            dt = time.perf_counter() - t0
        finally:
            os.chdir(cwd)
            import shutil
            shutil.rmtree(tmp, ignore_errors=True)
    return alg.no_accepted + alg.no_rejected + alg.no_prerejected, dt


def _port_worker(args):
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
    apply_o = ORB.RbfApply(d, m, kernel_type=w["kernel_type"]) if w["surrogate"] == "rbf" else OP.PolyApply(d, m)

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


def cpu_arm(wname, SOL, truth_SOL, seconds, cores=None):
    """All host cores, one sampler process each, `seconds` of DAMH steps: aggregate chain steps/s."""
    cores = cores or len(os.sched_getaffinity(0))
    ref = reference_available()
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        res = pool.map(_ref_worker if ref else _port_worker, [(wname, SOL, truth_SOL, seconds, 1000 * r) for r in range(cores)])
        wall = time.perf_counter() - t0
    steps = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    what = ("the reference's Algorithm_DAMH.run (unmodified modules, oracle/_ref), CSV output on, in-process proxies (no MPI hop), "
            "surrogate frozen between collector swaps") if ref else "numpy oracle DAMH steps (frozen surrogate, in-process proxies, no MPI hop)"
    return dict(value=steps / busy, unit="chain steps/s", cores=cores, kind="reference" if ref else "port",
                sample="%d processes x %.0f s of %s; %d steps" % (cores, seconds, what, steps), wall_s=wall)


def cpu_models(w):
    """The surrogate (and cfg5's frozen truth) the CPU samplers evaluate, fitted on the host by the same CPU implementation."""
    d, m = w["d"], w["m"]
    if reference_available():
        import contextlib
        import io
        import warnings
        from oracle import ref_import as R
        ns = R.load()

        def fit(kind, X, Y, **kw):
            with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
                warnings.simplefilter("ignore")
                up = (ns.rbf if kind == "rbf" else ns.poly).Surrogate_update(d, m, **kw)
                up.add_data([_Snap(X[i], Y[i]) for i in range(X.shape[0])])
                return up.update()[0]
    else:
        from oracle import surrogate_poly as OP
        from oracle import surrogate_rbf as ORB

        def fit(kind, X, Y, **kw):
            up = (ORB.RbfUpdate if kind == "rbf" else OP.PolyUpdate)(d, m, **kw)
            up.add_arrays(X, Y)
            return up.update()[0]
    t0 = time.perf_counter()
    if w["surrogate"] == "rbf":
        X, Y = training_set(w, w["centers"], 1)
        SOL = fit("rbf", X, Y, kernel_type=w["kernel_type"], solver_type="direct")
    else:
        X, Y = training_set(w, 2000, 1)
        SOL = fit("poly", X, Y, max_degree=w["degree"])
    fit_s = time.perf_counter() - t0
    truth_SOL = None
    if w["truth"] == "rbf":
        Xt, Yt = training_set(w, w["truth_centers"], 2)
        truth_SOL = fit("rbf", Xt, Yt, kernel_type=1, solver_type="direct")
    return SOL, truth_SOL, fit_s


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
def build_models(w):
    """Initial surrogate (GPU refit of a synthetic training set) and, for cfg5, the frozen truth."""
    from surrdamh_b200 import surrogate_poly as SP
    from surrdamh_b200 import surrogate_rbf as SR
    from surrdamh_b200.device import DeviceModel
    d, m = w["d"], w["m"]
    if w["surrogate"] == "rbf":
        X, Y = training_set(w, w["centers"], 1)
        updater = SR.Surrogate_update(d, m, kernel_type=w["kernel_type"], solver_type=w.get("solver_type", "direct"),
                                      no_keep=w["centers"])
    else:
        X, Y = training_set(w, 2000, 1)
        updater = SP.Surrogate_update(d, m, max_degree=w["degree"])
    updater.add_arrays(X, Y)
    sur, _ = updater.update_device()
    truth = DeviceModel.toy()
    if w["truth"] == "rbf":
        Xt, Yt = training_set(w, w["truth_centers"], 2)
        tu = SR.Surrogate_update(d, m, kernel_type=1, solver_type="direct")
        tu.add_arrays(Xt, Yt)
        truth, _ = tu.update_device()
    return updater, sur, truth


class Run:
    """One measured configuration on this rank: C_ chains of workload w, synchronous or asynchronous collector."""

    def __init__(self, w, C_, rank, world, dev, collector="sync", refit_mode="broadcast", lag=1, precision="f64", reserve_sms=16):
        import torch

        from surrdamh_b200.chains import ChainBlock, Proposal, Trace
        from surrdamh_b200.collector import AsyncCollector, Collector, PipelinedCollector
        from surrdamh_b200.device import DeviceProblem
        self.torch = torch
        self.w, self.C, self.rank, self.world, self.dev = w, C_, rank, world, dev
        d, m, K = w["d"], w["m"], w["block"]
        self.d, self.m, self.K = d, m, K
        self.updater, self.sur, self.truth = build_models(w)
        self.prob = DeviceProblem(d, **w["problem"], device=dev)
        self.blk = ChainBlock(self.prob, C_, chain0=rank * C_, seed=2024, device=dev)
        self.blk.surrogate_f32 = precision == "f32"
        self.precision = precision
        self.prop = Proposal(d, w["proposal_std"], device=dev)
        self.col = self.acol = None
        self.lag = lag if collector != "sync" else 0
        self.reserve_sms = reserve_sms if collector != "sync" else 0
        if w["updated"] and self.updater.uses_cluster_kernels():
            self.reserve_sms = 0          # the LU's 16-CTA cluster cannot be placed on scattered SMs: the round runs between blocks
        if w["updated"]:
            self.col = Collector(self.updater, "rbf", d, m, kernel_type=w["kernel_type"], rank=rank, world=world,
                                 snapshots_per_refit=w["snapshots_per_refit"], refit_mode=refit_mode, device=dev)
            self.col.model = self.sur
            if collector == "async":
                self.acol = PipelinedCollector(self.col, dev, lag=lag)
            elif collector == "async-serial":
                self.acol = AsyncCollector(self.col, dev, lag=lag)
            self.traces = [Trace(K, C_, d, m, dev, snapshots=True, full=False) for _ in range(self.lag + 1)]
        else:
            self.traces = [None]
        self.flags_only = Trace(K, C_, d, m, dev, snapshots=False, full=False) if not w["updated"] else None
        self.flush = torch.empty((256 << 20) // 8, dtype=torch.float64, device=dev)      # > 126 MB of L2
        self.state = {"sur": self.sur, "swapped": False, "b": 0, "drained": 0}
        self.kern_ev, self.round_ev = [], []

    def one_block(self, timed, host_io=None):
        """One bench step.  host_io: the e2e variant (pinned buffers).  With the asynchronous collector the snapshots of
        block b are refitted on the side stream while blocks b+1 .. b+lag run, and the surrogate of round b is
        switched in at the start of block b + lag + 1."""
        torch, st, blk, K = self.torch, self.state, self.blk, self.K
        b = st["b"]
        st["b"] += 1
        main = torch.cuda.current_stream()
        acol, col = self.acol, self.col
        if acol is not None and b > self.lag:
            R, new = acol.result(b - self.lag - 1)
            main.wait_event(R)
            acol.forget(b - self.lag - 1)
            if new is not None:
                st["sur"], st["swapped"] = new, True
        trace = self.traces[b % len(self.traces)]
        if trace is None and host_io is not None:
            trace = self.flags_only
        self.flush.fill_(1.0)
        if host_io is not None:
            host_io.begin(b)                                                  # H2D: chain samples from pinned host memory (uploaded
            blk.prepare("DAMH", st["sur"], self.truth)                        # during the previous block); G(u), posteriors of them
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        blk.run("DAMH", K, self.prop, st["sur"], self.truth, trace=trace, refresh_pre=st["swapped"], coresident=self.reserve_sms)
        e1.record()
        if timed:
            self.kern_ev.append((e0, e1, st["swapped"]))
        st["swapped"] = False
        if acol is not None:
            acol.submit(b, trace, K, e1, blk.step0, timed)
        elif col is not None:
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record()
            if col.round(trace, K, blk.step0):
                st["sur"], st["swapped"] = col.model, True
            r1.record()
            if timed:
                self.round_ev.append((r0, r1))
        if host_io is not None:
            host_io.finish(b, trace)          # D2H: what a user gets back per block, on the copy stream beside the next block
            host_io.upload(b + 1)             # H2D of the next block's samples
            if b >= 1:
                host_io.result(b - 1)         # the host consumes block b-1's product while block b runs

    def drain(self):
        """Make the sampling stream wait for every collector round issued so far (their surrogates stay queued)."""
        if self.acol is None:
            return
        main = self.torch.cuda.current_stream()
        for b in sorted(self.acol.results):
            if b >= self.state["drained"]:
                main.wait_event(self.acol.result(b)[0])
        self.state["drained"] = self.state["b"]

    def barrier(self):
        self.drain()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def timed_region(self, n_steps, host_io=None):
        torch = self.torch
        self.barrier()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        self.base_event = t0
        for _ in range(n_steps):
            self.one_block(True, host_io)
        if host_io is not None:
            host_io.drain()      # every block's product has reached the host
        self.drain()             # the last collector rounds belong to the timed region
        t1.record()
        self.barrier()
        ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def warmup(self, n):
        self.blk.prepare("DAMH", self.sur, self.truth)
        for _ in range(n):
            self.one_block(False)

    def host_buffers(self):
        from surrdamh_b200.host_io import BlockIO
        io = BlockIO(self.blk, self.K)
        io.u_in.copy_(self.blk.u.cpu())
        return io, int(io.h2d_bytes), int(io.d2h_bytes)

    def measure(self, steps, warmup, e2e_steps):
        """-> dict of measured numbers for this configuration (device-resident and end-to-end)."""
        from surrdamh_b200 import _lib as L
        self.warmup(max(warmup, 3))
        L.lib().sdb_launch_count(1)
        wall0 = time.time()
        ms_total = self.timed_region(steps)
        wall1 = time.time()
        launches = int(L.lib().sdb_launch_count(1))
        kern = [(a.elapsed_time(b), sw) for a, b, sw in self.kern_ev]
        rounds = [a.elapsed_time(b) for a, b in (self.acol.round_events if self.acol is not None else self.round_ev)]
        if os.environ.get("SDB_TIMELINE") and getattr(self.acol, "stage_events", None):
            # development: device times (ms since the start of the timed region) of every block and round of this rank
            base = self.base_event
            tl = {"rank": self.rank, "chains": self.C,
                  "blocks": [[base.elapsed_time(a), base.elapsed_time(b)] for a, b, _ in self.kern_ev],
                  "rounds": [[base.elapsed_time(e) if e is not None else None for e in (t0_, g, t, s_, m_)]
                             for t0_, t, s_, m_, g in self.acol.stage_events],
                  "columns": "blocks: chain kernel start, end; rounds: start, gathered, thinned, solved (duty rank), assembled"}
            with open("%s_c%d_%s_rank%d.json" % (os.environ["SDB_TIMELINE"], self.C, self.precision, self.rank), "w") as f:
                json.dump(tl, f)
        stages = None
        if getattr(self.acol, "stage_events", None):
            ev = self.acol.stage_events
            solved = [t.elapsed_time(s_) for _, t, s_, _, _ in ev if s_ is not None]
            stages = {"thin": float(np.mean([a.elapsed_time(t) for a, t, _, _, _ in ev])),
                      "select_and_gather": float(np.mean([a.elapsed_time(g) for a, _, _, _, g in ev])),
                      "assemble_after_thin": float(np.mean([t.elapsed_time(m_) for _, t, _, m_, _ in ev])),
                      "solve_on_duty_rank": float(np.mean(solved)) if solved else None, "solves_here": len(solved)}
            ev.clear()
        self.kern_ev.clear(); self.round_ev.clear()
        if self.acol is not None:
            self.acol.round_events.clear()
        io, h2d, d2h = self.host_buffers()
        self.one_block(False, io)
        ms_e2e = self.timed_region(e2e_steps, io)
        self.kern_ev.clear(); self.round_ev.clear()
        if self.acol is not None:
            self.acol.round_events.clear()
            if getattr(self.acol, "stage_events", None):
                self.acol.stage_events.clear()
        failures = 0
        if self.col is not None:
            self.col.poll_status(wait=True)
            failures = sum(1 for s in self.col.status_log if s[1])
        total = float(self.C) * self.K * steps * self.world
        return dict(value=total / (ms_total * 1e-3), ms_per_step=ms_total / steps, steps=steps, launches=launches, kern=kern,
                    rounds=rounds, stages=stages, wall=(wall0, wall1), refit_failures=failures,
                    e2e=dict(value=float(self.C) * self.K * e2e_steps * self.world / (ms_e2e * 1e-3), unit="chain steps/s",
                             h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, steps=e2e_steps, ms_per_step=ms_e2e / e2e_steps,
                             returns="per block and chain: final sample, counters, running first / second moments of the chain, "
                                     "accept / reject flag of every step; double-buffered pinned transfers on a copy stream "
                                     "(surrdamh_b200/host_io.py): block b's download and block b+1's upload run beside a chain "
                                     "kernel, the host consumes a block's product one block later"))


def flops_per_point(w, sur):
    d, m = w["d"], w["m"]
    if w["surrogate"] == "rbf":
        return w["centers"] * (3 * d + 3 + 2 * m) + m * (2 * d + 1)
    P = int(sur.alpha.shape[0])
    return d * w["degree"] + d * (w["degree"] + 1) ** 2 + P * (d + 2 * m)


def fp64_peaks(dev):
    """Live FP64 FMA throughput of this GPU (roofline denominator; MEASURED_PEAKS.json has no FP64 entry)."""
    import torch

    from surrdamh_b200 import _lib as L
    info = L.device_info()
    pb, iters = info["sm_count"] * 8, 20000
    pout = torch.empty((pb * 256,), dtype=torch.float64, device=dev)
    out = []
    for fn in (L.lib().sdb_fp64_probe, L.lib().sdb_fp64_probe3):
        pout.zero_()
        L.check(fn(iters, pb, L.ptr(pout), L.stream_handle()))
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        L.check(fn(iters, pb, L.ptr(pout), L.stream_handle()))
        p1.record()
        p1.synchronize()
        out.append(pb * 256 * 8 * 2.0 * iters / (p0.elapsed_time(p1) * 1e-3) / 1e12)
    return out


def fp32_peak(dev):
    import torch

    from surrdamh_b200 import _lib as L
    info = L.device_info()
    pb, iters = info["sm_count"] * 8, 40000
    pout = torch.empty((pb * 256,), dtype=torch.float32, device=dev)
    L.check(L.lib().sdb_fp32_probe(iters, pb, L.ptr(pout), L.stream_handle()))
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    L.check(L.lib().sdb_fp32_probe(iters, pb, L.ptr(pout), L.stream_handle()))
    p1.record()
    p1.synchronize()
    return pb * 256 * 8 * 2.0 * iters / (p0.elapsed_time(p1) * 1e-3) / 1e12


def f32_decision_check(w, run, dev):
    """One lockstep step from the SAME state with the same random numbers, surrogate in FP64 and in FP32: fraction of chains
    whose stage-1 decision differs, and the relative difference of the surrogate log-posteriors."""
    import torch

    from surrdamh_b200.chains import ChainBlock, Trace
    d, m, C_ = w["d"], w["m"], run.C
    out = {}
    for f32 in (False, True):
        blk = ChainBlock(run.prob, C_, chain0=0, seed=77, device=dev)
        blk.u.copy_(run.blk.u)
        blk.surrogate_f32 = f32
        blk.prepare("DAMH", run.state["sur"], run.truth)
        tr = Trace(1, C_, d, m, dev, snapshots=False, full=True)
        blk.run("DAMH", 1, run.prop, run.state["sur"], run.truth, trace=tr)
        torch.cuda.synchronize(dev)
        out[f32] = (tr.flags[0].cpu().numpy(), tr.pre[0].cpu().numpy())
    p64, p32 = out[False][1], out[True][1]
    rel = np.abs(p32 - p64) / np.maximum(1.0, np.abs(p64))
    return {"stage1_decision_mismatch_rate": float(np.mean((out[False][0] != 0) != (out[True][0] != 0))),
            "accept_mismatch_rate": float(np.mean((out[False][0] == 1) != (out[True][0] == 1))),
            "pre_posterior_rel_diff_median": float(np.median(rel)), "pre_posterior_rel_diff_p99": float(np.percentile(rel, 99)),
            "chains": int(C_), "note": "one step from the same state and random numbers; stage 2 uses the true posterior in FP64 "
                                       "either way, so the sampled distribution is the exact posterior in both runs"}


def solve_alone_ms(run, reps=5):
    """The collector's dense solve of the current center set, alone on the GPU (default stream), ms per solve."""
    torch, up = run.torch, run.updater
    n0 = int(up.no_keep)
    if up._par.shape[0] != n0 or not hasattr(up, "solve_stage"):
        return None
    run.drain()
    torch.cuda.synchronize(run.dev)
    X, Y = up._par.clone(), up._obs.clone()
    meta = torch.tensor([float(n0), 0.0, 0.0, 0.0], dtype=torch.float64, device=run.dev)
    res = up.new_result()
    for _ in range(2):
        up.solve_stage(X, Y, meta, res)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        up.solve_stage(X, Y, meta, res)
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / reps


def kernel_alone_ms(run, reps=3):
    """The fused chain kernel of this run's shape on ALL SMs, nothing running beside it (ms per launch)."""
    torch, blk = run.torch, run.blk
    run.drain()
    torch.cuda.synchronize(run.dev)
    sur = run.state["sur"]
    blk.run("DAMH", run.K, run.prop, sur, run.truth, trace=run.traces[0])
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        blk.run("DAMH", run.K, run.prop, sur, run.truth, trace=run.traces[0])
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / reps


def kernel_name(w):
    if w["surrogate"] == "poly" and w["d"] == 2 and w["m"] == 1 and not os.environ.get("SDB_NO_LEAN"):
        return "chain_lean_kernel<%d,true>" % w["degree"]
    if w["surrogate"] == "rbf" and w.get("centers", 0) >= 1024 and not os.environ.get("SDB_NO_WT"):
        return "chain_kernel<%d,%d,0,true>" % (w["d"], w["m"])      # warp-owned chains (csrc/sdb_chain_wt.cu)
    return "chain_kernel<%d,%d>" % (w["d"], w["m"])


def roofline_of(w, run, rec, peak_tf, peak3_tf, headline):
    d, m, C_, K = w["d"], w["m"], run.C, run.K
    per_point = flops_per_point(w, run.sur)
    avg_ms = float(np.mean([k for k, _ in rec["kern"]]))
    frac_sw = float(np.mean([1.0 if sw else 0.0 for _, sw in rec["kern"]]))
    flops_launch = float(C_) * (K + frac_sw) * per_point
    achieved = flops_launch / (avg_ms * 1e-3) / 1e12
    out = {"bound": "fp64-pipe", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
           "traffic": 131635456 if headline else None, "kernel": kernel_name(w), "kernel_ms": avg_ms,
           "kernel_share_of_step": avg_ms * len(rec["kern"]) / (rec["ms_per_step"] * rec["steps"]), "flops_per_launch": flops_launch}
    plan = run.blk.plan("DAMH", run.state["sur"], run.truth, run.prop)
    if run.reserve_sms and plan["lanes_per_chain"] == 0:
        from surrdamh_b200 import _lib as L
        sms = L.device_info()["sm_count"]
        used = max(1, sms - run.reserve_sms)
        out.update({"sms_used": used, "sms_total": sms, "frac_of_the_sms_used": achieved / (peak_tf * used / sms),
                    "sms_note": "the launch leaves %d SMs to the asynchronous collector's kernels; `frac` is against the peak of "
                                "the WHOLE GPU" % run.reserve_sms})
        alone = kernel_alone_ms(run)
        if alone:
            fl1 = float(C_) * K * per_point
            out["alone_on_all_sms"] = {"kernel_ms": alone, "achieved": fl1 / (alone * 1e-3) / 1e12, "frac": fl1 / (alone * 1e-3) / 1e12 / peak_tf,
                                       "note": "the same kernel launched on every SM with nothing beside it (what --collector sync "
                                               "does), measured after the run: the kernel's own quality"}
    if headline:
        out.update({"traffic_note": "DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the ncu --set full capture in "
                                    "profiles/r02_chain_kernel_ncu.md; algorithmic bytes = C*K*25 B trace + state = 1.05e8",
                    "peak_three_register_fma": peak3_tf,
                    "peak_source": "sdb_fp64_probe (8 independent DFMA chains/thread), measured live in this run; "
                                   "MEASURED_PEAKS.json has no FP64 entry"})
    return out


def nccl_log_to_stderr(rank):
    """NCCL's INFO log (communicator lines: nranks, rings, NVLS) must be visible, but stdout carries exactly one JSON
    line: every rank logs into its own temporary file and copies it to stderr when the process ends."""
    if os.environ.get("SDB_KEEP_NCCL_DEBUG"):
        return
    import atexit
    path = os.path.join(tempfile.gettempdir(), "sdb_nccl_%d_rank%d.log" % (os.getpid(), rank))
    os.environ["NCCL_DEBUG"] = "INFO"
    os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT,ENV")
    os.environ["NCCL_DEBUG_FILE"] = path

    def dump():
        try:
            with open(path) as f:
                sys.stderr.write(f.read())
            sys.stderr.flush()
            os.remove(path)
        except OSError:
            pass
    atexit.register(dump)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from surrdamh_b200 import _lib as L

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    L.require_cuda()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        nccl_log_to_stderr(rank)
        dist.init_process_group("nccl", device_id=dev)
    wname = args.workload
    w = dict(WORKLOADS[wname])
    if args.chains:
        w["chains"] = args.chains
    if args.centers and w["surrogate"] == "rbf":
        w["centers"] = args.centers
    if args.solver and w["surrogate"] == "rbf":
        w["solver_type"] = args.solver
    collector = args.collector or "async"
    refit_mode = args.refit_mode or ("rotate" if world > 1 else "broadcast")
    # How many SMs the collector needs beside the chains, and the lag that goes with it.  One GPU solves every round itself: the
    # dense solve (2.3e10 flop) must fit one block period, which takes ~15 SMs (lag 1).  With the solve duty rotating over N GPUs a
    # rank solves every N-th round and may take N periods for it, so it gets by with fewer SMs (and the chains get more), at the
    # price of a surrogate that is a block or two older: 11 SMs / lag 2 on 2 GPUs, 6 SMs / lag 3 from 4 GPUs on (measured, 4 GPUs:
    # 14.55 -> 13.66 ms per step).  The strong-scaling record keeps 15 SMs: its blocks shrink with N, its solve window does not grow.
    auto_r, auto_lag = {1: (15, 1), 2: (11, 2)}.get(world, (6, 3)) if refit_mode == "rotate" or world == 1 else (15, 1)
    # the cooperative backward solve of the refit must fit on the reserved SMs (two CTAs each): 16 CTAs where >= 11 SMs are
    # reserved, the library's default of 8 (n <= 4096) with 6.  Read once by the library, before its first fit.
    if world <= 2:
        os.environ.setdefault("SDB_BWD_GRID", "16")
    weak_reserve = args.reserve_sms or auto_r
    weak_lag = args.lag or auto_lag
    strong_reserve = args.reserve_sms or 15

    # ---- CPU baseline (rank 0, N == 1 only), before the timed GPU region -------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.profile:
        SOL, truth_SOL, fit_s = cpu_models(WORKLOADS[wname])
        cpu = cpu_arm(wname, SOL, truth_SOL, args.cpu_seconds)
        if WORKLOADS[wname]["surrogate"] == "rbf":
            cpu["refit_s"] = fit_s
            cpu["refit_note"] = "one full Surrogate_update.update() of the same CPU implementation at n = %d (solver_type 'direct', all host threads)" % WORKLOADS[wname]["centers"]

    run = Run(w, w["chains"], rank, world, dev, collector, refit_mode, weak_lag, reserve_sms=weak_reserve)
    if args.profile:
        run.warmup(max(args.warmup, 3))
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.start()
        for _ in range(args.profile):
            run.one_block(False)
        run.drain()
        torch.cuda.synchronize(dev)
        torch.cuda.profiler.stop()
        print(json.dumps({"profile_blocks": args.profile, "note": "no measurement taken in profile mode"}))
        return
    clk_path = os.path.join(tempfile.gettempdir(), "sdb_clocks_%d.csv" % os.getpid())
    mon = clocks_monitor_start(clk_path) if rank == 0 else None
    peak_tf, peak3_tf = fp64_peaks(dev)
    e2e_steps = max(2, args.steps)          # the same number of blocks as the device-resident measurement: both amortise the final drain of the collector equally
    rec = run.measure(args.steps, args.warmup, e2e_steps)
    if mon is not None:
        mon.terminate()
        mon.wait()

    # ---- strong scaling: the same 65536 chains split evenly over the ranks ------------------------------------
    strong = None
    if world > 1 and w["updated"] and not args.no_extras and w["chains"] % world == 0:
        slag = args.strong_lag or world
        srun = Run(w, w["chains"] // world, rank, world, dev, collector, refit_mode, slag, reserve_sms=strong_reserve)
        srec = srun.measure(args.steps, args.warmup, e2e_steps)
        strong = {"scaling": "strong", "chains_total": w["chains"], "chains_per_gpu": w["chains"] // world, "value": srec["value"],
                  "unit": "chain steps/s", "ms_per_step": srec["ms_per_step"], "e2e": srec["e2e"],
                  "kernel_ms": float(np.mean([k for k, _ in srec["kern"]])),
                  "round_ms": float(np.mean(srec["rounds"])) if srec["rounds"] else None, "gpu_launches": srec["launches"],
                  "lag_blocks": srun.lag, "reserved_sms": srun.reserve_sms, "stages_ms": srec.get("stages")}
        del srun

    # ---- the FP32-surrogate variant of the same workload (second line; the headline stays FP64) ---------------------------
    fp32 = None
    if w["surrogate"] == "rbf" and not args.no_extras:
        frun = Run(w, w["chains"], rank, world, dev, collector, refit_mode, weak_lag, precision="f32", reserve_sms=strong_reserve)
        frec = frun.measure(max(10, args.steps // 2), 3, 2)
        if rank == 0:
            p32 = fp32_peak(dev)
            per_point = flops_per_point(w, frun.sur)
            k_ms = float(np.mean([k for k, _ in frec["kern"]]))
            frac_sw = float(np.mean([1.0 if sw else 0.0 for _, sw in frec["kern"]]))
            fl = float(frun.C) * (frun.K + frac_sw) * per_point
            fp32 = {"surrogate_precision": "f32", "value": frec["value"], "unit": "chain steps/s", "ms_per_step": frec["ms_per_step"],
                    "steps": frec["steps"], "e2e": frec["e2e"], "gpu_launches": frec["launches"],
                    "roofline": {"bound": "fp32-pipe", "achieved": fl / (k_ms * 1e-3) / 1e12, "peak": p32, "unit": "TFLOP/s",
                                 "frac": fl / (k_ms * 1e-3) / 1e12 / p32, "kernel": "chain_kernel<%d,%d,1>" % (w["d"], w["m"]),
                                 "kernel_ms": k_ms, "flops_per_launch": fl,
                                 "peak_source": "sdb_fp32_probe (8 independent FFMA chains/thread), measured live",
                                 "note": "same algorithmic flop count as the FP64 kernel (SURVEY 8d: 3d+3+2m per pair); per pair "
                                         "7 FP32 instructions + 1 MUFU.RSQ at d=2, m=1"},
                    "vs_f64": f32_decision_check(w, frun, dev)}
        del frun

    # ---- short records of the other single-GPU workloads ---------------------------------------------------------
    others = {}
    if world == 1 and not args.no_extras and wname == "cfg3" and not args.chains and not args.centers:
        for on in ("cfg2", "cfg5"):
            ow = dict(WORKLOADS[on])
            orun = Run(ow, ow["chains"], rank, world, dev)
            orec = orun.measure(10, 3, 2)
            others[on] = {"config": static_config(on, ow), "value": orec["value"], "unit": "chain steps/s",
                          "ms_per_step": orec["ms_per_step"], "steps": 10, "e2e": orec["e2e"],
                          "roofline": roofline_of(ow, orun, orec, peak_tf, peak3_tf, False), "gpu_launches": orec["launches"]}
            del orun

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    d, m, C_, K = w["d"], w["m"], w["chains"], w["block"]
    rl = roofline_of(w, run, rec, peak_tf, peak3_tf, wname == "cfg3" and not args.chains and not args.centers)
    avg_ms = rl["kernel_ms"]
    hbm_bytes = float(C_) * K * (25 if w["updated"] else 0) + 2.0 * C_ * (8 * (d + m + 2) + 16)
    out = {
        "metric": METRIC, "value": rec["value"], "unit": "chain steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": rec["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": static_config(wname, w),
        "e2e": rec["e2e"], "gpu_launches": rec["launches"], "roofline": rl,
        "roofline_hbm": {"bound": "hbm", "frac": hbm_bytes / (avg_ms * 1e-3) / 1e9 / 6548.2, "achieved": hbm_bytes / (avg_ms * 1e-3) / 1e9,
                         "peak": 6548.2, "unit": "GB/s",
                         "note": "the same kernel against the measured HBM copy bandwidth of MEASURED_PEAKS.json: algorithmic bytes "
                                 "(flags + snapshot trace, state read + written once per launch) over the launch time -- far from "
                                 "the HBM roofline, which is why the FP64 pipe is the bound reported above"},
        "clocks": clocks_summary(clk_path, local, *rec["wall"]),
    }
    if rec["rounds"]:
        n_c = w["centers"]
        round_ms = float(np.mean(rec["rounds"]))
        flop = n_c ** 3 / 3.0
        out["collector"] = {"mode": collector, "refit_mode": refit_mode, "lag_blocks": run.lag, "reserved_sms": run.reserve_sms, "round_ms": round_ms, "stages_ms": rec.get("stages"),
                            "rounds": len(rec["rounds"]), "refit_failures": rec["refit_failures"], "solver": w.get("solver_type"),
                            "system_size": n_c + d + 1,
                            "what": "select + all-gather + masked thinning + system build + dense solve + broadcast; "
                                    + ("between blocks on the sampling stream" if collector == "sync" else
                                       ("two pipelined stages on side streams (thin: every rank; solve: the rank on duty), "
                                        if collector == "async" else "one serial round on a side stream, ")
                                       + "surrogate of round b used from block b+%d; round_ms = latency from the end of "
                                         "the block to the assembled model, on the %d SMs the chain kernel leaves free"
                                       % (run.lag + 1, run.reserve_sms))}
        alone_ms = solve_alone_ms(run) if w.get("solver_type") in ("nullspace", "direct") else None
        ref_ms = alone_ms if alone_ms else round_ms
        out["roofline_refit"] = {"bound": "fp64-tensor (DMMA)", "achieved": flop / (ref_ms * 1e-3) / 1e12, "peak": peak_tf,
                                 "unit": "TFLOP/s", "frac": flop / (ref_ms * 1e-3) / 1e12 / peak_tf, "flops_per_round": flop,
                                 "solve_alone_ms": alone_ms, "round_ms": round_ms,
                                 "note": "n^3/3 flop of the Cholesky over the dense solve (system build + null-space projection + "
                                         "factorisation + substitutions) timed ALONE on the whole GPU after the run, same live FP64 "
                                         "peak; at n = 4096 the solve is latency-bound (DESIGN.md 4.2).  Inside the run it shares "
                                         "the GPU with the chain kernel: collector.stages_ms"}
    if fp32 is not None:
        out["fp32_surrogate"] = fp32
    if strong is not None:
        out["strong"] = strong
    if others:
        out["other_workloads"] = others
    if cpu is not None:
        out["cpu_baseline"] = cpu
    sys.stdout.flush()
    print(json.dumps(out))
    sys.stdout.flush()
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    """The reference arm: the reference's own CPU implementation of the path (Algorithm_DAMH.run of the unmodified modules
    in oracle/_ref; the numpy oracle when that copy is absent), all host cores, each bench step a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wname = args.workload
    w = WORKLOADS[wname]
    SOL, truth_SOL, fit_s = cpu_models(w)
    cores = len(os.sched_getaffinity(0))
    per_step = max(1.0, min(args.cpu_seconds, 120.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_arm(wname, SOL, truth_SOL, min(per_step, 1.0), cores)
    steps, busy, kind = 0.0, 0.0, "port"
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = cpu_arm(wname, SOL, truth_SOL, per_step, cores)
        kind = r["kind"]
        steps += r["value"] * per_step
        busy += per_step
    wall = time.perf_counter() - t0
    value = steps / busy
    what = ("the reference's Algorithm_DAMH.run (unmodified modules from oracle/_ref), CSV output on" if kind == "reference"
            else "numpy oracle DAMH steps")
    sample = "%d processes x %d samples of %.1f s of %s, in-process proxies (no MPI hop), surrogate frozen between collector swaps" % (
        cores, args.steps, per_step, what)
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "chain steps/s",
           "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": 1e3 * wall / max(1, args.steps), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": static_config(wname, w),
           "arm_note": "CPU arm: one chain per host core; the collector's refit is timed separately (refit_s) because the reference "
                       "runs it in its own process beside the samplers",
           "cpu_baseline": {"value": value, "unit": "chain steps/s", "cores": cores, "kind": kind, "sample": sample,
                            "refit_s": fit_s if w["surrogate"] == "rbf" else None},
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
    ap.add_argument("--no-extras", action="store_true", help="skip the strong-scaling record (N > 1) and the cfg2 / cfg5 records (N = 1)")
    ap.add_argument("--collector", default="", choices=["", "sync", "async", "async-serial"],
                    help="async (default): the collector round runs on a side stream beside the next block and its surrogate is "
                         "used from block b+2 on, like the reference's separate collector process; sync: between blocks")
    ap.add_argument("--refit-mode", default="", choices=["", "broadcast", "rotate", "replicated"])
    ap.add_argument("--lag", type=int, default=0, help="blocks of lag of the asynchronous collector, weak-scaling record (0: by the number of GPUs)")
    ap.add_argument("--strong-lag", type=int, default=0, help="the same for the strong-scaling record (0: the number of GPUs)")
    ap.add_argument("--reserve-sms", type=int, default=0,
                    help="SMs the chain kernel leaves to the asynchronous collector's kernels, weak-scaling record (0: by the number of GPUs)")
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
