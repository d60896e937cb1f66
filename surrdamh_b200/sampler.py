"""The sampler stage loop for C lockstep chains per GPU.

Host-side counterpart of surrDAMH/process_SAMPLER.py:25-88 (build problem / proposal / proxies, initial
samples, run the stages of `samplers_list` in order handing `current_sample` on) with the chain loop
of classes_SAMPLER.py:147-220 executed by the fused CUDA kernel (csrc/sdb_chain.cu) for all chains
at once, and the collector round-trip (process_COLLECTOR.py) reduced to Collector.round.

Two data paths per stage:
  fused : the solver plugin offers device_model() -> K steps per kernel launch, G evaluated in-kernel
  split : any other plugin (the reference's set_parameters / get_observations on the host) ->
          per step: PROPOSE kernel, host G for the proposals that passed the surrogate test, DECIDE kernel
Schedule (deterministic; SURVEY.md section 7 trap 7): in a stage with surrogate_is_updated the snapshots of every
`refit_every` steps (and of the last, shorter block) are exchanged and the surrogate is refitted;
the new model is used from the next step on (`collector_async`: from `refit_lag` blocks later, the
refit running beside the chains).  A DAMH stage with no fitted surrogate is an error, as in the reference (trap 10).
"""
import time

import numpy as np
import torch

from . import _lib as L
from .chains import ChainBlock, Proposal, Trace
from .collector import AsyncCollector, Collector, PipelinedCollector
from .device import DeviceProblem
from .lhs_normal import lhs_normal


class StageResult:
    def __init__(self, index, kind, steps, counters, seconds, refits):
        self.index, self.kind, self.steps = index, kind, steps
        self.counters = counters                    # [C][4] accepted, rejected, prerejected, rej_current
        self.seconds, self.refits = seconds, refits

    @property
    def acceptance_rate(self):
        return self.counters[:, 0] / np.maximum(1, self.steps)

    def summary(self):
        acc, rej, pre = [int(v) for v in self.counters[:, :3].sum(axis=0)]
        return dict(stage=self.index, type=self.kind, steps=self.steps, chains=int(self.counters.shape[0]), accepted=acc,
                    rejected=rej, prerejected=pre, seconds=self.seconds, refits=self.refits)


class LockstepSampler:
    def __init__(self, conf, no_chains=None, rank=0, world=1, group=None, device=None, solver=None, trace_sink=None,
                 streams=None, moments=False):
        """conf: surrdamh_b200.configuration.Configuration.  no_chains: chains of THIS rank (default
        conf.no_chains).  streams: optional callable (stage, step0, K) -> (z [K][d][C], uni [K][2][C]) device
        tensors replacing Philox (parity runs).  trace_sink: optional callable (stage_index, step0, K, Trace)
        receiving the per-step trace of every launch (CSV writer, tests)."""
        L.require_cuda()
        self.conf = conf
        self.rank, self.world, self.group = rank, world, group
        self.device = torch.device(device or "cuda")
        self.C = int(no_chains if no_chains is not None else conf.no_chains)
        self.d, self.m = conf.no_parameters, conf.no_observations
        pp = conf.problem_parameters
        self.problem = DeviceProblem(self.d, prior_mean=pp["prior_mean"], prior_std=pp["prior_std"], noise_std=pp["noise_std"],
                                     no_observations=self.m, observations=pp["observations"], device=self.device)
        self.solver = solver if solver is not None else conf.child_solver_init(**conf.child_solver_parameters)
        self.truth = None
        if conf.device_solver and hasattr(self.solver, "device_model"):
            self.truth = self.solver.device_model()
        # host-evaluated G with no_solvers > 1: a pool of solver processes, as the reference's process_SOLVER / process_CHILD
        self.pool = None
        if self.truth is None and solver is None and getattr(conf, "no_full_solvers", 1) > 1 and "solver_module_path" in conf.conf:
            from .solver_pool import HostSolverPool
            self.pool = HostSolverPool.from_conf(conf)
        self.trace_sink, self.streams = trace_sink, streams
        self.chain0 = rank * self.C                               # global index of this rank's first chain
        self.block = ChainBlock(self.problem, self.C, initial=self._initial_samples(), chain0=self.chain0, seed=conf.seed,
                                moments=moments, device=self.device)
        self.block.surrogate_f32 = conf.surrogate_precision == "f32"
        self.collector = None
        if conf.use_surrogate:
            kt = conf.surr_updater_parameters.get("kernel_type", 0) if conf.surrogate_type == "rbf" else 0
            self.apply_kernel_type = conf.surr_solver_parameters.get("kernel_type", 0) if conf.surrogate_type == "rbf" else 0
            # every rank owns an updater: rank 0 (or every replica) refits on the general path, any rank can on the
            # fixed-shape path, where the packed state travels with the broadcast
            kw = {k: v for k, v in conf.surr_updater_parameters.items() if k not in ("no_parameters", "no_observations")}
            if conf.surrogate_type == "rbf":
                kw.setdefault("drop_coincident", conf.drop_coincident)
            updater = conf.surr_updater_init(self.d, self.m, **kw)
            self.collector = Collector(updater, conf.surrogate_type, self.d, self.m, kernel_type=self.apply_kernel_type,
                                       group=group, rank=rank, world=world, snapshots_per_refit=conf.snapshots_per_refit,
                                       refit_mode=conf.refit_mode, device=self.device)
        self.global_step = 0
        self.results = []
        self.no_G_calls_host = 0
        self.overlap_stats = {"queued_while_busy": 0}     # split path: batches queued on the pool while another was in flight

    # process_SAMPLER.py:48-52 -- LHS rows `rank` of an N-point design, or the prior mean
    def _initial_samples(self):
        if self.conf.initial_sample_type == "lhs":
            total = self.C * self.world      # the reference's O(n^2) memory is avoided, any chain count works
            pp = self.conf.problem_parameters
            pts = lhs_normal(self.d, pp["prior_mean"], pp["prior_std"], total, 0)
            return pts[self.rank * self.C:(self.rank + 1) * self.C]
        return None

    # ---- one stage -------------------------------------------------------------------------------------
    def _proposal(self, st):
        if st.get("proposal_differs", False):
            per = st["proposal_std"][self.chain0:self.chain0 + self.C]
            return Proposal(self.d, per, per_chain=True, device=self.device)
        return Proposal(self.d, st["proposal_std"], device=self.device)

    def _host_G(self, params):
        if self.pool is not None:
            out = self.pool.evaluate(params)
            self.no_G_calls_host += params.shape[0]
            return out.reshape(params.shape[0], self.m)
        out = np.empty((params.shape[0], self.m))
        for i, p in enumerate(params):
            self.solver.set_parameters(p)
            out[i] = np.asarray(self.solver.get_observations(), dtype=float).reshape(-1)
        self.no_G_calls_host += params.shape[0]
        return out

    def run_stage(self, i, st):
        kind, n_steps = st["type"], int(st["max_samples"])
        time_limit = float(st.get("time_limit", float("inf")))
        blk, col = self.block, self.collector
        collects = col is not None and st["surrogate_is_updated"] is True
        prop = self._proposal(st)
        if "initial_sample" in st:
            blk.set_current(st["initial_sample"])
        blk.reset_counters()
        blk.stream_id = i                                   # a fresh acceptance / proposal sub-stream per stage (seed0+2+i there)
        surrogate = col.model if (col is not None and kind == "DAMH") else None
        if kind == "DAMH" and surrogate is None:
            raise TypeError("cannot unpack non-iterable NoneType object: DAMH stage %d starts before any surrogate was "
                            "fitted (the reference fails the same way; run an MH stage with surrogate_is_updated first)" % i)
        t0 = time.time()
        if self.truth is not None:
            blk.prepare(kind, surrogate, self.truth)
        else:
            blk.prepare(kind, surrogate, None, G_host=self._host_G(blk.current_numpy()))
        if hasattr(self.trace_sink, "begin_stage"):
            self.trace_sink.begin_stage(i, st, self)
        R = self.conf.refit_every if collects else min(n_steps, 1024)
        refits0 = col.no_refits if col is not None else 0
        done, swapped = 0, False
        use_async = collects and self.conf.collector_async and self.truth is not None
        acol = None
        if use_async:
            acol = (PipelinedCollector if self.conf.collector_pipelined else AsyncCollector)(col, self.device, lag=self.conf.refit_lag)
        traces, b = [], 0
        main = torch.cuda.current_stream(self.device)
        while done < n_steps:
            K = min(R, n_steps - done)
            need_trace = collects or self.trace_sink is not None
            trace = None
            if need_trace:
                slot = b % (acol.lag + 1) if acol is not None else 0        # a round still reads the trace of its block
                while len(traces) <= slot:
                    traces.append(None)
                if traces[slot] is None or traces[slot].K < K:
                    traces[slot] = Trace(K, self.C, self.d, self.m, self.device, snapshots=collects,
                                         full=self.trace_sink is not None or self.truth is None)
                trace = traces[slot]
            if acol is not None and b > acol.lag:
                ev, new = acol.result(b - acol.lag - 1)
                main.wait_event(ev)
                acol.forget(b - acol.lag - 1)
                if new is not None and kind == "DAMH":
                    surrogate, swapped = new, True
            strm = None if self.streams is None else self.streams(i, done, K)
            if self.truth is not None:
                reserve = 0
                if acol is not None and not getattr(col.updater, "uses_cluster_kernels", lambda: False)():
                    reserve = self.conf.reserve_sms
                blk.run(kind, K, prop, surrogate, self.truth, trace=trace, streams=strm, refresh_pre=swapped, coresident=reserve)
            else:
                self._run_split(kind, K, prop, surrogate, trace, strm, swapped, collects)
            swapped = False
            if self.trace_sink is not None:
                self.trace_sink(i, done, K, trace)
            done += K
            self.global_step += K
            if acol is not None:
                ev = torch.cuda.Event()
                ev.record(main)
                acol.submit(b, trace, K, ev, self.global_step)
            elif collects:
                if col.round(trace, K, self.global_step):
                    if kind == "DAMH":
                        surrogate, swapped = col.model, True
            b += 1
            if time_limit != float("inf") and self._time_is_up(t0, time_limit):
                break
        if acol is not None:
            # the stage ends with every issued round finished and its surrogate in place (the next stage starts from it)
            for bb in sorted(acol.results):
                main.wait_event(acol.result(bb)[0])
            acol.close()
        torch.cuda.synchronize(self.device)
        if col is not None:
            col.poll_status(wait=True)
        if hasattr(self.trace_sink, "end_stage"):
            self.trace_sink.end_stage(i, st, self, done)
        res = StageResult(i, kind, done, blk.counters_numpy(), time.time() - t0, (col.no_refits - refits0) if col is not None else 0)
        self.results.append(res)
        return res

    def _time_is_up(self, t0, time_limit):
        """classes_SAMPLER.py:156,215 checks wall time of finished work after every step; here after every block, and
        the decision is COLLECTIVE (max over ranks), because the collector round inside the loop is: a rank leaving
        alone would pair its next collectives with the wrong block.  The stream is only synchronised when the host is
        within one block's duration of the limit (or has no estimate yet); queued-but-unfinished blocks are few since
        every round's gather keeps the ranks in step."""
        elapsed = time.time() - t0
        est = getattr(self, "_block_seconds", None)
        if est is None or elapsed + 4.0 * est > time_limit:
            torch.cuda.current_stream(self.device).synchronize()
            now = time.time() - t0
            nb = getattr(self, "_blocks_seen", 0) + 1
            self._blocks_seen = nb
            self._block_seconds = now / nb if est is None else max(est, 0.0) * 0.5 + 0.5 * (now - getattr(self, "_last_sync", 0.0))
            self._last_sync = now
            elapsed = now
        up = elapsed > time_limit
        if self.world > 1:
            import torch.distributed as dist
            flag = torch.tensor([1 if up else 0], dtype=torch.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
            up = bool(flag.item())
        return up

    def _run_split(self, kind, K, prop, surrogate, trace, strm, swapped, collects):
        """K steps with the true model evaluated by the host plugin (one PROPOSE / DECIDE kernel pair per step).

        With a solver pool the chains are worked in two halves, software-pipelined against the pool: while the workers
        evaluate G for the stage-1 survivors of half A, the GPU decides half B's previous step and proposes its next one,
        whose survivors are queued behind A's -- the solvers never wait for the GPU or the copies (the reference's pool
        keeps its solvers busy by queueing the samplers' requests, process_SOLVER.py:60-106).  Every chain still takes
        its steps in order with its own (chain, step) random numbers, so the result is identical to the unsplit loop."""
        blk, C_ = self.block, self.C
        one = Trace(1, C_, self.d, self.m, self.device, snapshots=collects)
        Gd = torch.zeros((self.m, C_), dtype=torch.float64, device=self.device)
        halves = [(0, C_)] if (self.pool is None or C_ < 2) else [(0, C_ // 2), (C_ // 2, C_ - C_ // 2)]
        base = blk.step0
        inflight = {}                      # half index -> (step k, chain indices of the survivors, handle | G array)

        def stream_of(k):
            return None if strm is None else (strm[0][k:k + 1].contiguous(), strm[1][k:k + 1].contiguous())

        def propose(h, k):
            lo, n = halves[h]
            blk.propose(kind, prop, surrogate, one, streams=stream_of(k), refresh_pre=swapped and k == 0, chains=(lo, n),
                        step0=base + k)
            flags = one.flags[0, lo:lo + n].cpu().numpy()
            idx = np.nonzero(flags == L.FLAG_PENDING)[0]
            if idx.size == 0:
                inflight[h] = (k, idx, np.empty((0, self.m)))
                return
            props = one.prop[0][:, lo:lo + n].cpu().numpy().T[idx]
            if self.pool is not None:
                if any(isinstance(v[2], tuple) for v in inflight.values()):
                    self.overlap_stats["queued_while_busy"] += 1        # the other half's G is still being evaluated
                inflight[h] = (k, idx, ("pool", self.pool.submit(props)))
                self.no_G_calls_host += props.shape[0]
            else:
                inflight[h] = (k, idx, self._host_G(props))

        def decide(h):
            lo, n = halves[h]
            k, idx, G = inflight.pop(h)
            if isinstance(G, tuple):
                G = self.pool.collect(G[1]).reshape(idx.size, self.m)
            if idx.size:
                full = np.zeros((n, self.m))
                full[idx] = G
                Gd[:, lo:lo + n] = L.dev(np.ascontiguousarray(full.T), device=self.device)
            blk.decide(kind, prop, surrogate, one, Gd, streams=stream_of(k), chains=(lo, n), advance=False, step0=base + k)
            if trace is not None:
                sl = slice(lo, lo + n)
                trace.flags[k, sl] = one.flags[0, sl]
                if trace.prop is not None:
                    trace.prop[k, :, sl], trace.post[k, sl], trace.pre[k, sl] = one.prop[0][:, sl], one.post[0, sl], one.pre[0, sl]
                    trace.pre_cur[k, sl] = one.pre_cur[0, sl]
                if collects:
                    trace.snap_par[k, :, sl], trace.snap_obs[k, :, sl] = one.snap_par[0][:, sl], one.snap_obs[0][:, sl]
            return k

        for h in range(len(halves)):
            propose(h, 0)
        for k in range(K):
            for h in range(len(halves)):
                decide(h)
                if k + 1 < K:
                    propose(h, k + 1)
        blk.step0 = base + K
        if blk.moments is not None:
            blk.moment_steps += K

    def run(self):
        """All stages of samplers_list (process_SAMPLER.py:53-88)."""
        for i, st in enumerate(self.conf.list_alg):
            self.run_stage(i, st)
        return self.results
