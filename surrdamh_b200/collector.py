"""The collector's job for lockstep chains: gather the snapshots every rank produced since the last
refit, refit the surrogate, hand the new model to every rank.

Counterpart of surrDAMH/process_COLLECTOR.py:57-113 (whenever snapshots arrived: add_data + a FULL
update(), then push SOL to the samplers) and of the sampler-side batching / double buffer in
surrDAMH/modules/classes_communication.py:121-150.  Transport: the reference's pickled MPI
isend/irecv becomes, with one process per GPU, per round
    ONE all_gather of fixed-width snapshot blocks (count in a header row)   NCCL (gloo in the CPU tests)
    ONE broadcast of the packed state [COEFS | centers | observations | status] from the rank on refit duty
or no broadcast at all when every rank refits the identical gathered set itself ('replicated').

Two refit paths:
  fixed-shape (RBF, `snapshots_per_refit` set, updater holding exactly `no_keep` rows, dense solver): every buffer has
      a static size, nothing synchronises the host -- the round is a chain of kernel launches and two NCCL calls that
      can be enqueued on a side stream while the chains keep stepping (AsyncCollector).  The refit duty can rotate
      over the ranks (`refit_mode: "rotate"`), because the packed state makes every rank a full copy of the updater.
  general (everything else: the reference's unbounded snapshot stream, growing center sets, poly, MINRES): sizes are
      read back to the host, the model is broadcast from rank 0.

Which snapshots enter a refit when `snapshots_per_refit` = q is set (additive key; the reference forwards every
snapshot, which is what happens when it is unset): every rank contributes floor(q / world) (at least 1) of its valid
snapshots, evenly spaced over their (step, chain) order within the block; the gathered set is ordered (rank, step,
chain).  The reference's timing is a race between busy-polling ranks; here it is a schedule: the sampler calls
`round()` at fixed step counts, so a run is reproducible.
"""
import numpy as np
import torch

from . import _lib as L
from .device import DeviceModel


class Collector:
    def __init__(self, updater, surrogate_type, no_parameters, no_observations, kernel_type=0, group=None, rank=0, world=1,
                 snapshots_per_refit=None, refit_mode="broadcast", device=None):
        self.updater = updater                  # Surrogate_update of this package
        self.surrogate_type = surrogate_type    # 'poly' | 'rbf'
        self.d, self.m = no_parameters, no_observations
        self.kernel_type = kernel_type
        self.group, self.rank, self.world = group, rank, world
        self.cap = None if snapshots_per_refit is None else int(snapshots_per_refit)
        self.quota = None if self.cap is None else max(1, self.cap // world)
        self.refit_mode = refit_mode            # 'broadcast' (rank 0 refits) | 'rotate' | 'replicated'
        self.device = torch.device(device or "cuda")
        self.model = None                       # the surrogate the chains currently evaluate
        self.no_refits = 0
        self.no_snapshots = 0
        self.history = []                       # (global_step, snapshots offered to the refit | None when only the device knows)
        self.model_factory = self._device_model   # (kind, coefs, centers | degree) -> model; tests stub it on CPU
        self.refit_hook = None                    # optional (par_all, obs_all, refit_index) -> SOL: replaces the fit (parity runs)
        # fixed-shape rounds
        self.fixed = False                      # True once every rank holds the packed state
        self.ring, self.gen = [], 0             # packed states of the last generations (the chains may still read g-1, g-2)
        self._status_pending = []               # (event, pinned status copy, round index)
        self.status_log = []                    # host copies of the rounds' status rows, read one round late
        self.allow_fixed = True

    def _device_model(self, kind, coefs, extra):
        if kind == "rbf":
            return DeviceModel.rbf(coefs, extra, self.kernel_type, device=self.device)
        return DeviceModel.poly(coefs, int(extra), self.d, device=self.device)

    # ---- snapshot exchange ---------------------------------------------------------------------
    def select(self, trace, K):
        """This rank's block of the round: [(quota + 1) x (d + m)], header row = (rows kept, valid snapshots)."""
        C_ = trace.C
        quota = self.quota if self.quota is not None else K * C_
        block = torch.empty((quota + 1, self.d + self.m), dtype=torch.float64, device=trace.flags.device)
        scratch = torch.empty(((K * C_ + 1023) // 1024 + 2,), dtype=torch.int64, device=trace.flags.device)
        L.check(L.lib().sdb_snapshot_select(L.ptr(trace.flags), L.ptr(trace.snap_par), L.ptr(trace.snap_obs), K, C_, self.d, self.m,
                                            L.ptr(block), quota, L.ptr(scratch), L.stream_handle()))
        return block

    def gather(self, block, group=None):
        """[world x (quota + 1) x (d + m)], identical on every rank: the ONE collective of the exchange."""
        if self.world == 1:
            return block.unsqueeze(0)
        import torch.distributed as dist
        out = torch.empty((self.world * block.shape[0], block.shape[1]), dtype=block.dtype, device=block.device)
        dist.all_gather_into_tensor(out, block, group=group if group is not None else self.group)
        return out.view(self.world, block.shape[0], block.shape[1])

    @staticmethod
    def rows_of(gathered, d):
        """Host-synchronising: the valid rows of a gathered set in (rank, step, chain) order -> (par, obs)."""
        counts = gathered[:, 0, 0].to(torch.int64).cpu().numpy()
        rows = torch.cat([gathered[r, 1:1 + int(counts[r])] for r in range(gathered.shape[0])], dim=0)
        return rows[:, :d].contiguous(), rows[:, d:].contiguous()

    # ---- general path: sizes on the host, model from rank 0 -------------------------------------------------------
    def _broadcast_model(self, model):
        import torch.distributed as dist
        if self.surrogate_type == "rbf":
            meta = torch.tensor([model.centers.shape[0] if model is not None else 0, 0], dtype=torch.int64, device=self.device)
        else:
            meta = torch.tensor([model.coefs.shape[0] if model is not None else 0, model.degree if model is not None else 0],
                                dtype=torch.int64, device=self.device)
        dist.broadcast(meta, src=0, group=self.group)
        n0, deg = [int(v) for v in meta.cpu().numpy()]
        if self.surrogate_type == "rbf":
            coefs = model.coefs if self.rank == 0 else torch.empty((n0 + self.d + 1, self.m), dtype=torch.float64, device=self.device)
            centers = model.centers if self.rank == 0 else torch.empty((n0, self.d), dtype=torch.float64, device=self.device)
            dist.broadcast(coefs, src=0, group=self.group)
            dist.broadcast(centers, src=0, group=self.group)
            return model if self.rank == 0 else self.model_factory("rbf", coefs, centers)
        coefs = model.coefs if self.rank == 0 else torch.empty((n0, self.m), dtype=torch.float64, device=self.device)
        dist.broadcast(coefs, src=0, group=self.group)
        return model if self.rank == 0 else self.model_factory("poly", coefs, deg)

    def exchange_and_refit(self, par, obs, count=None, global_step=0):
        """General path on explicit arrays: par [n x d], obs [n x m] are THIS rank's snapshots of the round (valid rows
        first; count: device int64[1] or int, default all rows).  Returns True when a new surrogate is in place."""
        n_local = par.shape[0] if count is None else (int(count.item()) if isinstance(count, torch.Tensor) else int(count))
        n_local = min(n_local, par.shape[0])
        quota = self.quota if self.quota is not None else n_local
        if self.world > 1 and self.quota is None:
            import torch.distributed as dist
            mx = torch.tensor([n_local], dtype=torch.int64, device=par.device)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=self.group)
            quota = int(mx.item())
        n_local = min(n_local, quota)
        block = torch.zeros((max(quota, 1) + 1, self.d + self.m), dtype=torch.float64, device=par.device)
        block[0, 0] = n_local
        block[1:1 + n_local, :self.d] = par[:n_local]
        block[1:1 + n_local, self.d:] = obs[:n_local]
        return self._refit_general(self.gather(block), global_step)

    def _refit_general(self, gathered, global_step):
        par_all, obs_all = self.rows_of(gathered, self.d)
        if par_all.shape[0] == 0:
            return False
        if self.refit_hook is not None:
            SOL = self.refit_hook(par_all, obs_all, self.no_refits)
            self.model = self.model_factory(self.surrogate_type, SOL[0], SOL[1])
        else:
            model = None
            if self.world == 1 or self.refit_mode == "replicated" or self.rank == 0:
                self.updater.add_arrays(par_all, obs_all)
                model, n = self.updater.update_device()
                self.no_snapshots = n
            if self.world > 1 and self.refit_mode != "replicated":
                model = self._broadcast_model(model)
            self.model = model
        self.no_refits += 1
        self.history.append((global_step, int(par_all.shape[0])))
        return True

    # ---- fixed-shape path ------------------------------------------------------------------------------------------
    def _can_go_fixed(self):
        """Identical on every rank: decided from the model every rank holds, not from rank 0's updater."""
        up = self.updater
        return (self.allow_fixed and self.refit_hook is None and self.surrogate_type == "rbf" and self.cap is not None and up is not None
                and getattr(up, "no_keep", None) is not None and getattr(up, "solver_type", "") in ("direct", "nullspace")
                and self.model is not None and self.model.centers is not None
                and int(self.model.centers.shape[0]) == int(up.no_keep) and hasattr(up, "update_round_device"))

    def _enter_fixed(self):
        """One-off: pack rank 0's (or every replica's) state, hand it to all ranks."""
        up = self.updater
        self.ring = [up.new_pack() for _ in range(4)]
        pack = self.ring[0]
        if self.world == 1 or self.rank == 0 or self.refit_mode == "replicated":
            pack.copy_(up.pack_from_state(self.model))
        if self.world > 1 and self.refit_mode != "replicated":
            import torch.distributed as dist
            dist.broadcast(pack, src=0, group=self.group)
        self.model = up.adopt_pack(pack)
        self.gen, self.fixed = 0, True

    def duty_rank(self, round_index):
        return round_index % self.world if self.refit_mode == "rotate" else 0

    def _refit_fixed(self, gathered, global_step):
        up = self.updater
        cur = self.ring[self.gen % len(self.ring)]
        self.gen += 1
        out = self.ring[self.gen % len(self.ring)]
        duty = self.duty_rank(self.no_refits)
        if self.world == 1 or self.refit_mode == "replicated" or self.rank == duty:
            up.update_round_device(gathered.reshape(-1), self.world, self.quota, cur, out)
        if self.world > 1 and self.refit_mode != "replicated":
            import torch.distributed as dist
            dist.broadcast(out, src=duty, group=self.group)
        self.model = up.adopt_pack(out)
        self.no_snapshots = int(up.no_keep)
        self._queue_status(up.pack_views(out)[3])
        self.no_refits += 1
        self.history.append((global_step, None))
        return True

    def _queue_status(self, status):
        """The round's status row goes to pinned memory asynchronously (current stream) and is looked at one round later."""
        if status.is_cuda:
            pinned = torch.empty((16,), dtype=torch.float64).pin_memory()
            pinned.copy_(status, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            self._status_pending.append((ev, pinned, self.no_refits))
        else:
            self._status_pending.append((None, status.clone(), self.no_refits))
        self.poll_status()

    def poll_status(self, wait=False):
        """Move finished rounds' status rows to `status_log` [(round, failed, solver info, rows kept, new snapshots)]."""
        keep = []
        for ev, pinned, idx in self._status_pending:
            if wait and ev is not None:
                ev.synchronize()
            if ev is None or ev.query():
                s = pinned.numpy()
                self.status_log.append((idx, int(s[0]), int(s[1]), int(s[2]), int(s[3])))
                if s[0] != 0 and self.updater is not None:
                    self.updater.round_failures += 1
            else:
                keep.append((ev, pinned, idx))
        self._status_pending = keep
        return self.status_log

    # ---- one collector round -----------------------------------------------------------------------------------------
    def round(self, trace, K, global_step=0):
        """Snapshots of the launch in `trace` -> all ranks -> refit -> new model on every rank.  Returns True when a
        new surrogate is in place (always, on the fixed-shape path: whether the round had any snapshot is known to the
        device only; a round without one reproduces the previous coefficients)."""
        block = self.select(trace, K)
        if self.world > 1 and self.quota is None:
            # unbounded (the reference's semantics): blocks are as wide as the busiest rank's count
            par, obs = block[1:, :self.d], block[1:, self.d:]
            return self.exchange_and_refit(par.contiguous(), obs.contiguous(), block[0, 0].to(torch.int64).reshape(1), global_step)
        gathered = self.gather(block)
        if not self.fixed and self._can_go_fixed():
            self._enter_fixed()
        if self.fixed:
            return self._refit_fixed(gathered, global_step)
        return self._refit_general(gathered, global_step)

    def set_model_from_sol(self, SOL):
        """Start from a given surrogate (reference-shaped SOL list)."""
        self.model = self.model_factory(self.surrogate_type, SOL[0], SOL[1])
        self.fixed = False


def small_message_group(max_ctas=2):
    """A process group of all ranks for the collector's KB-sized collectives.  NCCL: at most `max_ctas` CTAs per
    collective -- the default (up to 32 channels, one CTA each) does not fit on the few SMs the chain kernel leaves to the
    asynchronous collector, and a collective whose CTAs cannot all start waits for the chain block to end (measured on 8
    GPUs: a 14 ms all-gather of 800 bytes).  Collective call: every rank must make it, in the same order."""
    import torch.distributed as dist
    if dist.get_backend() == "nccl":
        try:
            opts = dist.ProcessGroupNCCL.Options()
            opts.config.max_ctas = int(max_ctas)
            opts.config.min_ctas = 1
            return dist.new_group(pg_options=opts)
        except (AttributeError, TypeError, RuntimeError):
            pass
    return dist.new_group()


class _Lane:
    """A CUDA stream of its own -- or, on the CPU (the gloo tests drive the same code), nothing at all."""

    def __init__(self, device, priority=0, stream=None):
        self.cuda = torch.device(device).type == "cuda"
        self.stream = (stream if stream is not None else torch.cuda.Stream(device=device, priority=priority)) if self.cuda else None

    def __enter__(self):
        if self.cuda:
            self._ctx = torch.cuda.stream(self.stream)
            self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.cuda:
            self._ctx.__exit__(*exc)
        return False

    def wait(self, event):
        if self.cuda and event is not None:
            self.stream.wait_event(event)

    def record(self, timing=False):
        if not self.cuda:
            return None
        ev = torch.cuda.Event(enable_timing=timing)
        ev.record(self.stream)
        return ev


class AsyncCollector:
    """The collector as a concurrent actor, like the reference's separate COLLECTOR process
    (process_COLLECTOR.py runs beside the samplers, which keep stepping with the surrogate they have until a
    new one arrives -- classes_communication.py:137-150).  Here the asynchrony is deterministic: the snapshots
    of block b are refitted on a side CUDA stream WHILE the next `lag` blocks are being stepped, and the new surrogate
    is switched in at the start of block b + lag + 1, so a run is reproducible.

    On the fixed-shape path a round never synchronises the host, so it is simply enqueued on the side stream from the
    sampling thread; the sampling stream waits on the round's completion event when it needs that round's surrogate.
    (A general-path round -- the first rounds of a run, while the center set still grows -- blocks the host like the
    synchronous collector does.)
    """

    def __init__(self, collector, device, stream=None, lag=1, make_group=True):
        self.col = collector
        self.device = torch.device(device)
        # high priority: the round's critical-path kernels get the next free SM ahead of its own bulk updates
        self.lane = _Lane(self.device, priority=-1, stream=stream)
        self.stream = self.lane.stream
        self.lag = int(lag)
        self.results = {}
        self.round_events = []                # (start_event, end_event) pairs of the timed rounds
        if collector.world > 1 and collector.group is None and make_group:
            # collective: every rank constructs its asynchronous collector.  The round's collectives must fit on the SMs
            # the chain kernel leaves free.
            collector.group = small_message_group()

    def submit(self, b, trace, K, chain_done_event, step, timed=False):
        with self.lane:
            self.lane.wait(chain_done_event)
            r0 = self.lane.record(timing=True) if timed else None
            did = self.col.round(trace, K, step)
            r1 = self.lane.record(timing=timed)
            if timed:
                self.round_events.append((r0, r1))
        self.results[b] = (r1, self.col.model if did else None)

    def result(self, b):
        """(completion event, new surrogate or None) of round b."""
        return self.results[b]

    def forget(self, b):
        self.results.pop(b, None)

    def close(self):
        self.results.clear()


class PipelinedCollector(AsyncCollector):
    """The asynchronous collector with a round split into its two stages, each on a stream of its own:

      thin   select + ONE all-gather + candidates + greedy thinning + compaction.  Strictly sequential from round to round
             (round b+1 thins the centers round b kept) but cheap -- one SM for about a millisecond -- so EVERY rank runs
             it: the center sets stay identical on all ranks without being sent.
      solve  the dense solve of the thinned set.  It depends on the thin stage of ITS round only, so round b+1 is thinned
             while round b is still being solved, and with `refit_mode: "rotate"` the solves of consecutive rounds run on
             different ranks at the same time: rank b mod N solves round b and broadcasts [COEFS | status] (ONE
             broadcast, on a communicator of its own so that it never queues behind an all-gather).

    A third stream receives the broadcast and assembles the model of the round, [COEFS | centers | observations | status]
    -- or a copy of the previous round's model when the solve failed.  The sampling stream switches to the surrogate
    of round b at the start of block b + lag + 1, as with AsyncCollector; with N ranks the solve of a round may take N
    block periods, which is what makes a STRONG-scaling run (few chains per GPU, short blocks) keep its refit rate: the
    reference's collector is likewise a separate process the samplers never wait for (process_COLLECTOR.py:57-113).

    Rounds before the collector reaches its fixed shapes (center set still growing) run serially like AsyncCollector's.
    """

    def __init__(self, collector, device, lag=1, make_group=True):
        super().__init__(collector, device, lag=lag, make_group=make_group)
        col = collector
        self.thin = self.lane                           # the thin stage runs where AsyncCollector's serial rounds run
        self.solve = _Lane(device, priority=0)
        self.asm = _Lane(device, priority=-1)
        self.cuda = self.thin.cuda
        # communicators of its own (collective: every rank constructs its PipelinedCollector): the broadcast of round b
        # must not queue behind the all-gather of round b+1, and both have to fit beside the chain kernel
        self.ggroup = self.bgroup = col.group
        if col.world > 1 and make_group and col.refit_mode != "replicated":
            if getattr(col, "_bgroup", None) is None:
                col._bgroup = small_message_group()
            self.bgroup = col._bgroup
        self.nring = self.lag + 4
        self.slots = None
        self.n_piped = 0
        self.stage_events = []                          # timed rounds: (start, thinned, solved | None, assembled, gathered)
        self.last = None                                # slot of the newest round

    # ---- state ---------------------------------------------------------------------------------------------------
    def _start(self):
        """Enter the pipeline from the collector's current packed state (identical on every rank)."""
        col, up = self.col, self.col.updater
        cur = col.ring[col.gen % len(col.ring)]
        self.slots = []
        for _ in range(self.nring):
            X, Y, meta = up.new_thin_state()
            self.slots.append(dict(X=X, Y=Y, meta=meta, result=up.new_result(), pack=up.new_pack()))
        s0 = self.slots[-1]                              # "round -1": the state the first pipelined round starts from
        _, X0, Y0, _ = up.pack_views(cur)
        s0["X"].copy_(X0)
        s0["Y"].copy_(Y0)
        s0["pack"].copy_(cur)
        self.last = self.nring - 1
        if self.cuda:
            # Prime the solve path on EVERY rank now (scratch allocation, module loads, CUDA-graph capture of the solver): a
            # rank whose first solve duty came in the middle of a run stalled for ~200 ms (measured, 4 GPUs), and with it every
            # other rank at the next all-gather.
            ev = self.thin.record()
            with self.solve:
                self.solve.wait(ev)
                _, _, meta0 = up.new_thin_state()        # buffers of their own: slot 0 belongs to the first round already
                self._prime = (meta0, up.new_result())
                meta0[0] = float(int(up.no_keep))
                for _ in range(2):
                    up.solve_stage(s0["X"], s0["Y"], meta0, self._prime[1])

    def submit(self, b, trace, K, chain_done_event, step, timed=False):
        col, up = self.col, self.col.updater
        if not col.fixed:
            if not (col.model is not None and col._can_go_fixed()):
                return super().submit(b, trace, K, chain_done_event, step, timed)     # general-path round, serial
            with self.thin:
                col._enter_fixed()
        with self.thin:
            self.thin.wait(chain_done_event)
            if self.slots is None:
                self._start()
            prev, cur = self.slots[self.last], self.slots[(self.last + 1) % self.nring]
            t0 = self.thin.record(timing=True) if timed else None
            block = col.select(trace, K)
            gathered = col.gather(block, self.ggroup)
            ev_g = self.thin.record(timing=True) if timed else None
            up.thin_stage(prev["X"], prev["Y"], gathered.reshape(-1), col.world, col.quota, cur["X"], cur["Y"], cur["meta"])
            ev_t = self.thin.record(timing=timed)
        duty = col.duty_rank(col.no_refits)
        everyone = col.world == 1 or col.refit_mode == "replicated"
        ev_s = None
        if everyone or col.rank == duty:
            with self.solve:
                self.solve.wait(ev_t)
                up.solve_stage(cur["X"], cur["Y"], cur["meta"], cur["result"])
                ev_s = self.solve.record(timing=timed)
        with self.asm:
            self.asm.wait(ev_t)
            self.asm.wait(ev_s)
            if not everyone:
                import torch.distributed as dist
                dist.broadcast(cur["result"], src=duty, group=self.bgroup)
            up.assemble(cur["result"], cur["X"], cur["Y"], prev["pack"], cur["pack"])
            col._queue_status(up.pack_views(cur["pack"])[3])
            ev_m = self.asm.record(timing=timed)
        if timed:
            self.stage_events.append((t0, ev_t, ev_s, ev_m, ev_g))
            self.round_events.append((t0, ev_m))
        self.last = (self.last + 1) % self.nring
        self.n_piped += 1
        col.no_refits += 1
        col.history.append((step, None))
        self.results[b] = (ev_m, up.model_of_pack(cur["pack"]))

    def close(self):
        """End of a stage: the collector and its updater take over the newest round's state (call after the sampling
        stream has waited for every round)."""
        if self.slots is not None:
            col, up = self.col, self.col.updater
            cur = self.slots[self.last]
            if self.cuda:
                torch.cuda.current_stream().wait_stream(self.asm.stream)
            nxt = col.ring[(col.gen + 1) % len(col.ring)]
            nxt.copy_(cur["pack"])
            col.gen += 1
            col.model = up.adopt_pack(nxt)
            col.no_snapshots = int(up.no_keep)
            self.slots = None
        super().close()
