"""The collector's job for lockstep chains: gather the snapshots every rank produced since the last
refit, refit the surrogate, hand the new model to every rank.

Counterpart of surrDAMH/process_COLLECTOR.py:57-113 (whenever snapshots arrived: add_data + a FULL
update(), then push SOL to the samplers) and of the sampler-side batching / double buffer in
surrDAMH/modules/classes_communication.py:121-150.  Transport: the reference's pickled MPI
isend/irecv becomes, with one process per GPU,
    all_gather(counts) + all_gather(fixed-size snapshot blocks)   over NCCL (gloo in the CPU tests)
    broadcast(sizes) + broadcast(coefficients [, centers])        from rank 0  (refit_mode 'broadcast')
or no broadcast at all when every rank refits the identical gathered set itself ('replicated').
The reference's timing is a race between busy-polling ranks; here it is a schedule: the sampler
calls `exchange_and_refit` at fixed step counts, so a run is reproducible.
"""
import numpy as np
import torch

from . import _lib as L
from .device import DeviceModel


class Collector:
    def __init__(self, updater, surrogate_type, no_parameters, no_observations, kernel_type=0, group=None, rank=0, world=1,
                 snapshots_per_refit=None, refit_mode="broadcast", device=None):
        self.updater = updater                  # Surrogate_update of this package (owned by rank 0 / every rank)
        self.surrogate_type = surrogate_type    # 'poly' | 'rbf'
        self.d, self.m = no_parameters, no_observations
        self.kernel_type = kernel_type
        self.group, self.rank, self.world = group, rank, world
        self.cap = None if snapshots_per_refit is None else int(snapshots_per_refit)
        self.refit_mode = refit_mode
        self.device = torch.device(device or "cuda")
        self.model = None                       # the surrogate the chains currently evaluate
        self.no_refits = 0
        self.no_snapshots = 0
        self.history = []                       # (global_step, snapshots used by the refit)
        self.refit_ms = []
        self.model_factory = self._device_model   # (kind, coefs, centers | degree) -> model; tests stub it on CPU
        self.refit_hook = None                    # optional (par_all, obs_all, refit_index) -> SOL: replaces the fit (parity runs)

    def _device_model(self, kind, coefs, extra):
        if kind == "rbf":
            return DeviceModel.rbf(coefs, extra, self.kernel_type, device=self.device)
        return DeviceModel.poly(coefs, int(extra), self.d, device=self.device)

    # ---- snapshot exchange ---------------------------------------------------------------------
    def _gather(self, par, obs, count):
        """par [cap_local x d], obs [cap_local x m] (valid rows first), count: device int64[1] or int.
        -> (par_all [n x d], obs_all [n x m]) identical on every rank, rank-major order."""
        n_local = int(count.item()) if isinstance(count, torch.Tensor) else int(count)
        n_local = min(n_local, par.shape[0])
        if self.cap is not None:
            n_local = min(n_local, self.cap)
        if self.world == 1:
            return par[:n_local], obs[:n_local]
        import torch.distributed as dist
        counts = torch.zeros((self.world,), dtype=torch.int64, device=par.device)
        mine = torch.tensor([n_local], dtype=torch.int64, device=par.device)
        dist.all_gather_into_tensor(counts, mine, group=self.group)
        counts_h = counts.cpu().numpy()
        width = int(counts_h.max())
        if width == 0:
            return par[:0], obs[:0]
        block = torch.zeros((width, self.d + self.m), dtype=torch.float64, device=par.device)
        block[:n_local, :self.d] = par[:n_local]
        block[:n_local, self.d:] = obs[:n_local]
        gathered = torch.empty((self.world * width, self.d + self.m), dtype=torch.float64, device=par.device)
        dist.all_gather_into_tensor(gathered, block, group=self.group)
        rows = torch.cat([gathered[r * width:r * width + int(counts_h[r])] for r in range(self.world)], dim=0)
        if self.cap is not None:
            rows = rows[:self.cap]               # the first `cap` snapshots in (rank, step, chain) order
        return rows[:, :self.d].contiguous(), rows[:, self.d:].contiguous()

    # ---- refit + distribution --------------------------------------------------------------------
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

    def exchange_and_refit(self, par, obs, count, global_step=0, timed=False):
        """One collector round.  Returns True when a new surrogate is in place."""
        par_all, obs_all = self._gather(par, obs, count)
        if par_all.shape[0] == 0:
            return False
        ev0 = ev1 = None
        if timed and torch.cuda.is_available():
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        model = None
        if self.refit_hook is not None:
            SOL = self.refit_hook(par_all, obs_all, self.no_refits)
            self.model = self.model_factory(self.surrogate_type, SOL[0], SOL[1])
            self.no_refits += 1
            self.history.append((global_step, int(par_all.shape[0])))
            return True
        if self.world == 1 or self.refit_mode == "replicated" or self.rank == 0:
            self.updater.add_arrays(par_all, obs_all)
            model, n = self.updater.update_device()
            self.no_snapshots = n
        if self.world > 1 and self.refit_mode != "replicated":
            model = self._broadcast_model(model)
        if ev0 is not None:
            ev1.record()
            ev1.synchronize()
            self.refit_ms.append(ev0.elapsed_time(ev1))
        self.model = model
        self.no_refits += 1
        self.history.append((global_step, int(par_all.shape[0])))
        return True

    def set_model_from_sol(self, SOL):
        """Start from a given surrogate (reference-shaped SOL list)."""
        self.model = self.model_factory(self.surrogate_type, SOL[0], SOL[1])


class AsyncCollector:
    """The collector as a concurrent actor, like the reference's separate COLLECTOR process
    (process_COLLECTOR.py runs beside the samplers, which keep stepping with the surrogate they have until a
    new one arrives -- classes_communication.py:137-150).  Here the asynchrony is deterministic: the snapshots
    of block b are refitted on a side CUDA stream WHILE block b+1 is being stepped, and the new surrogate is
    switched in at the start of block b+2 (a fixed lag of one block), so a run is reproducible.

    A host thread drives the side stream, because one collector round contains host synchronisations (snapshot
    counts, solver status); the sampling thread never blocks on them -- it only makes its stream wait on the
    round's completion event when it needs that round's surrogate.
    """

    def __init__(self, collector, device, stream=None):
        import queue
        import threading
        self.col = collector
        self.device = torch.device(device)
        self.stream = stream if stream is not None else torch.cuda.Stream(device=self.device)
        self.jobs = queue.Queue()
        self.results = {}
        self.cv = threading.Condition()
        self.error = None
        self.round_ms = []                    # (start_event, end_event) pairs of the timed rounds
        self.thread = threading.Thread(target=self._work, daemon=True)
        self.thread.start()

    def _work(self):
        from .chains import compact_snapshots
        torch.cuda.set_device(self.device)
        while True:
            job = self.jobs.get()
            if job is None:
                return
            b, trace, K, cap, chain_done, step, timed = job
            try:
                with torch.cuda.stream(self.stream):
                    self.stream.wait_event(chain_done)
                    r0 = torch.cuda.Event(enable_timing=True)
                    r1 = torch.cuda.Event(enable_timing=True)
                    r0.record()
                    par, obs, cnt = compact_snapshots(trace, K, cap)
                    did = self.col.exchange_and_refit(par, obs, cnt, step)
                    r1.record()
                    if timed:
                        self.round_ms.append((r0, r1))
                    res = (r1, self.col.model if did else None)
            except BaseException as e:      # surfaced in the sampling thread by result()
                self.error = e
                res = (None, None)
            with self.cv:
                self.results[b] = res
                self.cv.notify_all()

    def submit(self, b, trace, K, cap, chain_done_event, step, timed=False):
        self.jobs.put((b, trace, K, cap, chain_done_event, step, timed))

    def result(self, b):
        """(completion event, new surrogate or None) of round b; waits (host side) until the round has been issued."""
        with self.cv:
            while b not in self.results:
                self.cv.wait(timeout=0.05)
                if self.error is not None:
                    raise self.error
            if self.error is not None:
                raise self.error
            return self.results[b]

    def forget(self, b):
        with self.cv:
            self.results.pop(b, None)

    def close(self):
        self.jobs.put(None)
        self.thread.join(timeout=10)
