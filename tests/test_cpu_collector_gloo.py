"""CPU, world_size 2 over gloo: the collector's snapshot exchange (ONE all-gather of fixed-width per-rank
blocks with the count in a header row, rank-major order, per-rank quota = snapshots_per_refit // world so that
every rank's snapshots enter the refit), the model broadcast of the general path, and the fixed-shape path
(packed state, ring of generations, rotating refit duty, one broadcast from the rank on duty) -- the N>1 logic of
surrdamh_b200/collector.py with the GPU refit replaced by recording stubs (host logic only)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as tmp


class StubModel:
    def __init__(self, coefs, centers=None, degree=0):
        self.coefs, self.centers, self.degree = coefs, centers, degree


class StubUpdater:
    """Keeps every snapshot; 'fits' coefficients that are a checksum of the data so the broadcast is checkable."""

    def __init__(self, d, m, kind):
        self.d, self.m, self.kind = d, m, kind
        self.par = torch.empty((0, d), dtype=torch.float64)
        self.obs = torch.empty((0, m), dtype=torch.float64)

    def add_arrays(self, par, obs):
        self.par = torch.cat([self.par, par])
        self.obs = torch.cat([self.obs, obs])

    def update_device(self):
        n = self.par.shape[0]
        if self.kind == "rbf":
            coefs = torch.arange((n + self.d + 1) * self.m, dtype=torch.float64).reshape(n + self.d + 1, self.m) + self.par.sum()
            return StubModel(coefs, self.par.clone()), n
        coefs = torch.full((6, self.m), float(self.obs.sum()), dtype=torch.float64)
        return StubModel(coefs, None, 2), n


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, kind, mode, cap, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from surrdamh_b200.collector import Collector
    d, m = 3, 2
    upd = StubUpdater(d, m, kind)
    col = Collector(upd, kind, d, m, rank=rank, world=world, snapshots_per_refit=cap, refit_mode=mode, device="cpu")
    col.model_factory = lambda k, coefs, extra: StubModel(coefs, extra if k == "rbf" else None, 0 if k == "rbf" else int(extra))
    res = []
    for rnd, counts in enumerate([(5, 2), (0, 3), (0, 0), (4, 4)]):
        n = counts[rank]
        par = torch.full((8, d), float(100 * rank + rnd), dtype=torch.float64) + torch.arange(8, dtype=torch.float64)[:, None]
        obs = torch.full((8, m), float(-rank - 10 * rnd), dtype=torch.float64)
        did = col.exchange_and_refit(par, obs, torch.tensor([n]), global_step=rnd)
        res.append((did, None if col.model is None else col.model.coefs.clone().numpy(),
                    None if col.model is None or col.model.centers is None else col.model.centers.clone().numpy(),
                    None if col.model is None else col.model.degree))
    out[rank] = (res, col.history, col.no_refits)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("kind,mode,cap", [("rbf", "broadcast", None), ("poly", "broadcast", 6), ("rbf", "replicated", 3)])
def test_collector_exchange_world2(kind, mode, cap):
    world, port = 2, _free_port()
    mgr = tmp.Manager()
    out = mgr.dict()
    tmp.spawn(_worker, args=(world, port, kind, mode, cap, out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    # every rank sees the same history and the same model after every round
    assert r0[1] == r1[1] and r0[2] == r1[2] == 3                   # round 2 had no snapshots anywhere: no refit
    for (d0, c0, z0, g0), (d1, c1, z1, g1) in zip(r0[0], r1[0]):
        assert d0 == d1
        assert (c0 is None) == (c1 is None)
        if c0 is not None:
            assert np.array_equal(c0, c1) and g0 == g1
            assert (z0 is None) == (z1 is None)
            if z0 is not None:
                assert np.array_equal(z0, z1)
    # snapshot counts per refit: rank-major concatenation, every rank capped at its quota = cap // world
    counts = [(5, 2), (0, 3), (4, 4)]
    quota = 99 if cap is None else max(1, cap // world)
    want = [min(a, quota) + min(b, quota) for a, b in counts]
    assert [h[1] for h in r0[1]] == want
    if kind == "rbf" and cap is not None:
        centers = r0[0][0][2]                                        # BOTH ranks' snapshots enter the first refit
        assert centers.shape[0] == want[0] and centers[0, 0] == 0.0 and centers[-1, 0] == 100.0
    if kind == "rbf" and cap is None:
        centers = r0[0][0][2]                                        # after round 0: 5 rows of rank 0 then 2 rows of rank 1
        assert centers.shape == (7, 3)
        assert np.array_equal(centers[:5, 0], np.arange(5.0)) and np.array_equal(centers[5:, 0], 100 + np.arange(2.0))



class FixedStub:
    """A CPU stand-in for the RBF updater's fixed-shape interface: state = the packed buffer; a 'refit' appends the sum of
    the round's alive rows to a running checksum and records who did it."""
    solver_type, round_failures = "direct", 0

    def __init__(self, d, m, no_keep, rank):
        self.d, self.m, self.no_keep, self.rank = d, m, no_keep, rank
        self.done_by_me = []

    def new_pack(self):
        return torch.zeros((self.no_keep * (self.d + self.m + 1) + 32,), dtype=torch.float64)

    def pack_views(self, pack):
        n, d, m = self.no_keep, self.d, self.m
        return (pack[:n].view(n, 1), pack[n:n + n * d].view(n, d), pack[n + n * d:n + n * d + n * m].view(n, m), pack[-16:])

    def pack_from_state(self, model):
        p = self.new_pack()
        self.pack_views(p)[1].copy_(model.centers)
        return p

    def adopt_pack(self, pack):
        coefs, X, Y, _ = self.pack_views(pack)
        return StubModel(coefs, X)

    def update_round_device(self, gathered, world, quota, cur, out):
        g = gathered.view(world, quota + 1, self.d + self.m)
        out.copy_(cur)
        tot = 0.0
        for r in range(world):
            k = int(g[r, 0, 0])
            tot += float(g[r, 1:1 + k].sum()) + 1000.0 * k
        self.pack_views(out)[0][0, 0] += tot
        self.pack_views(out)[3][3] = sum(int(g[r, 0, 0]) for r in range(world))
        self.done_by_me.append(tot)
        return out


def _worker_fixed(rank, world, port, mode, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from surrdamh_b200.collector import Collector
    d, m, n0 = 2, 1, 5
    upd = FixedStub(d, m, n0, rank)
    col = Collector(upd, "rbf", d, m, rank=rank, world=world, snapshots_per_refit=4, refit_mode=mode, device="cpu")
    col.model_factory = lambda k, coefs, extra: StubModel(coefs, extra)
    if rank == 0 or mode == "replicated":
        col.model = StubModel(torch.zeros((n0 + d + 1, m), dtype=torch.float64), torch.arange(n0 * d, dtype=torch.float64).view(n0, d))
    else:
        col.model = StubModel(torch.zeros((n0 + d + 1, m), dtype=torch.float64), torch.zeros((n0, d), dtype=torch.float64))
    sums = []
    for rnd in range(5):
        block = torch.zeros((col.quota + 1, d + m), dtype=torch.float64)
        k = (rnd + rank) % (col.quota + 1)
        block[0, 0] = k
        block[1:1 + k] = 10.0 * rank + rnd + 1.0
        gathered = col.gather(block)
        if not col.fixed and col._can_go_fixed():
            col._enter_fixed()
        assert col.fixed
        col._refit_fixed(gathered, rnd)
        sums.append(float(col.model.coefs[0, 0]))
    out[rank] = (sums, len(upd.done_by_me), col.model.centers.clone().numpy(), [s[4] for s in col.poll_status(wait=True)])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["broadcast", "rotate", "replicated"])
def test_collector_fixed_shape_world2(mode):
    world, port = 2, _free_port()
    mgr = tmp.Manager()
    out = mgr.dict()
    tmp.spawn(_worker_fixed, args=(world, port, mode, out), nprocs=world, join=True)
    (s0, n0, c0, t0), (s1, n1, c1, t1) = out[0], out[1]
    assert s0 == s1                                         # the same surrogate on both ranks after every round
    assert np.array_equal(c0, c1) and np.array_equal(c0.ravel(), np.arange(10.0))     # rank 0's state reached rank 1
    assert all(b > a for a, b in zip(s0, s0[1:]))            # every round built on the previous generation
    assert (n0, n1) == {"broadcast": (5, 0), "rotate": (3, 2), "replicated": (5, 5)}[mode]
    want = [((r) % 3) + ((r + 1) % 3) for r in range(5)]     # snapshots offered per round: both ranks' counts
    assert t0 == want and t1 == want


class PipeStub(FixedStub):
    """CPU stand-in for the staged interface (thin_stage / solve_stage / assemble): the thin stage folds the round's
    snapshots into a running checksum held in X[0, 0]; the solve stage 'fits' COEFS[0, 0] = 2 * that checksum and fails
    on demand."""

    def __init__(self, d, m, no_keep, rank, fail_rounds=()):
        super().__init__(d, m, no_keep, rank)
        self.fail_rounds, self.solved, self.thinned = set(fail_rounds), [], 0

    def new_thin_state(self):
        return (torch.zeros((self.no_keep, self.d), dtype=torch.float64), torch.zeros((self.no_keep, self.m), dtype=torch.float64),
                torch.zeros((4,), dtype=torch.float64))

    def new_result(self):
        return torch.zeros((self.no_keep + 16,), dtype=torch.float64)

    def thin_stage(self, X0, Y0, gathered, world, quota, Xn, Yn, meta):
        g = gathered.view(world, quota + 1, self.d + self.m)
        tot, k_all = 0.0, 0
        for r in range(world):
            k = int(g[r, 0, 0])
            tot += float(g[r, 1:1 + k].sum()) + 1000.0 * k
            k_all += k
        Xn.copy_(X0)
        Yn.copy_(Y0)
        Xn[0, 0] += tot
        meta.zero_()
        meta[0], meta[1] = self.no_keep, k_all
        self.thinned += 1

    def solve_stage(self, X, Y, meta, result):
        idx = len(self.solved)
        result.zero_()
        result[0] = 2.0 * float(X[0, 0])
        result[-16] = 1.0 if self.round_of_next_solve in self.fail_rounds else 0.0
        result[-16 + 3] = meta[1]
        self.solved.append(self.round_of_next_solve)

    def assemble(self, result, X, Y, pack_prev, pack_out):
        coefs, Xp, Yp, status = self.pack_views(pack_out)
        coefs.view(-1).copy_(result[:self.no_keep])
        Xp.copy_(X)
        Yp.copy_(Y)
        if float(result[-16]) != 0.0:
            pack_out.copy_(pack_prev)
        self.pack_views(pack_out)[3].copy_(result[-16:])

    def model_of_pack(self, pack):
        coefs, X, _, _ = self.pack_views(pack)
        return StubModel(coefs, X)


def _worker_pipe(rank, world, port, mode, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from surrdamh_b200.collector import Collector, PipelinedCollector
    d, m, n0 = 2, 1, 5
    upd = PipeStub(d, m, n0, rank, fail_rounds=(3,))
    col = Collector(upd, "rbf", d, m, rank=rank, world=world, snapshots_per_refit=4, refit_mode=mode, device="cpu")
    col.model_factory = lambda k, coefs, extra: StubModel(coefs, extra)
    col.select = lambda trace, K: trace                     # the test hands in ready-made blocks
    start = torch.arange(n0 * d, dtype=torch.float64).view(n0, d) if (rank == 0 or mode == "replicated") else torch.zeros((n0, d), dtype=torch.float64)
    col.model = StubModel(torch.zeros((n0 + d + 1, m), dtype=torch.float64), start)
    pc = PipelinedCollector(col, "cpu", lag=2)
    seen = []
    for rnd in range(6):
        block = torch.zeros((col.quota + 1, d + m), dtype=torch.float64)
        k = (rnd + rank) % (col.quota + 1)
        block[0, 0] = k
        block[1:1 + k] = 10.0 * rank + rnd + 1.0
        upd.round_of_next_solve = rnd
        pc.submit(rnd, block, 1, None, rnd)
        ev, model = pc.result(rnd)
        seen.append((float(model.coefs[0, 0]), float(model.centers[0, 0]), float(model.centers[1, 1])))
    pc.close()
    out[rank] = (seen, upd.solved, upd.thinned, float(col.model.coefs[0, 0]), [s[1] for s in col.poll_status(wait=True)])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["broadcast", "rotate", "replicated"])
def test_pipelined_collector_world2(mode):
    """PipelinedCollector over gloo: every rank runs the thin stage of every round, the solve runs on the rank on duty
    only, ONE broadcast per round hands [COEFS | status] to the others, a failed solve (round 3) leaves the previous
    round's model in place on every rank while the thinned state keeps evolving."""
    world, port = 2, _free_port()
    mgr = tmp.Manager()
    out = mgr.dict()
    tmp.spawn(_worker_pipe, args=(world, port, mode, out), nprocs=world, join=True)
    (s0, solved0, th0, last0, st0), (s1, solved1, th1, last1, st1) = out[0], out[1]
    assert s0 == s1 and last0 == last1                       # identical models on both ranks after every round
    assert th0 == th1 == 6                                   # the thin stage is replicated
    assert (solved0, solved1) == {"broadcast": ([0, 1, 2, 3, 4, 5], []), "rotate": ([0, 2, 4], [1, 3, 5]),
                                  "replicated": ([0, 1, 2, 3, 4, 5], [0, 1, 2, 3, 4, 5])}[mode]
    coefs = [c for c, _, _ in s0]
    assert all(c == 2.0 * x for (c, x, _) in s0)             # COEFS and centers of a model belong to the same round ...
    assert coefs[3] == coefs[2] and coefs[4] > coefs[3]      # ... the failed round 3 shows round 2's model, round 4 moved on
    assert all(y == 3.0 for _, _, y in s0)                   # rank 0's start state reached every rank (centers[1, 1] = 3)
    assert st0 == st1 == [0, 0, 0, 1, 0, 0]                  # the status log names the failed round
