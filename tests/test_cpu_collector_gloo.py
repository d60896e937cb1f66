"""CPU, world_size 2 over gloo: the collector's snapshot exchange (all-gather of ragged per-rank
blocks in rank-major order, per-refit cap) and the coefficient broadcast -- the N>1 path of
surrdamh_b200/collector.py with the GPU refit replaced by a recording stub (host logic only)."""
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
    upd = StubUpdater(d, m, kind) if (rank == 0 or mode == "replicated") else None
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
    # snapshot counts per refit: rank-major concatenation, capped
    counts = [(5, 2), (0, 3), (4, 4)]
    want = [min(a, cap or 99) + min(b, cap or 99) for a, b in counts]
    if cap is not None:
        want = [min(w, cap) for w in want]
    assert [h[1] for h in r0[1]] == want
    if kind == "rbf" and cap is None:
        centers = r0[0][0][2]                                        # after round 0: 5 rows of rank 0 then 2 rows of rank 1
        assert centers.shape == (7, 3)
        assert np.array_equal(centers[:5, 0], np.arange(5.0)) and np.array_equal(centers[5:, 0], 100 + np.arange(2.0))
