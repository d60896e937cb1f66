"""Every kernel of the library once, at small sizes (for compute-sanitizer runs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from surrdamh_b200 import _lib as L
from surrdamh_b200 import surrogate_poly as SP, surrogate_rbf as SR
from surrdamh_b200.chains import ChainBlock, Proposal, Trace, compact_snapshots
from surrdamh_b200.device import DeviceModel, DeviceProblem

rs = np.random.RandomState(0)
kw = dict(prior_mean=[5.0, 3.0], prior_std=[[4, -2], [-2, 4]], noise_std=2e-4, no_observations=1, observations=[-1e-3])
with torch.cuda.stream(torch.cuda.Stream()):
    # poly fit + eval
    X = rs.standard_normal((300, 2)) * 2 + np.array([5.0, 3.0]); Y = (-0.0375 * np.exp(-X[:, 0]) - 0.0125 * np.exp(-X[:, 1])).reshape(-1, 1)
    up = SP.Surrogate_update(2, 1, max_degree=5); up.add_arrays(X, Y); SOLp, _ = up.update()
    SP.Surrogate_apply(2, 1).apply(SOLp, X[:37])
    # rbf fits: LU, nullspace (twice: graph replay), minres, thinning
    for st in ("direct", "nullspace", "minres"):
        u = SR.Surrogate_update(2, 1, kernel_type=1, solver_type=st, no_keep=330, max_iterations=50)
        u.add_arrays(X, Y); u.update()
        u.add_arrays(X[:60] + 0.01, Y[:60]); SOLr, n = u.update()
        u.add_arrays(X[:60] + 0.02, Y[:60]); SOLr, n = u.update()
    SR.Surrogate_apply(2, 1, kernel_type=1).apply(SOLr, X[:41])
    Xb = rs.standard_normal((700, 4)); Yb = np.sin(Xb @ rs.standard_normal((4, 3)))
    for st in ("direct", "nullspace"):
        u = SR.Surrogate_update(4, 3, kernel_type=1, solver_type=st); u.add_arrays(Xb, Yb); SOL4, _ = u.update()
    SR.Surrogate_apply(4, 3, kernel_type=1).apply(SOL4, Xb[:100])
    # chains: MH, DAMH poly, DAMH rbf (L > 1), split phases, snapshots, compaction, philox dump
    prob = DeviceProblem(2, **kw)
    for C_, sur in ((70, DeviceModel.poly(SOLp[0], SOLp[1], 2)), (33, DeviceModel.rbf(SOLr[0], SOLr[1], 1))):
        blk = ChainBlock(prob, C_, seed=1, moments=True)
        prop = Proposal(2, [0.6, 0.9])
        tr = Trace(20, C_, 2, 1, blk.device, snapshots=True)
        blk.prepare("MH", None, DeviceModel.toy()); blk.run("MH", 20, prop, None, DeviceModel.toy(), trace=tr)
        compact_snapshots(tr, 20, 100)
        blk.prepare("DAMH", sur, DeviceModel.toy()); blk.run("DAMH", 20, prop, sur, DeviceModel.toy(), trace=tr, refresh_pre=True)
        compact_snapshots(tr, 20, 20 * C_)
        blk.philox_streams(5)
        one = Trace(1, C_, 2, 1, blk.device, snapshots=True)
        blk.propose("DAMH", prop, sur, one)
        blk.decide("DAMH", prop, sur, one, torch.zeros((1, C_), dtype=torch.float64, device="cuda"))
    prob4 = DeviceProblem(4, prior_mean=0.0, prior_std=1.0, noise_std=[0.2, 0.1, 0.1], no_observations=3, observations=[0.1, 0.2, 0.3])
    m4 = DeviceModel.rbf(SOL4[0], SOL4[1], 1)
    blk = ChainBlock(prob4, 50, seed=2)
    blk.prepare("DAMH", m4, m4); blk.run("DAMH", 10, Proposal(4, 0.2), m4, m4, lanes=8)
    torch.cuda.synchronize()
print("sanity ok")
