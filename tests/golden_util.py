"""Shared helpers for the parity tests: golden fixture access, problem blocks, streams."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SIMPLE_PROBLEM = dict(prior_mean=[5.0, 3.0], prior_std=[[4, -2], [-2, 4]], noise_std=2e-4,
                      no_observations=1, observations=[-1e-3])
DARCY_PROBLEM = dict(prior_mean=0.0, prior_std=1.0, noise_std=[0.2, 0.1, 0.1], no_observations=3,
                     observations=[10.24925199, -6.07103574, -4.17821625])
CORR_NOISE_PROBLEM = dict(prior_mean=[0.5, -0.5, 0.0], prior_std=[1.0, 2.0, 0.5],
                          noise_std=[[0.04, 0.01, 0.0], [0.01, 0.09, -0.02], [0.0, -0.02, 0.0625]],
                          no_observations=3, observations=[1.0, -2.0, 0.5])

# Value tolerance between two IEEE-754 implementations of the same formula whose
# exp/log/pow and BLAS reductions may differ in the last bits (SVML vs libm vs CUDA
# libdevice).  Decisions (flags, counters) and sample trajectories are compared exactly.
VAL_RTOL = 1e-12


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def stream_pair(seed, steps, d):
    """Same construction as oracle/make_golden.py:stream_pair."""
    z = np.random.RandomState(seed).standard_normal((steps, d))
    u = np.random.RandomState(seed + 1).uniform(size=(steps, 2))
    return z, u


def close(a, b, rtol=VAL_RTOL, atol_scale=None):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    if a.shape != b.shape:
        return False
    nan_ok = np.array_equal(np.isnan(a), np.isnan(b))
    finite = np.abs(b[np.isfinite(b)])
    scale = atol_scale if atol_scale is not None else (finite.max() if finite.size else 1.0)
    return nan_ok and np.allclose(a, b, rtol=rtol, atol=rtol * scale, equal_nan=True)
