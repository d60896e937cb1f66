"""-m gpu: the two additional ways the chain kernel holds an RBF surrogate (csrc/sdb_chain_kernel.cuh, SM = 1, 2).

Streamed surrogate (tables larger than shared memory; the reference has no limit on the number of centers,
surrogate_rbf.py:43,97): bit-identical to the resident path when forced on a table that would fit, and a DAMH run with
16 384 centers checked against the stand-alone evaluation kernel and the numpy oracle.

FP32 surrogate (surrogate_precision "f32", north_star: rtol 1e-4 for any fp32 path): surrogate values within 1e-4 of
the reference's float64 evaluation on its own golden vectors (all kernel types), the error bound
1e-6 * sum |w_j phi_j| on a fitted (cancelling) expansion, and a DAMH run whose posterior moments agree with the FP64
run within Monte Carlo error -- stage 2 corrects with the true posterior, so the chain is exact either way.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import problem as OPR
from oracle import surrogate_rbf as ORB
from tests.golden_util import DARCY_PROBLEM, SIMPLE_PROBLEM, load

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from surrdamh_b200 import _lib
    _lib.require_cuda()
    return _lib


def _prior_points(rs, n):
    Lc = np.linalg.cholesky(np.array([[4.0, -2.0], [-2.0, 4.0]]))
    return rs.standard_normal((n, 2)) @ Lc.T + np.array([5.0, 3.0])


def _toy_G(X):
    return (-0.0375 * np.exp(-X[:, 0]) - 0.0125 * np.exp(-X[:, 1])).reshape(-1, 1)


def _run_chains(L, kw, d, m, sur, truth, C_, K, std, seed=5, f32=False, lanes=0, cuts=(None,), trace_full=True):
    from surrdamh_b200.chains import ChainBlock, Proposal, Trace
    from surrdamh_b200.device import DeviceProblem
    prob = DeviceProblem(d, **kw)
    blk = ChainBlock(prob, C_, seed=seed)
    blk.surrogate_f32 = f32
    prop = Proposal(d, std)
    blk.prepare("DAMH", sur, truth)
    traces = []
    for k in (cuts if cuts[0] is not None else (K,)):
        tr = Trace(k, C_, d, m, blk.device, snapshots=False, full=trace_full)
        blk.run("DAMH", k, prop, sur, truth, trace=tr, lanes=lanes)
        traces.append(tr)
    torch.cuda.synchronize()
    return blk, traces


# ---------------------------------------------------------------------------------------------------------- streaming
_CHILD = r'''
import sys, json, numpy as np, torch
sys.path.insert(0, %(root)r)
from tests.test_gpu_precision_stream import _run_chains, _prior_points, _toy_G
from tests.golden_util import SIMPLE_PROBLEM, DARCY_PROBLEM
from surrdamh_b200 import _lib as L
from surrdamh_b200.device import DeviceModel
rs = np.random.RandomState(3)
out = {}
for tag, (d, m, n, kw, C_, lanes) in {"a": (2, 1, 1000, SIMPLE_PROBLEM, 3000, 0), "b": (4, 3, 777, DARCY_PROBLEM, 500, 0),
                                     "c": (3, 2, 300, None, 200, 0), "d": (2, 1, 1000, SIMPLE_PROBLEM, 64, 32)}.items():
    if kw is None:
        kw = dict(prior_mean=0.0, prior_std=1.0, noise_std=0.5, no_observations=2, observations=[0.3, -0.2])
    centers = rs.standard_normal((n, d)) * (2.0 if d == 2 else 1.0) + (np.array([5.0, 3.0]) if d == 2 else 0.0)
    coefs = rs.standard_normal((n + d + 1, m)) * 1e-3
    sur = DeviceModel.rbf(coefs, centers, 1)
    truth = DeviceModel.toy() if d == 2 else DeviceModel.zero(d, m)
    blk, trs = _run_chains(L, kw, d, m, sur, truth, C_, 12, 0.5, lanes=lanes)
    out[tag] = [blk.u.cpu().numpy().tolist(), blk.counters.cpu().numpy().tolist(), trs[0].pre.cpu().numpy().tolist()]
json.dump(out, open(sys.argv[1], "w"))
'''


def test_streamed_surrogate_is_bit_identical_to_resident(L, tmp_path):
    """The same launches with the tables resident and with SDB_STREAM_CHUNK forcing the TMA ring (32-, 96- and 512-center
    chunks: ragged last chunk, several chunks per evaluation, d = 2 / 3 / 4, 1 and 32 lanes per chain): identical final
    samples, counters and traced surrogate posteriors, bit for bit."""
    script = tmp_path / "child.py"
    script.write_text(_CHILD % {"root": ROOT})
    results = {}
    for chunk in ("0", "32", "96", "512"):
        env = dict(os.environ)
        env.pop("SDB_STREAM_CHUNK", None)
        if chunk != "0":
            env["SDB_STREAM_CHUNK"] = chunk
        out = tmp_path / ("out_%s.json" % chunk)
        subprocess.check_call([sys.executable, str(script), str(out)], env=env, cwd=ROOT)
        import json
        results[chunk] = json.load(open(out))
    for chunk in ("32", "96", "512"):
        for tag in results["0"]:
            a, b = results["0"][tag], results[chunk][tag]
            assert a[0] == b[0] and a[1] == b[1], (chunk, tag)
            assert np.array_equal(np.array(a[2]), np.array(b[2]), equal_nan=True), (chunk, tag)


def test_damh_with_16384_centers_streams_and_matches_oracle(L):
    """n = 16 384 centers, d = 2, m = 1: 393 KB of tables, more than an SM's shared memory -> the chain kernel streams them.
    Traced surrogate posteriors of the proposals equal the log-posterior of the stand-alone evaluation (rtol 1e-10) and
    of the numpy oracle at sampled proposals; every decision is re-derived from the traced posteriors and the Philox
    uniforms; cutting the launch (20 = 8 + 12 steps) changes nothing."""
    from surrdamh_b200 import surrogate_rbf as SR
    from surrdamh_b200.device import DeviceModel
    rs = np.random.RandomState(11)
    n, C_, K = 16384, 2048, 20
    centers = _prior_points(rs, n)
    coefs = rs.standard_normal((n + 3, 1)) * 2e-7
    coefs[n:, 0] = [1e-4, -5e-5, -1.2e-3]
    sur = DeviceModel.rbf(coefs, centers, 1)
    plan = None
    from surrdamh_b200.chains import ChainBlock, Proposal
    from surrdamh_b200.device import DeviceProblem
    blk0 = ChainBlock(DeviceProblem(2, **SIMPLE_PROBLEM), C_)
    plan = blk0.plan("DAMH", sur, DeviceModel.toy(), Proposal(2, 0.3))
    assert plan["smem"] < 16384 * 24                        # the tables are NOT resident
    blk, (tr,) = _run_chains(L, SIMPLE_PROBLEM, 2, 1, sur, DeviceModel.toy(), C_, K, 0.3, seed=9)
    blk2, trs = _run_chains(L, SIMPLE_PROBLEM, 2, 1, sur, DeviceModel.toy(), C_, K, 0.3, seed=9, cuts=(8, 12))
    assert torch.equal(blk.u, blk2.u) and torch.equal(blk.counters, blk2.counters)
    assert torch.equal(tr.flags, torch.cat([trs[0].flags, trs[1].flags]))
    props = tr.prop.cpu().numpy()                            # [K][d][C]
    pre = tr.pre.cpu().numpy()
    pts = props.transpose(0, 2, 1).reshape(-1, 2)
    S = SR.Surrogate_apply(2, 1, kernel_type=1).apply([coefs, centers], pts)
    prob = OPR.GaussProblem(2, **SIMPLE_PROBLEM)
    want = np.array([prob.log_posterior(pts[i], S[i]) for i in range(0, pts.shape[0], 97)])
    got = pre.reshape(-1)[::97]
    assert np.allclose(got, want, rtol=1e-10, atol=1e-10 * np.abs(want).max())
    idx = rs.choice(pts.shape[0], 40, replace=False)
    So = ORB.RbfApply(2, 1, kernel_type=1).apply([coefs, centers], pts[idx])
    assert np.allclose(S[idx], So, rtol=1e-10, atol=1e-10 * np.abs(So).max())
    # decisions from the traced values: stage 1 log(u0) < pre' - pre_cur, stage 2 log(u1) < (post' - post) - (pre' - pre_cur)
    z, uni = ChainBlock(DeviceProblem(2, **SIMPLE_PROBLEM), C_, seed=9).philox_streams(K)
    u0 = uni[:, 0].cpu().numpy()
    flags, pre_cur = tr.flags.cpu().numpy(), tr.pre_cur.cpu().numpy()
    with np.errstate(divide="ignore"):
        pass1 = np.log(u0) < (pre - pre_cur)
    assert np.array_equal(pass1, flags != 0)


# ---------------------------------------------------------------------------------------------------------- FP32
def test_f32_evaluation_within_1e4_of_reference_golden(L):
    """The reference's own golden evaluations (tests/golden/rbf.npz, all 8 kernel types + the two named shapes):
    rtol 1e-4 relative to the largest value (north_star's bound for an fp32 path)."""
    from surrdamh_b200 import surrogate_rbf as SR
    g = load("rbf")
    C_, CO, pts = g["apply_centers"], g["apply_coefs"], g["apply_pts"]
    for kt in range(8):
        want = g["apply_GS_kt%d" % kt]
        got = SR.Surrogate_apply(3, 2, kernel_type=kt, precision="f32").apply([CO, C_], pts)
        assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max(), kt
        assert np.abs(got - want).max() > 0                      # it really is the FP32 path
    for tag, (d, m) in {"cfg3": (2, 1), "cfg5": (4, 3)}.items():
        want = g[tag + "_GS"]
        got = SR.Surrogate_apply(d, m, kernel_type=1, precision="f32").apply([g[tag + "_coefs"], g[tag + "_centers"]], g[tag + "_pts"])
        assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max(), tag


def test_f32_error_is_bounded_by_the_cancellation_of_the_expansion(L):
    """A FITTED interpolant cancels heavily (sum |w_j phi_j| >> |value|): the FP32 error is bounded by 1e-6 * sum |w_j phi_j|
    (a few FP32 roundings per term), which is what decides whether 1e-4 of the value is reached; far-off parameters do
    not add to it (centers are stored relative to the first center)."""
    from surrdamh_b200 import surrogate_rbf as SR
    rs = np.random.RandomState(2)
    for shift in (0.0, 1e4):
        X = _prior_points(rs, 600)
        up = ORB.RbfUpdate(2, 1, kernel_type=1, solver_type="direct")
        up.add_arrays(X, _toy_G(X))
        SOL, _ = up.update()
        coefs, centers = SOL[0].copy(), SOL[1] + shift
        coefs[602, 0] -= shift * (coefs[600, 0] + coefs[601, 0])           # same function of the shifted parameters
        P = _prior_points(rs, 500) + shift
        want = ORB.RbfApply(2, 1, kernel_type=1).apply([coefs, centers], P)
        got = SR.Surrogate_apply(2, 1, kernel_type=1, precision="f32").apply([coefs, centers], P)
        D = ORB.pairwise_distance(P, centers)
        mass = (np.abs(coefs[:600, 0])[None, :] * D ** 3).sum(axis=1, keepdims=True)
        assert (np.abs(got - want) <= 1e-6 * mass + 1e-9 * np.abs(want).max()).all(), shift


def test_f32_surrogate_chain_samples_the_same_posterior(L):
    """cfg3 shape at reduced size: DAMH with the surrogate in FP32 vs FP64 -- both target the exact posterior (stage 2), so
    the pooled posterior mean / covariance agree within Monte Carlo error; the FP32 surrogate's stage-1 decisions differ
    from the FP64 ones only in a small fraction of steps."""
    from surrdamh_b200 import stats
    from surrdamh_b200 import surrogate_rbf as SR
    from surrdamh_b200.chains import ChainBlock, Proposal, Trace
    from surrdamh_b200.device import DeviceModel, DeviceProblem
    rs = np.random.RandomState(4)
    X = _prior_points(rs, 1024)
    up = SR.Surrogate_update(2, 1, kernel_type=1, solver_type="nullspace", no_keep=1024)
    up.add_arrays(X, _toy_G(X))
    sur, _ = up.update_device()
    prob = DeviceProblem(2, **SIMPLE_PROBLEM)
    prop = Proposal(2, 1.0)
    C_, burn, K = 8192, 200, 400
    res = {}
    for f32 in (False, True):
        blk = ChainBlock(prob, C_, seed=21, moments=True)
        blk.surrogate_f32 = f32
        blk.prepare("DAMH", sur, DeviceModel.toy())
        blk.run("DAMH", burn, prop, sur, DeviceModel.toy())
        blk.moments.zero_()
        blk.moment_steps = 0
        tr = Trace(K, C_, 2, 1, blk.device, snapshots=False, full=False)
        blk.run("DAMH", K, prop, sur, DeviceModel.toy(), trace=tr)
        mean, cov = stats.chain_moments(blk)
        per_chain = (blk.moments[:2] / K).cpu().numpy()          # chain means: their spread gives the Monte Carlo error
        res[f32] = (mean.cpu().numpy(), cov.cpu().numpy(), per_chain.std(axis=1) / np.sqrt(C_), tr.flags.cpu().numpy(),
                    blk.counters_numpy())
    m64, c64, se64, f64, n64 = res[False]
    m32, c32, se32, f32_, n32 = res[True]
    assert (np.abs(m32 - m64) <= 6.0 * np.sqrt(se64 ** 2 + se32 ** 2)).all(), (m32, m64, se64)
    assert np.allclose(c32, c64, rtol=0.05, atol=0.02)
    # first step of the measured window: same state? no -- the chains have diverged after the burn-in; compare rates instead
    acc64, acc32 = n64[:, 0].sum() / n64[:, :3].sum(), n32[:, 0].sum() / n32[:, :3].sum()
    assert abs(acc64 - acc32) < 0.01
    assert not np.array_equal(f64, f32_)                         # it really ran a different arithmetic


def test_f32_stage1_decisions_mostly_agree_with_f64(L):
    """From the SAME state and the same random numbers, one step: the fraction of chains whose stage-1 decision differs
    between the FP32 and the FP64 surrogate is small (it is the probability that log u falls between the two ratios)."""
    from surrdamh_b200 import surrogate_rbf as SR
    from surrdamh_b200.device import DeviceModel
    rs = np.random.RandomState(6)
    X = _prior_points(rs, 1024)
    up = SR.Surrogate_update(2, 1, kernel_type=1, solver_type="nullspace", no_keep=1024)
    up.add_arrays(X, _toy_G(X))
    sur, _ = up.update_device()
    flags = {}
    for f32 in (False, True):
        blk, (tr,) = _run_chains(L, SIMPLE_PROBLEM, 2, 1, sur, DeviceModel.toy(), 20000, 1, 1.0, seed=3, f32=f32)
        flags[f32] = (tr.flags.cpu().numpy()[0], tr.pre.cpu().numpy()[0])
    mismatch = np.mean((flags[False][0] != 0) != (flags[True][0] != 0))
    assert mismatch < 0.02, mismatch
    rel = np.abs(flags[True][1] - flags[False][1]) / np.maximum(1.0, np.abs(flags[False][1]))
    assert np.median(rel) < 0.05
