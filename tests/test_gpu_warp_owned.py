"""-m gpu: the warp-owned launch shape of the chain kernel (csrc/sdb_chain_kernel.cuh, WT; sdb_wt_points in sdb_models.cuh).

A warp owns up to 32 chains and all 32 lanes split the centers of every surrogate evaluation; how many chains a warp
gets depends on the number of chains, on the SMs left to the collector (`coresident`) and on the number of in-kernel
waves.  The 32 partial sums of a point are always combined by the same tree, so the results must not depend on any of
that: bit-identical samples, counters and traced posteriors for every split.  Against the lane-group shape
(explicit `lanes`) the summation order differs, so surrogate posteriors agree to rounding and -- on these benign
expansions -- every decision is the same.  (Evaluation follows surrogate_rbf.py:20-40, the chain classes_SAMPLER.py:165-220.)
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import problem as OPR
from oracle import surrogate_rbf as ORB
from tests.golden_util import DARCY_PROBLEM, SIMPLE_PROBLEM

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from surrdamh_b200 import _lib
    _lib.require_cuda()
    return _lib


def _models(d, m, n, seed):
    from surrdamh_b200.device import DeviceModel
    rs = np.random.RandomState(seed)
    if d == 2:
        # a FITTED interpolant of the toy model (random weights would be rejected wholesale at noise_std 2e-4)
        from surrdamh_b200 import surrogate_rbf as SR
        Lc = np.linalg.cholesky(np.array([[4.0, -2.0], [-2.0, 4.0]]))
        X = rs.standard_normal((n, 2)) @ Lc.T + np.array([5.0, 3.0])
        Y = (-0.0375 * np.exp(-X[:, 0]) - 0.0125 * np.exp(-X[:, 1])).reshape(-1, 1)
        up = SR.Surrogate_update(2, 1, kernel_type=1, solver_type="direct")
        up.add_arrays(X, Y)
        sur, _ = up.update_device()
        SOL = sur.sol_numpy()
        return sur, DeviceModel.toy(), (SOL[0], SOL[1])
    centers = rs.standard_normal((n, d)) * (2.0 if d == 2 else 1.0) + (np.array([5.0, 3.0]) if d == 2 else 0.0)
    coefs = rs.standard_normal((n + d + 1, m)) * 1e-3
    truth = DeviceModel.toy() if d == 2 else DeviceModel.zero(d, m)
    return DeviceModel.rbf(coefs, centers, 1), truth, (coefs, centers)


def _run(kw, d, m, sur, truth, C_, K, std, coresident=0, lanes=0, f32=False, chain_ranges=None, seed=5):
    lanes = lanes or -1                             # 0 would pick lane groups for these small tables: force the warp-owned shape
    from surrdamh_b200.chains import ChainBlock, Proposal, Trace
    from surrdamh_b200.device import DeviceProblem
    prob = DeviceProblem(d, **kw)
    prop = Proposal(d, std)
    if chain_ranges is None:
        blk = ChainBlock(prob, C_, seed=seed)
        blk.surrogate_f32 = f32
        blk.prepare("DAMH", sur, truth)
        tr = Trace(K, C_, d, m, blk.device, snapshots=True, full=True)
        blk.run("DAMH", K, prop, sur, truth, trace=tr, lanes=lanes, coresident=coresident)
        torch.cuda.synchronize()
        return dict(u=blk.u.clone(), counters=blk.counters.clone(), post=blk.post_u.clone(), pre_u=blk.pre_u.clone(),
                    pre=tr.pre.clone(), flags=tr.flags.clone(), prop=tr.prop.clone(),
                    plan=blk.plan("DAMH", sur, truth, prop, lanes=lanes))
    parts = []
    for lo, n in chain_ranges:                    # separate blocks of chains, as ranks of a sharded run hold them
        blk = ChainBlock(prob, n, chain0=lo, seed=seed)
        blk.surrogate_f32 = f32
        blk.prepare("DAMH", sur, truth)
        tr = Trace(K, n, d, m, blk.device, snapshots=True, full=True)
        blk.run("DAMH", K, prop, sur, truth, trace=tr, lanes=lanes, coresident=coresident)
        torch.cuda.synchronize()
        parts.append(dict(u=blk.u.clone(), counters=blk.counters.clone(), post=blk.post_u.clone(), pre_u=blk.pre_u.clone(),
                          pre=tr.pre.clone(), flags=tr.flags.clone()))
    return {k: torch.cat([p[k] for p in parts], dim=-1) for k in parts[0]}


def _same(a, b):
    for k in ("u", "counters", "post", "pre_u", "flags"):
        assert torch.equal(a[k], b[k]), k
    assert np.array_equal(a["pre"].cpu().numpy(), b["pre"].cpu().numpy(), equal_nan=True)


@pytest.mark.parametrize("shape", ["d2m1", "d4m3", "d3m2"])
def test_results_do_not_depend_on_the_split_of_chains_over_warps(L, shape):
    """5000 chains: default launch, 100 / 140 / 146 SMs left free (more chains per warp, then several in-kernel waves),
    and the chains run as three separately launched shards -- identical bit for bit."""
    if shape == "d2m1":
        d, m, kw, n = 2, 1, SIMPLE_PROBLEM, 1000
    elif shape == "d4m3":
        d, m, kw, n = 4, 3, DARCY_PROBLEM, 777
    else:
        d, m, n = 3, 2, 300
        kw = dict(prior_mean=0.0, prior_std=1.0, noise_std=0.5, no_observations=2, observations=[0.3, -0.2])
    sur, truth, _ = _models(d, m, n, 3)
    C_, K = 5000, 10
    ref = _run(kw, d, m, sur, truth, C_, K, 0.5)
    assert ref["plan"]["lanes_per_chain"] == 0                  # the warp-owned shape ran (forced: the tables are small)
    cnt = ref["counters"].cpu().numpy()
    assert cnt[2].sum() > 0 and cnt[0].sum() + cnt[1].sum() > 0                         # both stages were exercised
    if shape == "d2m1":
        assert cnt[0].sum() > 0                                                          # ... and chains moved
    for free in (100, 140, 146):
        _same(ref, _run(kw, d, m, sur, truth, C_, K, 0.5, coresident=free))
    _same(ref, _run(kw, d, m, sur, truth, C_, K, 0.5, chain_ranges=[(0, 1), (1, 2047), (2048, 2952)]))


def test_single_chain_and_ragged_counts(L):
    """1, 2, 31, 33 and 149 chains (fewer chains than warps, one chain per warp, a ragged last warp)."""
    sur, truth, _ = _models(2, 1, 257, 4)
    big = _run(SIMPLE_PROBLEM, 2, 1, sur, truth, 149, 8, 0.5)
    for C_ in (1, 2, 31, 33):
        small = _run(SIMPLE_PROBLEM, 2, 1, sur, truth, C_, 8, 0.5)
        for k in ("u", "counters", "post", "pre_u", "flags"):
            assert torch.equal(small[k], big[k][..., :C_]), (C_, k)


def test_warp_owned_agrees_with_lane_groups_and_the_oracle(L):
    """Against the lane-group shape (lanes = 8, 32) the traced surrogate posteriors agree to rounding, all decisions are
    the same; sampled proposals are re-evaluated by the numpy oracle (rtol 1e-10)."""
    d, m, n, C_, K = 2, 1, 1000, 3000, 12
    sur, truth, (coefs, centers) = _models(d, m, n, 3)
    wt = _run(SIMPLE_PROBLEM, d, m, sur, truth, C_, K, 0.5)
    ev = ORB.RbfApply(d, m, kernel_type=1)
    gp = OPR.GaussProblem(d, **SIMPLE_PROBLEM)
    sigma2 = float(np.asarray(SIMPLE_PROBLEM["noise_std"]).ravel()[0]) ** 2

    def tolerance(P, S, want):
        # the fitted weights alternate in sign: two summation orders of sum_i w_i phi_i differ by a few n eps sum_i |w_i phi_i|,
        # and the likelihood multiplies a difference in S by |obs - S| / sigma^2 (same bound as tests/test_gpu_fullsize.py)
        S_abs = ev.apply([np.abs(coefs), centers], P)[:, 0]
        amp = np.abs(np.asarray(SIMPLE_PROBLEM["observations"])[0] - S[:, 0]) / sigma2
        return 1e-10 * np.abs(want) + amp * 256 * 2.2e-16 * S_abs

    prop_all = wt["prop"].cpu().numpy()                           # [K][d][C]
    for lanes in (8, 32):
        lg = _run(SIMPLE_PROBLEM, d, m, sur, truth, C_, K, 0.5, lanes=lanes)
        assert torch.equal(wt["flags"], lg["flags"]) and torch.equal(wt["counters"], lg["counters"])
        assert torch.equal(wt["u"], lg["u"])
        a, b = wt["pre"].cpu().numpy(), lg["pre"].cpu().numpy()
        assert np.array_equal(np.isfinite(a), np.isfinite(b))
        for k in (0, K - 1):
            P = np.ascontiguousarray(prop_all[k].T)
            S = ev.apply([coefs, centers], P)
            assert np.all(np.abs(a[k] - b[k]) <= 2 * tolerance(P, S, a[k])), (lanes, k)
    rs = np.random.RandomState(0)
    chains = np.unique(np.concatenate([[0, 1, 31, 32, C_ - 1], rs.randint(0, C_, 40)]))
    prop = prop_all[:, :, chains]
    pre = wt["pre"].cpu().numpy()[:, chains]
    for k in range(K):
        P = np.ascontiguousarray(prop[k].T)
        S = ev.apply([coefs, centers], P)
        want = np.array([gp.log_likelihood(S[i]) + gp.log_prior(P[i]) for i in range(len(chains))])
        err = np.abs(pre[k] - want)
        assert np.all(err <= tolerance(P, S, want)), (k, float(err.max()))
        assert np.all(err <= 1e-6 * np.abs(want) + 1e-9), (k, float(err.max()))


def test_f32_surrogate_split_invariance(L):
    """The FP32-surrogate instantiation of the warp-owned kernel: same invariance."""
    sur, truth, _ = _models(2, 1, 1000, 3)
    ref = _run(SIMPLE_PROBLEM, 2, 1, sur, truth, 4000, 8, 0.5, f32=True)
    _same(ref, _run(SIMPLE_PROBLEM, 2, 1, sur, truth, 4000, 8, 0.5, f32=True, coresident=144))


@pytest.mark.parametrize("d,m,n", [(2, 1, 1000), (4, 3, 777), (3, 2, 300)])
def test_stand_alone_evaluation_of_many_points(L, d, m, n):
    """sdb_rbf_eval with enough points for the persistent warp-owned kernel (rbf_eval_wt_kernel: >= 16 warps x 32 points per
    SM): 80 007 points (a ragged last tile) against the numpy oracle, all kernel types (surrogate_rbf.py:20-40, 146-173)."""
    from surrdamh_b200 import surrogate_rbf as SR
    rs = np.random.RandomState(d * 100 + m)
    N = 80007
    centers = rs.standard_normal((n, d))
    coefs = rs.standard_normal((n + d + 1, m))
    pts = rs.standard_normal((N, d)) * 1.5
    for kt in (0, 1, 2, 3, 4, 5, 7):
        got = SR.Surrogate_apply(d, m, kernel_type=kt).apply([coefs, centers], pts)
        want = ORB.RbfApply(d, m, kernel_type=kt).apply([coefs, centers], pts)
        scale = ORB.RbfApply(d, m, kernel_type=kt).apply([np.abs(coefs), centers], np.abs(pts))       # sum |w phi| (+ |tail|)
        assert got.shape == want.shape == (N, m)
        assert np.all(np.abs(got - want) <= 1e-10 * np.abs(want) + 64 * 2.2e-16 * np.abs(scale)), kt


def test_split_phases_use_the_warp_owned_shape_and_change_nothing(L):
    """PROPOSE / DECIDE launches (host-evaluated G, classes_SAMPLER.py:77-81 through the solver plugin) with a 1200-center RBF surrogate
    -- the library picks the warp-owned shape on its own -- against the fused launch with the zero model as G: identical samples,
    counters and posteriors, also when the two phases are issued for two halves of the chains (the pipelined split path)."""
    from surrdamh_b200.chains import ChainBlock, Proposal, Trace
    from surrdamh_b200.device import DeviceModel, DeviceProblem
    d, m, C_, K = 2, 1, 2500, 7
    rs = np.random.RandomState(12)
    centers = rs.standard_normal((1200, 2)) * 2.0 + np.array([5.0, 3.0])
    coefs = rs.standard_normal((1203, 1)) * 1e-9
    coefs[1200:, 0] = [2e-5, -1e-5, -1.05e-3]
    sur, zero = DeviceModel.rbf(coefs, centers, 1), DeviceModel.zero(d, m)
    prop = Proposal(d, [0.4, 0.2])
    prob = DeviceProblem(d, **SIMPLE_PROBLEM)

    fused = ChainBlock(prob, C_, seed=21)
    assert fused.plan("DAMH", sur, zero, prop)["lanes_per_chain"] == 0          # warp-owned, chosen by the library
    fused.prepare("DAMH", sur, zero)
    fused.run("DAMH", K, prop, sur, zero)
    torch.cuda.synchronize()
    assert int(fused.counters[0].sum()) > 0 and int(fused.counters[2].sum()) > 0

    for halves in ([(0, C_)], [(0, 1100), (1100, C_ - 1100)]):
        blk = ChainBlock(prob, C_, seed=21)
        G0 = np.zeros((C_, m))
        blk.prepare("DAMH", sur, None, G_host=G0)
        one = Trace(1, C_, d, m, blk.device, snapshots=False)
        Gd = torch.zeros((m, C_), dtype=torch.float64, device=blk.device)
        for k in range(K):
            for lo, n in halves:
                blk.propose("DAMH", prop, sur, one, chains=(lo, n), step0=k)
            for i, (lo, n) in enumerate(halves):
                blk.decide("DAMH", prop, sur, one, Gd, chains=(lo, n), advance=(i == len(halves) - 1), step0=k)
        torch.cuda.synchronize()
        for a, b in ((fused.u, blk.u), (fused.counters, blk.counters), (fused.post_u, blk.post_u), (fused.pre_u, blk.pre_u)):
            assert torch.equal(a, b), halves
