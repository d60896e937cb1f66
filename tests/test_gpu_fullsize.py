"""BASELINE.json configs[2] at FULL size on one GPU -- 65 536 chains, cubic RBF surrogate with 4096 centers, DAMH -- checked through
properties that do not need the (slow) oracle to walk 65 536 chains:

* the result does not depend on how the steps are cut into launches (one launch of 6 steps == 2 + 4);
* nor on how the chains are cut into blocks (65 536 == 2 x 32 768 with chain0 offsets): Philox is keyed by the global chain id;
* the traced surrogate log-posterior of a sample of (step, chain) proposals equals the numpy oracle's RBF evaluation
  (surrogate_rbf.py:20-40) + Gaussian log-posterior (classes_SAMPLER.py:295-316) at those proposals, to rtol 1e-10 plus the
  rounding bound of the 4096-term sum itself (stated in the test);
* the traced flags obey the delayed-acceptance logic given the traced posteriors and the Philox uniforms (classes_SAMPLER.py:
  193-195): re-deriving every decision of the sampled chains from (pre, post, uniforms) reproduces the flags bit for bit;
* counters add up: accepted + rejected + pre-rejected == steps for every chain.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import problem as OPR
from oracle import surrogate_rbf as ORB
from tests.golden_util import SIMPLE_PROBLEM

pytestmark = pytest.mark.gpu

C_FULL, N_CENTERS, K = 65536, 4096, 6


def _toy_G(X):
    return (-0.0375 * np.exp(-X[:, 0]) - 0.0125 * np.exp(-X[:, 1])).reshape(-1, 1)


@pytest.fixture(scope="module")
def setup():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from surrdamh_b200 import surrogate_rbf as SR
    from surrdamh_b200.device import DeviceModel, DeviceProblem
    rs = np.random.RandomState(1)
    Lc = np.linalg.cholesky(np.array(SIMPLE_PROBLEM["prior_std"], dtype=float))
    X = rs.standard_normal((N_CENTERS, 2)) @ Lc.T + np.array(SIMPLE_PROBLEM["prior_mean"])
    up = SR.Surrogate_update(2, 1, kernel_type=1, solver_type="nullspace", no_keep=N_CENTERS)
    up.add_arrays(X, _toy_G(X))
    sur, _ = up.update_device()
    assert getattr(up, "fallbacks", 0) == 0
    return DeviceProblem(2, **SIMPLE_PROBLEM), sur, DeviceModel.toy()


def _run(prob, sur, truth, n_chains, chain0, cuts, trace=None):
    from surrdamh_b200.chains import ChainBlock, Proposal
    blk = ChainBlock(prob, n_chains, chain0=chain0, seed=11)
    blk.prepare("DAMH", sur, truth)
    prop = Proposal(2, [0.4, 0.2])
    for k in cuts:
        blk.run("DAMH", k, prop, sur, truth, trace=trace if len(cuts) == 1 else None)
    torch.cuda.synchronize()
    return blk


def test_cfg3_full_size_properties(setup):
    from surrdamh_b200 import _lib as L
    from surrdamh_b200.chains import ChainBlock, Proposal, Trace
    prob, sur, truth = setup
    tr = Trace(K, C_FULL, 2, 1, torch.device("cuda"), snapshots=False, full=True)
    whole = _run(prob, sur, truth, C_FULL, 0, [K], trace=tr)
    plan = whole.plan("DAMH", sur, truth, Proposal(2, [0.4, 0.2]))
    assert plan["blocks"] * plan["threads"] >= C_FULL * plan["lanes_per_chain"]            # the bench's launch shape

    # launches: 6 == 2 + 4
    cut = _run(prob, sur, truth, C_FULL, 0, [2, 4])
    assert torch.equal(whole.u, cut.u) and torch.equal(whole.counters, cut.counters)
    assert torch.equal(whole.post_u, cut.post_u) and torch.equal(whole.pre_u, cut.pre_u)

    # blocks: 65 536 == 2 x 32 768
    halves = [_run(prob, sur, truth, C_FULL // 2, r * (C_FULL // 2), [K]) for r in range(2)]
    assert torch.equal(whole.u, torch.cat([h.u for h in halves], dim=1))
    assert torch.equal(whole.counters, torch.cat([h.counters for h in halves], dim=1))

    # counters add up
    cnt = whole.counters.cpu().numpy()
    assert np.array_equal(cnt[0] + cnt[1] + cnt[2], np.full(C_FULL, K))
    assert cnt[0].sum() > 0 and cnt[2].sum() > 0          # the run exercised accepts and pre-rejections

    # traced surrogate posteriors of sampled proposals vs the oracle
    rs = np.random.RandomState(2)
    chains = np.unique(np.concatenate([[0, 1, C_FULL // 2 - 1, C_FULL // 2, C_FULL - 1], rs.randint(0, C_FULL, 59)]))
    prop_t = tr.prop.cpu().numpy()[:, :, chains]          # [K][d][c]
    pre_t = tr.pre.cpu().numpy()[:, chains]
    post_t = tr.post.cpu().numpy()[:, chains]
    flags = tr.flags.cpu().numpy()[:, chains]
    pre_cur = tr.pre_cur.cpu().numpy()[:, chains]
    SOL = sur.sol_numpy()
    ev = ORB.RbfApply(2, 1, kernel_type=1)
    gp = OPR.GaussProblem(2, **SIMPLE_PROBLEM)
    for k in range(K):
        P = np.ascontiguousarray(prop_t[k].T)            # [c][d]
        S = ev.apply(SOL, P)
        want = np.array([gp.log_likelihood(S[i]) + gp.log_prior(P[i]) for i in range(len(chains))])
        # The fitted coefficients of 4096 scattered centers alternate in sign (cond ~ 1e9): ANY two summation orders of
        # sum_i w_i phi_i differ by a few n*eps*sum_i |w_i phi_i|, and the likelihood multiplies a difference in S by
        # |obs - S| / sigma^2.  Tolerance = the north_star's rtol 1e-10 + that rounding bound of the sum itself.
        S_abs = ev.apply([np.abs(SOL[0]), SOL[1]], np.abs(P))[:, 0]
        sigma2 = float(np.asarray(SIMPLE_PROBLEM["noise_std"]).ravel()[0]) ** 2
        amp = np.abs(np.asarray(SIMPLE_PROBLEM["observations"])[0] - S[:, 0]) / sigma2
        tol = 1e-10 * np.abs(want) + amp * 256 * 2.2e-16 * S_abs
        err = np.abs(pre_t[k] - want)
        assert np.all(err <= tol), (k, float((err / tol).max()), float(err.max()), float(tol.min()))
        assert np.all(err <= 1e-6 * np.abs(want) + 1e-9), (k, float(err.max()))      # and never worse than this, whatever the bound says

    # re-derive every decision of the sampled chains from the traced posteriors and the Philox uniforms
    z, uni = whole.philox_streams(K, step0=0)
    uni = uni.cpu().numpy().reshape(K, 2, C_FULL)[:, :, chains]
    # the current posterior before step 0 is not traced: take it from the prepared state of a fresh block and walk forward
    fresh = ChainBlock(prob, C_FULL, chain0=0, seed=11)
    fresh.prepare("DAMH", sur, truth)
    torch.cuda.synchronize()
    post_cur = fresh.post_u.cpu().numpy()[chains].copy()
    with np.errstate(divide="ignore"):
        lu = np.log(uni)
    for k in range(K):
        with np.errstate(invalid="ignore"):               # NaN posteriors of pre-rejected steps compare False, as intended
            pass1 = lu[k, 0] < (pre_t[k] - pre_cur[k])
            acc = pass1 & (lu[k, 1] < ((post_t[k] - post_cur) - (pre_t[k] - pre_cur[k])))
        want_flag = np.where(acc, L.FLAG_ACCEPTED, np.where(pass1, L.FLAG_REJECTED, L.FLAG_PREREJECTED))
        assert np.array_equal(flags[k], want_flag), k
        post_cur = np.where(acc, post_t[k], post_cur)


@pytest.mark.parametrize("name", ["cfg5", "cfg2"])
def test_other_bench_shapes_full_size_cut_invariance(name):
    """BASELINE.json configs[4] (131 072 chains, frozen RBF 500 centers, 2000-center RBF truth, d = 4, m = 3) and configs[1]
    (4096 chains, frozen polynomial surrogate: the lean kernel) with the models bench.py builds: cutting the steps into
    launches or the chains into blocks does not change a bit, and the counters add up."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import bench
    from surrdamh_b200.chains import ChainBlock, Proposal
    from surrdamh_b200.device import DeviceProblem
    w = bench.WORKLOADS[name]
    _, sur, truth = bench.build_models(w)
    prob = DeviceProblem(w["d"], **w["problem"])
    C_, steps = w["chains"], 5

    def run(n_chains, chain0, cuts):
        blk = ChainBlock(prob, n_chains, chain0=chain0, seed=23)
        blk.prepare("DAMH", sur, truth)
        prop = Proposal(w["d"], w["proposal_std"])
        for k in cuts:
            blk.run("DAMH", k, prop, sur, truth)
        torch.cuda.synchronize()
        return blk

    whole = run(C_, 0, [steps])
    cut = run(C_, 0, [1, 4])
    assert torch.equal(whole.u, cut.u) and torch.equal(whole.counters, cut.counters) and torch.equal(whole.post_u, cut.post_u)
    quarters = [run(C_ // 4, r * (C_ // 4), [steps]) for r in range(4)]
    assert torch.equal(whole.u, torch.cat([q.u for q in quarters], dim=1))
    assert torch.equal(whole.G_u, torch.cat([q.G_u for q in quarters], dim=1))
    assert torch.equal(whole.counters, torch.cat([q.counters for q in quarters], dim=1))
    cnt = whole.counters.cpu().numpy()
    assert np.array_equal(cnt[0] + cnt[1] + cnt[2], np.full(C_, steps))
    assert cnt[0].sum() > 0 and cnt[2].sum() > 0


def test_refit_at_the_largest_sweep_size_residual():
    """BASELINE.json configs[3] at its largest size, n = 32 768 snapshots (d = 4, m = 3): no host solver finishes on that in test
    time, so the fit is checked through its own defining equations ON THE DEVICE -- the saddle system S [c; lambda] = [Y; 0]
    (surrogate_rbf.py:109-127) rebuilt by sdb_rbf_system and multiplied out with torch: normwise backward error of LAPACK
    quality, polynomial side conditions P^T c = 0 met, and the fitted surrogate interpolates its training values."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from surrdamh_b200 import _lib as L
    from surrdamh_b200 import surrogate_rbf as SR
    n, d, m = 32768, 4, 3
    if torch.cuda.get_device_properties(0).total_memory < 40e9:
        pytest.skip("needs 40 GB of device memory")
    rs = np.random.RandomState(9)
    X = rs.standard_normal((n, d))
    Y = np.stack([np.sin((k + 1) * X).sum(axis=1) for k in range(m)], axis=1)
    up = SR.Surrogate_update(d, m, kernel_type=1, solver_type="nullspace")
    up.add_arrays(X, Y)
    model, nn = up.update_device()
    assert nn == n and getattr(up, "fallbacks", 0) == 0
    CO = model.coefs                                      # [n + d + 1][m] on the device
    N = n + d + 1
    S = torch.empty((N, N), dtype=torch.float64, device="cuda")
    R = torch.empty((m, N), dtype=torch.float64, device="cuda")
    dX, dY = L.dev(X), L.dev(Y)
    L.check(L.lib().sdb_rbf_system(L.ptr(dX), L.ptr(dY), n, d, m, 1, L.ptr(S), L.ptr(R), L.stream_handle()))
    res = S @ CO - R.T
    berr = (torch.linalg.norm(res) / (torch.linalg.norm(S) * torch.linalg.norm(CO))).item()
    assert berr <= 1e-15, berr                            # LAPACK's dgesv reaches ~1e-17 on such systems (profiles/r01_refit_sweep.md)
    side = torch.linalg.norm(res[n:]).item()              # rows n.. of S are [P^T 0]: the side conditions
    assert side <= 1e-9 * torch.linalg.norm(CO[:n]).item(), side
    fit = (S[:n] @ CO)                                    # the surrogate at its own centers
    assert torch.allclose(fit, dY, rtol=0, atol=1e-7 * float(np.abs(Y).max())), (fit - dY).abs().max().item()
