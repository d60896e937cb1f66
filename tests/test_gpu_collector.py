"""-m gpu: the fixed-shape collector round (surrdamh_b200/collector.py, csrc/sdb_util.cu, csrc/sdb_fit.cu)
against numpy restatements and against the general path:
  * fair snapshot selection (evenly spaced over the (step, chain) order, header row)
  * masked greedy thinning == thinning of the compacted set (surrogate_rbf.py:97-120); coincident rows dropped
  * one fixed-shape round == add_data + update() of the general path, bit for bit
  * a failing solve keeps the previous surrogate and is reported one round late
  * all chains starting at the prior mean with an RBF 'direct' refit (the singular-system trap) runs through
  * asynchronous collector with a lag: deterministic, same surrogates as the synchronous schedule shifted by the lag
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from tests.test_oracle_golden import SIMPLE_CONF

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from surrdamh_b200 import _lib
    _lib.require_cuda()
    return _lib


def _random_trace(L, K, C_, d, m, seed, p_valid=0.3):
    from surrdamh_b200.chains import Trace
    rs = np.random.RandomState(seed)
    tr = Trace(K, C_, d, m, torch.device("cuda"), snapshots=True, full=False)
    flags = rs.choice([0, 1, 2], size=(K, C_), p=[1 - p_valid, p_valid / 2, p_valid / 2]).astype(np.uint8)
    par = rs.standard_normal((K, d, C_))
    obs = rs.standard_normal((K, m, C_))
    tr.flags.copy_(torch.from_numpy(flags))
    tr.snap_par.copy_(torch.from_numpy(par))
    tr.snap_obs.copy_(torch.from_numpy(obs))
    return tr, flags, par, obs


def _select_numpy(flags, par, obs, quota):
    K, C_ = flags.shape
    rows = []
    for k in range(K):
        for c in range(C_):
            if flags[k, c] in (1, 2):
                rows.append(np.concatenate([par[k, :, c], obs[k, :, c]]))
    T = len(rows)
    if T <= quota:
        return T, rows
    out = [None] * quota
    for o in range(T):
        a, b = (o * quota) // T, ((o + 1) * quota) // T
        if b > a:
            out[b - 1] = rows[o]
    return T, out


@pytest.mark.parametrize("K,C_,d,m,quota,p", [(7, 333, 2, 1, 64, 0.3), (3, 50, 4, 3, 500, 0.2), (5, 1100, 3, 2, 1, 0.5),
                                              (2, 40, 2, 1, 16, 0.0)])
def test_snapshot_select_is_fair_and_ordered(L, K, C_, d, m, quota, p):
    from surrdamh_b200.collector import Collector
    tr, flags, par, obs = _random_trace(L, K, C_, d, m, seed=K * 1000 + C_, p_valid=p)
    col = Collector(None, "rbf", d, m, snapshots_per_refit=quota)
    block = col.select(tr, K).cpu().numpy()
    T, want = _select_numpy(flags, par, obs, quota)
    assert block.shape == (quota + 1, d + m)
    assert block[0, 0] == min(T, quota) and block[0, 1] == T
    n = min(T, quota)
    if n:
        assert np.array_equal(block[1:1 + n], np.array(want[:n]))       # bit-exact copies, in (step, chain) order
    assert not block[1 + n:].any()
    if T > quota and quota >= 8:
        # evenly spaced: the ordinals of consecutive picks differ by floor(T/quota) or ceil(T/quota)
        ordinals = [o for o in range(T) if ((o + 1) * quota) // T > (o * quota) // T]
        gaps = np.diff(ordinals)
        assert gaps.min() >= T // quota and gaps.max() <= -(-T // quota)


def _thin(L, X, no_keep, alive=None, dedupe=0):
    n, d = X.shape
    Xd = L.dev(X)
    keep = torch.empty((n,), dtype=torch.uint8, device="cuda")
    work = torch.empty((L.lib().sdb_rbf_thin_work(n),), dtype=torch.float64, device="cuda")
    al = None if alive is None else torch.from_numpy(alive.astype(np.uint8)).cuda()
    L.check(L.lib().sdb_rbf_thin_masked(L.ptr(Xd), n, d, no_keep, L.ptr(al), dedupe, L.ptr(keep), L.ptr(work), L.stream_handle()))
    return keep.cpu().numpy().astype(bool)


@pytest.mark.parametrize("n,d,frac,no_keep", [(300, 2, 0.7, 150), (2000, 2, 0.9, 1500), (5000, 4, 0.5, 2000), (9000, 2, 0.8, 7000),
                                              (120, 7, 0.6, 200)])
def test_masked_thinning_equals_thinning_of_the_compacted_set(L, n, d, frac, no_keep):
    from oracle import surrogate_rbf as ORB
    rs = np.random.RandomState(n + d)
    X = np.round(rs.standard_normal((n, d)), 2)          # rounding provokes distance ties
    alive = rs.uniform(size=n) < frac
    keep_m = _thin(L, X, no_keep, alive)
    assert not keep_m[~alive].any()
    keep_c = _thin(L, X[alive], no_keep)
    assert np.array_equal(keep_m[alive], keep_c)
    assert keep_m.sum() == min(no_keep, alive.sum())
    if n <= 2000:
        # and both equal the reference's greedy rule (oracle restatement of surrogate_rbf.py:97-120)
        Xa = X[alive]
        if Xa.shape[0] > no_keep:
            kept = ORB.greedy_thinning(ORB.pairwise_distance(Xa, Xa), no_keep)
            assert np.array_equal(keep_c, np.asarray(kept, dtype=bool))


def test_coincident_rows_are_dropped(L):
    rs = np.random.RandomState(3)
    base = rs.standard_normal((40, 3))
    X = np.concatenate([base, base[:10], base[5:8], rs.standard_normal((20, 3))])
    keep = _thin(L, X, 10 ** 6, None, dedupe=1)                       # nothing to thin, only duplicates go
    Xk = X[keep]
    assert Xk.shape[0] == 60
    assert np.unique(Xk, axis=0).shape[0] == 60
    keep2 = _thin(L, X, 30, None, dedupe=1)                           # quota first, then no duplicate is left either way
    assert keep2.sum() == 30 and np.unique(X[keep2], axis=0).shape[0] == 30
    assert _thin(L, X, 10 ** 6, None, dedupe=0).all()


def _toy_G(X):
    return (-0.0375 * np.exp(-X[:, 0]) - 0.0125 * np.exp(-X[:, 1])).reshape(-1, 1)


def _prior_points(rs, n):
    Lc = np.linalg.cholesky(np.array([[4.0, -2.0], [-2.0, 4.0]]))
    return rs.standard_normal((n, 2)) @ Lc.T + np.array([5.0, 3.0])


@pytest.mark.parametrize("solver", ["nullspace", "direct"])
def test_fixed_shape_round_equals_general_update(L, solver):
    """Collector.round on the fixed-shape path vs add_arrays + update_device on the same snapshots: same kept centers
    (order included) and the same coefficients bit for bit -- the two paths run the same kernels on the same rows."""
    from surrdamh_b200 import surrogate_rbf as SR
    from surrdamh_b200.collector import Collector
    rs = np.random.RandomState(5)
    n0, K, C_, quota = 600, 6, 500, 96
    X0 = _prior_points(rs, n0)
    ups = [SR.Surrogate_update(2, 1, kernel_type=1, solver_type=solver, no_keep=n0) for _ in range(2)]
    models = []
    for up in ups:
        up.add_arrays(X0, _toy_G(X0))
        models.append(up.update_device()[0])
    col = Collector(ups[0], "rbf", 2, 1, kernel_type=1, snapshots_per_refit=quota)
    col.model = models[0]
    for rnd in range(3):
        tr, flags, par, obs = _random_trace(L, K, C_, 2, 1, seed=100 + rnd, p_valid=0.2)
        par = _prior_points(rs, K * C_).T.reshape(2, K, C_).transpose(1, 0, 2).copy()
        obs = _toy_G(par.transpose(0, 2, 1).reshape(-1, 2)).reshape(K, C_, 1).transpose(0, 2, 1).copy()
        tr.snap_par.copy_(torch.from_numpy(par))
        tr.snap_obs.copy_(torch.from_numpy(obs))
        assert col.round(tr, K, rnd)
        assert col.fixed
        T, rows = _select_numpy(flags, par, obs, quota)
        rows = np.array(rows[:min(T, quota)])
        ups[1].add_arrays(rows[:, :2], rows[:, 2:])
        ref, n = ups[1].update_device()
        assert n == n0
        assert np.array_equal(col.model.centers.cpu().numpy(), ref.centers.cpu().numpy())
        assert np.array_equal(col.model.coefs.cpu().numpy(), ref.coefs.cpu().numpy())
    log = col.poll_status(wait=True)
    assert [(s[1], s[3]) for s in log] == [(0, n0)] * 3 and all(s[4] == quota for s in log)
    assert ups[0].round_failures == 0


def test_failing_round_keeps_previous_surrogate(L):
    """Snapshots identical to each other AND beyond the thinning surplus make the projected kernel matrix singular:
    the round reports it and the chains keep the previous model and centers."""
    from surrdamh_b200 import surrogate_rbf as SR
    from surrdamh_b200.collector import Collector
    rs = np.random.RandomState(8)
    n0, K, C_, quota = 300, 2, 64, 32
    X0 = _prior_points(rs, n0)
    X0[7] = X0[3]                                            # the held set itself already contains a coincident pair
    # 'direct': two identical rows stay identical through the elimination until one is the pivot row, then the other is
    # exactly zero -> an exact zero pivot, reported by the LU (the Cholesky of the projected matrix would only see a tiny one)
    up = SR.Surrogate_update(2, 1, kernel_type=1, solver_type="direct", no_keep=n0)
    good = SR.Surrogate_update(2, 1, kernel_type=1, solver_type="direct", no_keep=n0)
    Xg = _prior_points(rs, n0)
    good.add_arrays(Xg, _toy_G(Xg))
    model, _ = good.update_device()
    up._par, up._obs = L.dev(X0), L.dev(_toy_G(X0))          # state with the duplicate, model from a healthy fit
    col = Collector(up, "rbf", 2, 1, kernel_type=1, snapshots_per_refit=quota)
    col.model = model
    tr, flags, par, obs = _random_trace(L, K, C_, 2, 1, seed=1, p_valid=0.0)       # no snapshot at all this round
    before = (model.coefs.cpu().numpy().copy(), model.centers.cpu().numpy().copy())
    col.round(tr, K, 0)
    assert col.fixed
    log = col.poll_status(wait=True)
    assert log[-1][1] == 1 and up.round_failures == 1
    # pack_from_state took the healthy model's coefficients with the updater's rows: what must survive is that pair
    assert np.array_equal(col.model.coefs.cpu().numpy(), before[0])
    assert np.array_equal(col.model.centers.cpu().numpy(), X0)


def _conf(**extra):
    from surrdamh_b200.configuration import Configuration
    c = dict(SIMPLE_CONF, solver_module_path=os.path.join(ROOT, "surrdamh_b200", "solvers.py"), solver_module_name="solvers",
             solver_init_name="Solver_linela2exp_local", solver_parameters={}, no_solvers=1, seed=4)
    c.update(extra)
    return Configuration(4, conf=c)


def test_prior_mean_start_with_direct_rbf_refit_does_not_go_singular(L):
    """every lockstep chain starts at the prior mean, so each chain's first acceptance emits the SAME
    displaced snapshot; with n <= no_keep nothing thins them away and the reference-faithful 'direct' solve would be
    singular.  The sampler's updater drops coincident rows (drop_coincident, default on); switched off, the singular
    system raises LinAlgError like np.linalg.solve."""
    from surrdamh_b200.sampler import LockstepSampler
    stages = [{"type": "MH", "max_samples": 40, "time_limit": 600, "proposal_std": 1.0, "surrogate_is_updated": True},
              {"type": "DAMH", "max_samples": 60, "time_limit": 600, "proposal_std": 1.0, "surrogate_is_updated": True}]
    kw = dict(surrogate_type="rbf", surr_solver_parameters={"kernel_type": 1}, initial_sample_type="prior_mean",
              surr_updater_parameters={"kernel_type": 1, "solver_type": "direct", "no_keep": 400}, refit_every=20,
              snapshots_per_refit=256, samplers_list=stages)
    smp = LockstepSampler(_conf(**kw), no_chains=512)
    res = smp.run()
    assert smp.collector.no_refits == 2 + 3
    mo = smp.collector.model
    assert torch.isfinite(mo.coefs).all() and mo.centers.shape[0] == 400
    C_ = mo.centers.cpu().numpy()
    assert np.unique(C_, axis=0).shape[0] == C_.shape[0]
    assert res[1].counters[:, 0].sum() > 0                       # the DAMH stage accepts: the surrogate is usable
    assert smp.collector.fixed                                   # 400 centers reached -> later rounds ran without host syncs
    assert smp.collector.updater.round_failures == 0
    with pytest.raises(np.linalg.LinAlgError):
        LockstepSampler(_conf(drop_coincident=False, **kw), no_chains=512).run()


def test_async_collector_is_the_synchronous_schedule_with_a_lag(L):
    """collector_async: round b runs beside block b+1 and its surrogate is used from block b+2.  The run is
    deterministic (two runs agree bit for bit) and its models are those of a synchronous run whose chains were fed the
    same lagging surrogates -- checked through the first asynchronous round, which sees the same snapshots as the
    synchronous one (block 0 runs under the same surrogate in both)."""
    from surrdamh_b200.sampler import LockstepSampler
    stages = [{"type": "MH", "max_samples": 60, "time_limit": 600, "proposal_std": 1.0, "surrogate_is_updated": True},
              {"type": "DAMH", "max_samples": 200, "time_limit": 600, "proposal_std": 1.0, "surrogate_is_updated": True}]
    kw = dict(surrogate_type="rbf", surr_solver_parameters={"kernel_type": 1},
              surr_updater_parameters={"kernel_type": 1, "solver_type": "nullspace", "no_keep": 300}, refit_every=20,
              snapshots_per_refit=128, samplers_list=stages)
    runs = []
    for mode in (True, True, False):
        smp = LockstepSampler(_conf(collector_async=mode, **kw), no_chains=300)
        smp.run()
        runs.append((smp.block.current_numpy(), smp.block.counters_numpy(), smp.collector.model.coefs.cpu().numpy(),
                     smp.collector.no_refits))
    assert np.array_equal(runs[0][0], runs[1][0]) and np.array_equal(runs[0][1], runs[1][1])
    assert np.array_equal(runs[0][2], runs[1][2])
    assert runs[0][3] == runs[2][3] == 3 + 10
    # same number of rounds as the synchronous schedule, different (lagging) surrogates -> different trajectories
    assert not np.array_equal(runs[0][0], runs[2][0])


def test_pipelined_collector_equals_the_serial_asynchronous_one(L):
    """collector_pipelined (thin stage and solve stage on streams of their own, PipelinedCollector) against the serial
    asynchronous round with the same lag: the same thinning of the same candidates and the same dense solve, so the
    chains, the counters and the final surrogate agree bit for bit; lag 1 and lag 3 (several rounds in flight)."""
    from surrdamh_b200.sampler import LockstepSampler
    stages = [{"type": "MH", "max_samples": 60, "time_limit": 600, "proposal_std": 1.0, "surrogate_is_updated": True},
              {"type": "DAMH", "max_samples": 240, "time_limit": 600, "proposal_std": 1.0, "surrogate_is_updated": True},
              {"type": "DAMH", "max_samples": 60, "time_limit": 600, "proposal_std": 1.0, "surrogate_is_updated": True}]
    kw = dict(surrogate_type="rbf", surr_solver_parameters={"kernel_type": 1},
              surr_updater_parameters={"kernel_type": 1, "solver_type": "nullspace", "no_keep": 300}, refit_every=20,
              snapshots_per_refit=128, samplers_list=stages, collector_async=True)
    for lag in (1, 3):
        runs = []
        for piped in (True, False):
            smp = LockstepSampler(_conf(collector_pipelined=piped, refit_lag=lag, **kw), no_chains=300)
            smp.run()
            assert smp.collector.fixed and smp.collector.updater.round_failures == 0
            runs.append((smp.block.current_numpy(), smp.block.counters_numpy(), smp.collector.model.coefs.cpu().numpy(),
                         smp.collector.model.centers.cpu().numpy(), smp.collector.no_refits,
                         smp.collector.updater.processed_par))
        for a, b in zip(runs[0], runs[1]):
            assert np.array_equal(a, b), lag
        assert np.array_equal(runs[0][3], runs[0][5])            # the updater took over the last round's center set


def test_pipelined_collector_failing_round_keeps_previous_model(L):
    """The pipelined collector on the GPU with a held set that contains a coincident pair ('direct' solve: exact zero pivot) and
    rounds without snapshots: every round's solve fails, the status log says so, and the model handed to the chains stays the
    healthy one the collector started from -- coefficients AND centers -- while the thin stage's state keeps the updater's rows."""
    from surrdamh_b200 import surrogate_rbf as SR
    from surrdamh_b200.collector import Collector, PipelinedCollector
    rs = np.random.RandomState(8)
    n0, K, C_, quota = 300, 2, 64, 32
    X0 = _prior_points(rs, n0)
    X0[7] = X0[3]
    up = SR.Surrogate_update(2, 1, kernel_type=1, solver_type="direct", no_keep=n0)
    good = SR.Surrogate_update(2, 1, kernel_type=1, solver_type="direct", no_keep=n0)
    Xg = _prior_points(rs, n0)
    good.add_arrays(Xg, _toy_G(Xg))
    model, _ = good.update_device()
    up._par, up._obs = L.dev(X0), L.dev(_toy_G(X0))
    col = Collector(up, "rbf", 2, 1, kernel_type=1, snapshots_per_refit=quota)
    col.model = model
    before = model.coefs.cpu().numpy().copy()
    pc = PipelinedCollector(col, "cuda", lag=1)
    tr, _, _, _ = _random_trace(L, K, C_, 2, 1, seed=1, p_valid=0.0)            # no snapshot at all
    main = torch.cuda.current_stream()
    for b in range(3):
        ev = torch.cuda.Event()
        ev.record()
        pc.submit(b, tr, K, ev, b)
    for b in range(3):
        ev, mo = pc.result(b)
        main.wait_event(ev)
        torch.cuda.synchronize()
        assert np.array_equal(mo.coefs.cpu().numpy(), before), b            # the pack the pipeline started from, round after round
        assert np.array_equal(mo.centers.cpu().numpy(), X0), b
    pc.close()
    log = col.poll_status(wait=True)
    assert [s[1] for s in log[-3:]] == [1, 1, 1] and up.round_failures >= 3
    assert np.array_equal(col.model.coefs.cpu().numpy(), before)
    assert np.array_equal(up.processed_par, X0)
