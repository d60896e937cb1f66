"""GPU parity tests of the host driver (LockstepSampler + Collector): the reference's simple.json run
for N = 4 samplers under the deterministic lockstep schedule (golden vectors produced by driving the
live reference's own classes, oracle/make_golden.py:gen_simple_lockstep), the host-plugin (split)
path, and the full GPU pipeline (GPU refits) against the numpy oracle in distribution."""
import contextlib
import io
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import lockstep as OLS
from oracle import solvers as OS
from oracle import surrogate_poly as OP
from tests.golden_util import close, load
from tests.test_oracle_golden import SIMPLE_CONF, simple_lockstep_generators

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _conf(extra=None, **kw):
    from surrdamh_b200.configuration import Configuration
    c = dict(SIMPLE_CONF, solver_module_path=os.path.join(ROOT, "surrdamh_b200", "solvers.py"), solver_module_name="solvers",
             solver_init_name="Solver_linela2exp_local", solver_parameters={}, no_solvers=1)
    c.update(extra or {})
    return Configuration(kw.pop("no_samplers", 4), conf=c)


def _streams_for(N, conf, device):
    """The (z, u) streams of the golden run, as the sampler's `streams` callable."""
    from surrdamh_b200 import _lib as L
    gens = simple_lockstep_generators(N, conf)
    z = np.stack([g.z for g in gens], axis=-1)            # [total][d][C]
    u = np.stack([g.u for g in gens], axis=-1)            # [total][2][C]
    offs = np.cumsum([0] + [s["max_samples"] for s in conf["samplers_list"]])

    def streams(stage, step0, K):
        a = offs[stage] + step0
        return L.dev(np.ascontiguousarray(z[a:a + K])), L.dev(np.ascontiguousarray(u[a:a + K]))
    return streams


@pytest.mark.parametrize("device_solver", [True, False, "pool"])
def test_simple_json_lockstep_bit_identical_to_reference(device_solver):
    """simple.json, 4 chains, refit every 50 steps.  Chains, proposal / uniform streams and the refit
    schedule are those of the golden run; the surrogate coefficients installed at each refit are the
    reference's own (SOL_k of the fixture), so every accept / reject decision of all three stages must
    be bit-identical.  device_solver=False routes the true model through the host plugin interface
    (set_parameters / get_observations), the reference's own data path."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from surrdamh_b200.sampler import LockstepSampler
    g = load("simple_lockstep")
    N, R, total = [int(x) for x in g["meta"]]
    use_pool = device_solver == "pool"       # no_solvers = 2: the host G goes through the pool of solver processes
    if use_pool:
        device_solver = False
    conf = _conf({"refit_every": R, "device_solver": device_solver, "no_solvers": 2 if use_pool else 1})
    flags = {}

    def sink(stage, step0, K, tr):
        flags.setdefault(stage, []).append(tr.flags[:K].cpu().numpy().copy())
    smp = LockstepSampler(conf, no_chains=N, streams=_streams_for(N, SIMPLE_CONF, "cuda"), trace_sink=sink)
    assert np.array_equal(smp.block.current_numpy(), g["init"])              # LHS start, row c for chain c
    hist = g["history"]

    def hook(par, obs, k):
        return [g["SOL_%d" % k], int(hist[k][2])]          # the reference's own coefficients and degree of refit k
    smp.collector.refit_hook = hook
    res = smp.run()
    steps_at = [a for a, _ in smp.collector.history]
    cum = list(np.cumsum([b for _, b in smp.collector.history]))
    assert steps_at == [int(r[0]) for r in hist] and cum == [int(r[1]) for r in hist]       # same schedule, same snapshot counts
    for i, r in enumerate(res):
        f = np.concatenate(flags[i], axis=0)
        for c in range(N):
            assert np.array_equal(f[:, c], g[f"st{i}_c{c}_flags"]), (i, c)
            assert list(r.counters[c][:3]) == list(g[f"st{i}_c{c}_counters"])
        if i == len(res) - 1:
            for c in range(N):
                assert np.array_equal(smp.block.current_numpy()[c], g[f"st{i}_c{c}_final"])
    if not device_solver:
        assert smp.no_G_calls_host > 0
    assert (smp.pool is not None) == use_pool
    if smp.pool is not None:
        smp.pool.close()


def test_full_gpu_pipeline_matches_oracle_in_distribution():
    """Everything on the GPU (Philox, fused kernel, GPU polynomial refits), 2048 chains: acceptance
    statistics and posterior moments of the last (frozen-surrogate) stage agree with the numpy oracle's
    lockstep run within Monte-Carlo error."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from surrdamh_b200.sampler import LockstepSampler
    conf_d = dict(SIMPLE_CONF)
    conf_d["samplers_list"] = [dict(s) for s in SIMPLE_CONF["samplers_list"]]
    conf_d["samplers_list"][2]["max_samples"] = 1500
    C_ = 2048
    conf = _conf({"samplers_list": conf_d["samplers_list"], "refit_every": 50, "initial_sample_type": "prior_mean",
                  "snapshots_per_refit": 2000, "seed": 5})
    smp = LockstepSampler(conf, no_chains=C_, moments=True)
    smp.run_stage(0, conf.list_alg[0])
    smp.run_stage(1, conf.list_alg[1])
    smp.block.moments.zero_()
    r2 = smp.run_stage(2, conf.list_alg[2])
    mom = smp.block.moments.sum(dim=1).cpu().numpy() / (C_ * r2.steps)
    mean_g = mom[:2]
    cov_g = np.array([[mom[2], mom[3]], [mom[3], mom[4]]]) - np.outer(mean_g, mean_g)
    acc_g = r2.counters[:, 0].sum() / (C_ * r2.steps)
    pre_g = r2.counters[:, 2].sum() / (C_ * r2.steps)
    # oracle: 24 chains, same stages, default numpy generators
    n_or = 24
    with contextlib.redirect_stdout(io.StringIO()):
        stages, shared = OLS.run_lockstep(conf_d, n_or, OS.LinElast2Exp, apply_obj=OP.PolyApply(2, 1),
                                          updater=OP.PolyUpdate(2, 1, max_degree=5), refit_every=50,
                                          initial_samples=[np.array([5.0, 3.0])] * n_or)
    last = stages[2]
    samples, acc_o, pre_o = [], 0, 0
    for ch in last:
        f, props, _, _ = ch.trace_arrays()
        cur = np.array([5.0, 3.0])
        acc_o += ch.no_accepted
        pre_o += ch.no_prerejected
        rows = np.array(ch.data_rows)
        samples.append(np.repeat(rows[:, 1:3], rows[:, 0].astype(int), axis=0))
    S = np.vstack(samples)
    n_steps = sum(len(ch.trace) for ch in last)
    acc_o, pre_o = acc_o / n_steps, pre_o / n_steps
    mean_o, cov_o = S.mean(axis=0), np.cov(S.T)
    # the posterior is a thin curved ridge: compare with tolerances a few times the oracle's own MC error
    assert abs(acc_g - acc_o) < 0.05 and abs(pre_g - pre_o) < 0.08, (acc_g, acc_o, pre_g, pre_o)
    assert np.all(np.abs(mean_g - mean_o) < 0.5), (mean_g, mean_o)
    assert np.all(np.abs(np.sqrt(np.diag(cov_g)) - np.sqrt(np.diag(cov_o))) < 0.5), (cov_g, cov_o)
    assert smp.collector.no_refits == 2 + 20


def test_csv_traces_match_reference_files(tmp_path):
    """The on-disk layout of classes_SAMPLER.py written from the GPU trace: data/ (compressed chain rows),
    raw_data/ (per-step labels), notes/, DAMH_accepted/ and DAMH_rejected/ equal what the reference's own
    Algorithm objects wrote in the golden run (weights and samples exact, log-posteriors to 1e-9)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import csv
    from surrdamh_b200.sampler import LockstepSampler
    from surrdamh_b200.trace_csv import CsvTraceWriter
    g = load("simple_lockstep")
    N, R, total = [int(x) for x in g["meta"]]
    hist = g["history"]
    conf = _conf({"refit_every": R, "saved_samples_name": "csvtest"})
    sink = CsvTraceWriter("csvtest", root=str(tmp_path))
    smp = LockstepSampler(conf, no_chains=N, streams=_streams_for(N, SIMPLE_CONF, "cuda"), trace_sink=sink)
    smp.collector.refit_hook = lambda par, obs, k: [g["SOL_%d" % k], int(hist[k][2])]
    res = smp.run()
    base = tmp_path / "saved_samples" / "csvtest"
    code = {"prerejected": 0, "accepted": 1, "rejected": 2}
    for i, st in enumerate(SIMPLE_CONF["samplers_list"]):
        for c in range(N):
            name = "alg%03d%s_rank%d" % (i, st["type"], c)
            rows = np.array([[float(x) for x in r] for r in csv.reader(open(base / "data" / (name + ".csv")))])
            want = g[f"st{i}_c{c}_data"]
            assert rows.shape == want.shape, (i, c)
            assert np.array_equal(rows[:, :3], want[:, :3])
            assert close(rows[:, 3:], want[:, 3:], rtol=1e-9), (i, c)
            raw = [r for r in csv.reader(open(base / "raw_data" / (name + ".csv")))]
            assert np.array_equal(np.array([code[r[0]] for r in raw], dtype=np.uint8), g[f"st{i}_c{c}_flags"])
            notes = [r for r in csv.reader(open(base / "notes" / (name + ".csv")))]
            assert notes[0][0] == "name" and notes[1][0] == name
            assert [int(notes[1][2]), int(notes[1][3]), int(notes[1][4])] == list(g[f"st{i}_c{c}_counters"])
            if st["type"] == "DAMH":
                acc = [r for r in csv.reader(open(base / "DAMH_accepted" / (name + ".csv")))]
                rej = [r for r in csv.reader(open(base / "DAMH_rejected" / (name + ".csv")))]
                assert len(acc) == g[f"st{i}_c{c}_counters"][0] and len(rej) == g[f"st{i}_c{c}_counters"][1]


def test_stage_options_initial_sample_proposal_differs_time_limit_and_errors():
    """process_SAMPLER.py:53-64 stage keys: per-stage `initial_sample`, `proposal_differs` (per-chain proposal_std indexed by
    the global chain), `time_limit`; a DAMH stage before any surrogate fails like the reference (TypeError on SOL = None)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from surrdamh_b200.sampler import LockstepSampler
    C_ = 8
    stds = [0.1 * (c + 1) for c in range(C_)]
    conf = _conf({"samplers_list": [
        {"type": "MH", "max_samples": 30, "time_limit": 60, "proposal_std": stds, "proposal_differs": True, "surrogate_is_updated": True,
         "initial_sample": [6.5, 3.5]},
        {"type": "DAMH", "max_samples": 10 ** 9, "time_limit": 0.3, "proposal_std": 0.5, "surrogate_is_updated": False}],
        "initial_sample_type": "prior_mean", "refit_every": 10, "seed": 4})
    seen = {}

    def sink(stage, step0, K, tr):
        if stage == 0 and step0 == 0:
            seen["prop0"] = tr.prop[0].cpu().numpy().copy()
    smp = LockstepSampler(conf, no_chains=C_, trace_sink=sink)
    z, _ = smp.block.philox_streams(1, step0=0)
    r0 = smp.run_stage(0, conf.list_alg[0])
    # first proposal of chain c = initial_sample + std_c * z_c   (stage sub-stream 0 == the block's default stream id)
    zn = z.cpu().numpy()[0]
    want = np.array([6.5, 3.5])[:, None] + np.array(stds)[None, :] * zn
    assert np.array_equal(seen["prop0"], want)
    assert r0.steps == 30 and r0.refits == 3
    r1 = smp.run_stage(1, conf.list_alg[1])                      # stops on the time limit, between blocks
    assert 0 < r1.steps < 10 ** 9 and r1.seconds < 5.0
    assert int(r1.counters[:, :3].sum()) == r1.steps * C_
    conf2 = _conf({"samplers_list": [{"type": "DAMH", "max_samples": 5, "time_limit": 60, "proposal_std": 0.5,
                                      "surrogate_is_updated": True}], "initial_sample_type": "prior_mean"})
    with pytest.raises(TypeError):
        LockstepSampler(conf2, no_chains=4).run()


def test_minres_initial_iteration_and_iteration_limit():
    """surr_updater_parameters of the reference's RBF updater: `initial_iteration` (x0 of MINRES) and the iteration limit
    (scipy's info = maxiter when it is hit)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import scipy.sparse.linalg as splin
    from oracle import surrogate_rbf as ORB
    from surrdamh_b200 import surrogate_rbf as SR
    rs = np.random.RandomState(31)
    X, Y = rs.standard_normal((150, 3)), rs.standard_normal((150, 1))
    ref = ORB.RbfUpdate(3, 1, kernel_type=1, solver_type="direct")
    ref.add_arrays(X, Y)
    (COd, _), _ = ref.update()
    S, b = ref.last_system, ref.last_rhs[:, 0]
    x0 = COd[:, 0] * (1 + 1e-3 * rs.standard_normal(COd.shape[0]))       # a good starting point: few iterations
    up = SR.Surrogate_update(3, 1, kernel_type=1, solver_type="minres", initial_iteration=x0)
    up.add_arrays(X, Y)
    (CO, _), _ = up.update()
    want, info_ref = splin.minres(S, b, x0=x0, rtol=1e-6)
    assert up.last_info[0] == info_ref == 0
    assert np.linalg.norm(CO[:, 0] - want) <= 1e-2 * np.linalg.norm(want)
    assert np.linalg.norm(S @ CO[:, 0] - b) <= 2.0 * np.linalg.norm(S @ want - b) + 1e-12
    up2 = SR.Surrogate_update(3, 1, kernel_type=1, solver_type="minres", max_iterations=7)
    up2.add_arrays(X, Y)
    up2.update()
    assert up2.last_info[0] == 7 and up2.last_info[1] == 7 and up2.last_info[2] == 6     # scipy: info = maxiter, istop = 6
