"""CPU: the numpy oracle against the golden vectors produced by the live reference
(oracle/make_golden.py).  Decisions exact; values within VAL_RTOL."""
import json
import os

import numpy as np
import pytest

from oracle import chain as OC
from oracle import lhs as OL
from oracle import lockstep as OLS
from oracle import problem as OPR
from oracle import solvers as OS
from oracle import surrogate_poly as OP
from oracle import surrogate_rbf as ORB
from tests.golden_util import (CORR_NOISE_PROBLEM, DARCY_PROBLEM, SIMPLE_PROBLEM, close, load, stream_pair)


def test_poly_tables_and_apply():
    g = load("poly")
    for i, (d, m, deg) in enumerate(g["apply_cases"]):
        assert np.array_equal(OP.multi_indices(d, deg), g[f"apply{i}_alpha"])
        assert np.array_equal(OP.hermite_table(deg), g[f"apply{i}_herm"])
        got = OP.PolyApply(d, m).apply([g[f"apply{i}_MAT"], int(deg)], g[f"apply{i}_pts"])
        assert close(got, g[f"apply{i}_GS"])


def test_poly_update():
    g = load("poly")
    for i, (d, m, maxdeg, n1, n2, ill) in enumerate(g["update_cases"]):
        X, Y = g[f"update{i}_X"], g[f"update{i}_Y"]
        up = OP.PolyUpdate(d, m, max_degree=maxdeg)
        up.add_arrays(X[:n1], Y[:n1])
        (M1, deg1), c1 = up.update()
        up.add_arrays(X[n1:], Y[n1:])
        (M2, deg2), c2 = up.update()
        assert [deg1, deg2, c1, c2] == list(g[f"update{i}_deg"])
        # pinv of an ill-conditioned normal matrix amplifies last-bit differences of the BLAS
        # products by cond(A); the well-conditioned cases are held to 1e-10.
        rtol = 1e-4 if ill else 1e-10
        assert close(M1, g[f"update{i}_MAT1"], rtol=rtol)
        assert close(M2, g[f"update{i}_MAT2"], rtol=rtol)


def test_rbf_apply_all_kernels():
    g = load("rbf")
    for kt in range(8):
        got = ORB.RbfApply(3, 2, kernel_type=kt).apply([g["apply_coefs"], g["apply_centers"]], g["apply_pts"])
        assert close(got, g[f"apply_GS_kt{kt}"])
    assert np.array_equal(g["apply_GS_kt6"], g["apply_GS_kt0"])   # kernel 6 is a no-op in the reference
    for tag, (d, m) in {"cfg3": (2, 1), "cfg5": (4, 3)}.items():
        got = ORB.RbfApply(d, m, kernel_type=1).apply([g[f"{tag}_coefs"], g[f"{tag}_centers"]], g[f"{tag}_pts"])
        assert close(got, g[f"{tag}_GS"])


def test_rbf_update():
    g = load("rbf")
    X, Y = g["update_X"], g["update_Y"]
    for i, (kt, minres, keep) in enumerate(g["update_cases"]):
        up = ORB.RbfUpdate(4, 3, kernel_type=int(kt), solver_type="minres" if minres else "direct",
                           no_keep=None if keep < 0 else int(keep))
        up.add_arrays(X[:70], Y[:70])
        (CO1, C1), n1 = up.update()
        up.add_arrays(X[70:], Y[70:])
        (CO2, C2), n2 = up.update()
        assert [n1, n2] == list(g[f"update{i}_n"])
        assert np.array_equal(C1, g[f"update{i}_centers1"]) and np.array_equal(C2, g[f"update{i}_centers2"])
        # coefficient agreement is bounded by cond(system)*eps (1e6..1e10 here), SURVEY section 7 trap 2
        assert close(CO1, g[f"update{i}_coefs1"], rtol=1e-5)
        assert close(CO2, g[f"update{i}_coefs2"], rtol=1e-5)


def test_lhs_toy_and_posteriors():
    g = load("misc")
    lcases = [(4, 2, [5.0, 3.0], [[4, -2], [-2, 4]]), (7, 4, 0.0, 1.0), (16, 3, [1.0, 2.0, 3.0], [1.0, 0.5, 2.0])]
    for i, (n, d, mean, std) in enumerate(lcases):
        assert close(OL.lhs_normal(d, mean, std, n, 0), g[f"lhs{i}"])
    sol = OS.LinElast2Exp()
    got = []
    for p in g["toy_params"]:
        sol.set_parameters(p)
        got.append(sol.get_observations())
    assert close(got, g["toy_G"])
    for tag, (d, kw) in {"simple": (2, SIMPLE_PROBLEM), "darcy": (4, DARCY_PROBLEM), "corrnoise": (3, CORR_NOISE_PROBLEM)}.items():
        prob = OPR.GaussProblem(d, **kw)
        U, G = g[f"post_{tag}_U"], g[f"post_{tag}_G"]
        assert close([prob.log_posterior(U[j], G[j]) for j in range(len(U))], g[f"post_{tag}_val"])
        assert close([prob.log_likelihood(G[j]) for j in range(len(U))], g[f"post_{tag}_ll"])
        assert close([prob.log_prior(U[j]) for j in range(len(U))], g[f"post_{tag}_lp"])


class _Frozen:
    def __init__(self, a, SOL):
        self.a, self.SOL, self.queue = a, SOL, []

    def send_parameters(self, p):
        self.p = p.copy()

    def recv_observations(self):
        return self.a.apply(self.SOL, self.p)

    def send_to_data_collector(self, s):
        self.queue.append(s)


def _run_oracle_chain(kind, d, kw, solver, surrogate, steps, std, streams, updated, init):
    prob = OPR.GaussProblem(d, **kw)
    gen = OPR.StreamGenerators(*streams)
    prop = OPR.RandomWalkProposal(d, proposal_std=std, generator=gen)
    ch = OC.Chain(prob, prop, OS.LocalSolverProxy(solver), kind=kind, initial_sample=init, surrogate=surrogate,
                  surrogate_is_updated=updated, uniform=gen)
    return ch.run(steps)


def _check_chain(ch, g, prefix):
    flags, props, post, pre = ch.trace_arrays()
    assert np.array_equal(flags, g[prefix + "flags"])
    assert np.array_equal(props, g[prefix + "props"])           # u' = u + std*z: exact
    assert [ch.no_accepted, ch.no_rejected, ch.no_prerejected] == list(g[prefix + "counters"])
    assert np.array_equal(ch.current, g[prefix + "final"])
    assert close(post, g[prefix + "post"]) and close(pre, g[prefix + "pre"])
    rows = np.array(ch.data_rows, dtype=float)
    assert np.array_equal(rows[:, :-2], g[prefix + "data_rows"][:, :-2])      # weight + sample exact
    assert close(rows[:, -2:], g[prefix + "data_rows"][:, -2:])


def test_chains_under_streams():
    g = load("chains")
    C, K = g["chains_meta"]
    SOL = [g["poly_MAT"], int(g["poly_deg"])]
    for c in range(C):
        st = stream_pair(100 + 10 * c, K, 2)
        s = _Frozen(OP.PolyApply(2, 1), SOL)
        ch = _run_oracle_chain("MH", 2, SIMPLE_PROBLEM, OS.LinElast2Exp(), s, K, 1.0, st, True, g["simple_init"][c])
        _check_chain(ch, g, f"mh_simple_c{c}_")
        assert np.array_equal(np.array([q.sample for q in s.queue]), g[f"mh_simple_c{c}_snap_par"])
        assert [q.weight for q in s.queue] == list(g[f"mh_simple_c{c}_snap_wei"])
        s = _Frozen(OP.PolyApply(2, 1), SOL)
        ch = _run_oracle_chain("DAMH", 2, SIMPLE_PROBLEM, OS.LinElast2Exp(), s, K, [0.6, 0.9], st, True, g["simple_init"][c])
        _check_chain(ch, g, f"damh_poly_c{c}_")
        assert np.array_equal(np.array([q.sample for q in s.queue]).reshape(-1, 2), g[f"damh_poly_c{c}_snap_par"])
        assert [q.weight for q in s.queue] == list(g[f"damh_poly_c{c}_snap_wei"])
    C, K = g["darcy_meta"]
    for c in range(C):
        st = stream_pair(500 + 10 * c, K, 4)
        truth = OS.SurrogateAsSolver(ORB.RbfApply(4, 3, kernel_type=1), [g["darcy_truth_coefs"], g["darcy_truth_centers"]])
        s = _Frozen(ORB.RbfApply(4, 3, kernel_type=1), [g["darcy_surr_coefs"], g["darcy_surr_centers"]])
        ch = _run_oracle_chain("DAMH", 4, DARCY_PROBLEM, truth, s, K, 0.2, st, False, None)
        _check_chain(ch, g, f"damh_darcy_c{c}_")
    st = stream_pair(900, 800, 3)
    truth = OS.SurrogateAsSolver(ORB.RbfApply(3, 3, kernel_type=1), [g["corr_truth_coefs"], g["corr_truth_centers"]])
    ch = _run_oracle_chain("MH", 3, CORR_NOISE_PROBLEM, truth, None, 800, [0.3, 0.5, 0.2], st, True, None)
    _check_chain(ch, g, "mh_corr_")


SIMPLE_CONF = {
    "no_parameters": 2, "no_observations": 1,
    "problem_parameters": {"prior_mean": [5.0, 3.0], "prior_std": [[4, -2], [-2, 4]], "observations": [-1e-3], "noise_std": 2e-4},
    "samplers_list": [
        {"type": "MH", "max_samples": 100, "time_limit": 60, "proposal_std": 1.0, "surrogate_is_updated": True},
        {"type": "DAMH", "max_samples": 1000, "time_limit": 60, "proposal_std": 1.0, "surrogate_is_updated": True},
        {"type": "DAMH", "max_samples": 10000, "time_limit": 60, "proposal_std": 1.0, "surrogate_is_updated": False}],
    "surrogate_type": "poly", "surr_updater_parameters": {"max_degree": 5}}


def simple_lockstep_generators(N, conf):
    """z from RandomState(seed0+1) (what the reference's own proposal generator would
    draw), u[(stage step), slot] from RandomState(seed0+2+stage) -- make_golden.py."""
    total = sum(s["max_samples"] for s in conf["samplers_list"])
    gens = []
    for c in range(N):
        seed0 = OLS.sampler_seed0(c, N)
        z = np.random.RandomState(seed0 + 1).standard_normal((total, conf["no_parameters"]))
        u = np.vstack([np.random.RandomState(seed0 + 2 + i).uniform(size=(s["max_samples"], 2))
                       for i, s in enumerate(conf["samplers_list"])])
        gens.append(OPR.StreamGenerators(z, u))
    return gens


def test_simple_json_lockstep_schedule():
    """examples/simple.json, N=4, refit every 50 steps: the oracle's lockstep harness
    against the reference's own classes driven under the same schedule."""
    g = load("simple_lockstep")
    N, R, total = g["meta"]
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        stages, shared = OLS.run_lockstep(SIMPLE_CONF, N, OS.LinElast2Exp, apply_obj=OP.PolyApply(2, 1),
                                          updater=OP.PolyUpdate(2, 1, max_degree=5), refit_every=int(R),
                                          initial_samples=g["init"], generators=simple_lockstep_generators(N, SIMPLE_CONF))
    hist = g["history"]
    assert [(a, b) for a, b in shared.history] == [(int(r[0]), int(r[1])) for r in hist]
    for i, chains in enumerate(stages):
        for c, ch in enumerate(chains):
            flags = ch.trace_arrays()[0]
            assert np.array_equal(flags, g[f"st{i}_c{c}_flags"]), (i, c)
            assert [ch.no_accepted, ch.no_rejected, ch.no_prerejected] == list(g[f"st{i}_c{c}_counters"])
            assert np.array_equal(ch.current, g[f"st{i}_c{c}_final"])
            rows = np.array(ch.data_rows, dtype=float)
            assert np.array_equal(rows[:, :3], g[f"st{i}_c{c}_data"][:, :3])
            assert close(rows[:, 3:], g[f"st{i}_c{c}_data"][:, 3:], rtol=1e-9)


def test_manifest_lists_all_fixtures():
    with open(os.path.join(os.path.dirname(__file__), "golden", "MANIFEST.json")) as f:
        man = json.load(f)
    for name in man["files"]:
        assert os.path.isfile(os.path.join(os.path.dirname(__file__), "golden", name))
