"""TEST INFRASTRUCTURE -- generate tests/golden/*.npz from the LIVE reference.

Run in the build container only (needs /root/reference):
    python -m oracle.make_golden
Every expected value below is produced by the unmodified reference modules
(imported through oracle/ref_import.py); the oracle restatement is NOT used to
produce any expectation.  Inputs are derived from fixed numpy RandomState seeds
so the fixtures stay small; where an input is cheap it is stored as well.

Environment the fixtures were generated with is recorded in golden/MANIFEST.json
(numpy / scipy / BLAS versions): transcendental functions (exp, log, pow) and
BLAS reductions differ at the ulp level between machines, so value comparisons
in the tests carry a stated tolerance; decisions (flags, counters) are exact.
"""
import contextlib
import csv
import io
import json
import os
import sys
import tempfile
import threading
import warnings

import numpy as np

from . import ref_import as R

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

SIMPLE_PROBLEM = dict(prior_mean=[5.0, 3.0], prior_std=[[4, -2], [-2, 4]], noise_std=2e-4,
                      no_observations=1, observations=[-1e-3])
DARCY_PROBLEM = dict(prior_mean=0.0, prior_std=1.0, noise_std=[0.2, 0.1, 0.1], no_observations=3,
                     observations=[10.24925199, -6.07103574, -4.17821625])
CORR_NOISE_PROBLEM = dict(prior_mean=[0.5, -0.5, 0.0], prior_std=[1.0, 2.0, 0.5],
                          noise_std=[[0.04, 0.01, 0.0], [0.01, 0.09, -0.02], [0.0, -0.02, 0.0625]],
                          no_observations=3, observations=[1.0, -2.0, 0.5])


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        yield


class Snap:
    def __init__(self, sample, G_sample, weight=1):
        self.sample, self.G_sample, self.weight = sample, G_sample, weight


def smooth_map(X, W, shift):
    return shift + 2.0 * np.sin(X @ W)


# --------------------------------------------------------------------------- unit fixtures
def gen_poly(ns):
    rs = np.random.RandomState(11)
    out = {}
    cases = [(2, 1, 5), (2, 1, 6), (4, 3, 5), (3, 2, 4), (1, 1, 3), (6, 2, 2), (2, 1, 1)]
    out["apply_cases"] = np.array(cases)
    for i, (d, m, deg) in enumerate(cases):
        with quiet():
            ra = ns.poly.Surrogate_apply(d, m)
        P = ns.poly.generate_polynomials_degree(d, deg).shape[0]
        MAT = rs.standard_normal((P, m))
        pts = rs.standard_normal((33, d)) * 1.5 + 0.5
        out[f"apply{i}_MAT"], out[f"apply{i}_pts"] = MAT, pts
        out[f"apply{i}_GS"] = ra.apply([MAT, deg], pts)
        out[f"apply{i}_alpha"] = ns.poly.generate_polynomials_degree(d, deg)
        out[f"apply{i}_herm"] = ns.poly.hermite_poly_normalized(deg)
    # update: well conditioned (X ~ N(0,I)) in two batches, and the toy-problem-shaped ill-conditioned one
    ucases = [(2, 1, 5, 40, 600, 0), (4, 3, 5, 60, 900, 0), (3, 2, 3, 20, 100, 0), (2, 1, 5, 50, 700, 1)]
    out["update_cases"] = np.array(ucases)
    for i, (d, m, maxdeg, n1, n2, ill) in enumerate(ucases):
        X = rs.standard_normal((n1 + n2, d))
        if ill:
            X = X * 0.7 + np.array([8.0, 3.0])
            sol = ns.solvers.Solver_linela2exp_local()
            Y = []
            for x in X:
                sol.set_parameters(x)
                Y.append([sol.get_observations()])
            Y = np.array(Y)
        else:
            Y = np.sin(X @ rs.standard_normal((d, m)))
        with quiet():
            ru = ns.poly.Surrogate_update(d, m, max_degree=maxdeg)
            ru.add_data([Snap(X[j], Y[j]) for j in range(n1)])
            (MAT1, deg1), cnt1 = ru.update()
            ru.add_data([Snap(X[j], Y[j]) for j in range(n1, n1 + n2)])
            (MAT2, deg2), cnt2 = ru.update()
        out[f"update{i}_X"], out[f"update{i}_Y"] = X, Y
        out[f"update{i}_MAT1"], out[f"update{i}_MAT2"] = MAT1, MAT2
        out[f"update{i}_deg"] = np.array([deg1, deg2, cnt1, cnt2])
    np.savez_compressed(os.path.join(GOLDEN, "poly.npz"), **out)


def gen_rbf(ns):
    rs = np.random.RandomState(12)
    out = {}
    d, m, n = 3, 2, 50
    C = rs.standard_normal((n, d))
    CO = rs.standard_normal((n + d + 1, m))
    pts = rs.standard_normal((17, d))
    out["apply_centers"], out["apply_coefs"], out["apply_pts"] = C, CO, pts
    for kt in range(8):
        out[f"apply_GS_kt{kt}"] = ns.rbf.Surrogate_apply(d, m, kernel_type=kt).apply([CO, C], pts)
    # the two named shapes
    for tag, (d, m, n) in {"cfg3": (2, 1, 300), "cfg5": (4, 3, 200)}.items():
        C = rs.standard_normal((n, d))
        CO = rs.standard_normal((n + d + 1, m))
        pts = rs.standard_normal((64, d))
        out[f"{tag}_centers"], out[f"{tag}_coefs"], out[f"{tag}_pts"] = C, CO, pts
        out[f"{tag}_GS"] = ns.rbf.Surrogate_apply(d, m, kernel_type=1).apply([CO, C], pts)
    # update
    ucases = [(1, "direct", -1), (1, "minres", -1), (0, "direct", -1), (2, "direct", -1), (4, "direct", -1),
              (5, "direct", -1), (3, "direct", -1), (7, "direct", -1), (1, "direct", 60), (1, "direct", 1000)]
    out["update_cases"] = np.array([(kt, 0 if s == "direct" else 1, keep) for kt, s, keep in ucases])
    d, m = 4, 3
    X = rs.standard_normal((160, d))
    Y = smooth_map(X, rs.standard_normal((d, m)), np.array([1.0, -1.0, 0.5]))
    out["update_X"], out["update_Y"] = X, Y
    for i, (kt, solver, keep) in enumerate(ucases):
        ru = ns.rbf.Surrogate_update(d, m, kernel_type=kt, solver_type=solver, no_keep=None if keep < 0 else keep)
        ru.add_data([Snap(X[j], Y[j]) for j in range(70)])
        (CO1, C1), n1 = ru.update()
        ru.add_data([Snap(X[j], Y[j]) for j in range(70, 160)])
        (CO2, C2), n2 = ru.update()
        out[f"update{i}_coefs1"], out[f"update{i}_centers1"] = CO1, C1
        out[f"update{i}_coefs2"], out[f"update{i}_centers2"] = CO2, C2
        out[f"update{i}_n"] = np.array([n1, n2])
    np.savez_compressed(os.path.join(GOLDEN, "rbf.npz"), **out)


def gen_misc(ns):
    out = {}
    lcases = [(4, 2, [5.0, 3.0], [[4, -2], [-2, 4]]), (7, 4, 0.0, 1.0), (16, 3, [1.0, 2.0, 3.0], [1.0, 0.5, 2.0])]
    for i, (n, d, mean, std) in enumerate(lcases):
        out[f"lhs{i}"] = ns.lhs.lhs_normal(d, mean, std, n, 0)
    rs = np.random.RandomState(13)
    P = rs.standard_normal((500, 2)) * 2 + np.array([5.0, 3.0])
    sol = ns.solvers.Solver_linela2exp_local()
    G = []
    for p in P:
        sol.set_parameters(p)
        G.append(sol.get_observations())
    out["toy_params"], out["toy_G"] = P, np.array(G)
    # log-posterior samples for the three problem shapes
    for tag, (d, m, kw) in {"simple": (2, 1, SIMPLE_PROBLEM), "darcy": (4, 3, DARCY_PROBLEM),
                            "corrnoise": (3, 3, CORR_NOISE_PROBLEM)}.items():
        prob = ns.cS.Problem_Gauss(d, **kw)
        U = rs.standard_normal((200, d)) * 1.5 + np.array(prob.prior_mean)
        Gs = rs.standard_normal((200, m)) * 0.5 + np.array(kw["observations"])
        out[f"post_{tag}_U"], out[f"post_{tag}_G"] = U, Gs
        out[f"post_{tag}_val"] = np.array([prob.get_log_posterior(U[j], Gs[j]) for j in range(200)])
        out[f"post_{tag}_ll"] = np.array([prob.get_log_likelihood(Gs[j]) for j in range(200)])
        out[f"post_{tag}_lp"] = np.array([prob.get_log_prior(U[j]) for j in range(200)])
    np.savez_compressed(os.path.join(GOLDEN, "misc.npz"), **out)


# --------------------------------------------------------------------------- chains under explicit streams
class SolverProxy:
    def __init__(self, solver):
        self.solver = solver

    def send_parameters(self, p):
        self.solver.set_parameters(p)

    def recv_observations(self):
        return self.solver.get_observations()


class RefSurrogateAsSolver:
    def __init__(self, apply_obj, SOL):
        self.a, self.SOL = apply_obj, SOL

    def set_parameters(self, p):
        self.p = np.array(p)

    def get_observations(self):
        return self.a.apply(self.SOL, self.p)[0, :]


class FrozenSurrogateProxy:
    def __init__(self, apply_obj, SOL, queue=None):
        self.a, self.SOL, self.queue = apply_obj, SOL, [] if queue is None else queue

    def send_parameters(self, p):
        self.p = p.copy()

    def recv_observations(self):
        return self.a.apply(self.SOL, self.p)

    def send_to_data_collector(self, snap):
        self.queue.append(snap)


def stream_pair(seed, steps, d):
    z = np.random.RandomState(seed).standard_normal((steps, d))
    u = np.random.RandomState(seed + 1).uniform(size=(steps, 2))
    return z, u


def read_rows(path):
    with open(path) as f:
        return [row for row in csv.reader(f)]


def run_ref_chain(ns, kind, d, prob_kw, solver, surrogate, steps, std, streams, updated, init):
    """One reference Algorithm_MH / Algorithm_DAMH run; returns the per-step record."""
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    os.chdir(tmp)
    try:
        prob = ns.cS.Problem_Gauss(d, **prob_kw, saved_samples_name="g")
        prop = ns.cS.Proposal_GaussRandomWalk(d, proposal_std=std, seed=0)
        cls = ns.cS.Algorithm_MH if kind == "MH" else ns.cS.Algorithm_DAMH
        alg = cls(prob, prop, SolverProxy(solver), Surrogate=surrogate, surrogate_is_updated=updated,
                  max_samples=steps, name="a", seed=0, save_raw_data=True, initial_sample=init)
        R.inject_streams(alg, prop, R.Streams(*streams))
        log = []
        inner = alg.compute_posterior

        def recording(sample, G):
            v = inner(sample, G)
            log.append(v)
            return v

        alg.compute_posterior = recording      # instance attribute; the reference file is untouched
        with quiet():
            alg.run()
        raw = read_rows("saved_samples/g/raw_data/a.csv")
        data = read_rows("saved_samples/g/data/a.csv")
    finally:
        os.chdir(cwd)
    code = {"prerejected": 0, "accepted": 1, "rejected": 2}
    flags = np.array([code[r[0]] for r in raw], dtype=np.uint8)
    props = np.array([[float(x) for x in r[1:]] for r in raw])
    post = np.full(steps, np.nan)
    pre = np.full(steps, np.nan)
    pos = 1 + (1 if kind == "DAMH" else 0)       # prepare(): posterior, then DAMH prologue pre-posterior
    for k in range(steps):
        if kind == "MH":
            post[k] = log[pos]
            pos += 1
        else:
            pre[k] = log[pos + 1]
            pos += 2
            if flags[k] != 0:
                post[k] = log[pos]
                pos += 1
    assert pos == len(log)
    rows = np.array([[float(x) for x in r] for r in data])
    return dict(flags=flags, props=props, post=post, pre=pre, data_rows=rows,
                counters=np.array([alg.no_accepted, alg.no_rejected, alg.no_prerejected]),
                final=np.array(alg.current_sample), post0=log[0])


def gen_chains(ns):
    out = {}
    rs = np.random.RandomState(14)
    # --- poly surrogate for the toy problem: least squares on prior-spread samples (deg 5)
    X = rs.standard_normal((1500, 2)) @ np.linalg.cholesky(np.array([[4.0, -2.0], [-2.0, 4.0]])).T + np.array([5.0, 3.0])
    toy = ns.solvers.Solver_linela2exp_local()
    Y = []
    for x in X:
        toy.set_parameters(x)
        Y.append([toy.get_observations()])
    with quiet():
        pu = ns.poly.Surrogate_update(2, 1, max_degree=5)
        pu.add_data([Snap(X[j], np.array(Y[j])) for j in range(len(X))])
        (MAT, deg), _ = pu.update()
    out["poly_MAT"], out["poly_deg"] = MAT, np.array(deg)
    C, K = 4, 1200
    init = ns.lhs.lhs_normal(2, SIMPLE_PROBLEM["prior_mean"], SIMPLE_PROBLEM["prior_std"], C, 0)
    out["simple_init"] = init
    for c in range(C):
        st = stream_pair(100 + 10 * c, K, 2)
        with quiet():
            pa = ns.poly.Surrogate_apply(2, 1)
        q = []
        r = run_ref_chain(ns, "MH", 2, SIMPLE_PROBLEM, ns.solvers.Solver_linela2exp_local(),
                          FrozenSurrogateProxy(pa, [MAT, deg], q), K, 1.0, st, True, init[c])
        for k, v in r.items():
            out[f"mh_simple_c{c}_{k}"] = v
        out[f"mh_simple_c{c}_snap_par"] = np.array([s.sample for s in q])
        out[f"mh_simple_c{c}_snap_obs"] = np.array([np.atleast_1d(s.G_sample) for s in q])
        out[f"mh_simple_c{c}_snap_wei"] = np.array([s.weight for s in q])
        q = []
        r = run_ref_chain(ns, "DAMH", 2, SIMPLE_PROBLEM, ns.solvers.Solver_linela2exp_local(),
                          FrozenSurrogateProxy(pa, [MAT, deg], q), K, [0.6, 0.9], st, True, init[c])
        for k, v in r.items():
            out[f"damh_poly_c{c}_{k}"] = v
        out[f"damh_poly_c{c}_snap_par"] = np.array([s.sample for s in q]).reshape(-1, 2)
        out[f"damh_poly_c{c}_snap_obs"] = np.array([np.atleast_1d(s.G_sample) for s in q]).reshape(-1, 1)
        out[f"damh_poly_c{c}_snap_wei"] = np.array([s.weight for s in q])
    out["chains_meta"] = np.array([C, K])
    # --- Darcy shape: cubic RBF surrogate (80 centers), cubic RBF "truth" (300 centers)
    d, m = 4, 3
    Xc = rs.standard_normal((300, d))
    W = rs.standard_normal((d, m)) * 0.6
    shift = np.array(DARCY_PROBLEM["observations"]) - np.array([0.3, -0.2, 0.1])
    tu = ns.rbf.Surrogate_update(d, m, kernel_type=1, solver_type="direct")
    tu.add_data([Snap(Xc[j], smooth_map(Xc[j], W, shift)) for j in range(300)])
    (COt, Ct), _ = tu.update()
    su = ns.rbf.Surrogate_update(d, m, kernel_type=1, solver_type="direct")
    su.add_data([Snap(Xc[j], smooth_map(Xc[j], W, shift)) for j in range(80)])
    (COs, Cs), _ = su.update()
    out["darcy_truth_coefs"], out["darcy_truth_centers"] = COt, Ct
    out["darcy_surr_coefs"], out["darcy_surr_centers"] = COs, Cs
    Kd = 1000
    for c in range(C):
        st = stream_pair(500 + 10 * c, Kd, d)
        truth = RefSurrogateAsSolver(ns.rbf.Surrogate_apply(d, m, kernel_type=1), [COt, Ct])
        surr = FrozenSurrogateProxy(ns.rbf.Surrogate_apply(d, m, kernel_type=1), [COs, Cs])
        r = run_ref_chain(ns, "DAMH", d, DARCY_PROBLEM, truth, surr, Kd, 0.2, st, False, None)
        for k, v in r.items():
            out[f"damh_darcy_c{c}_{k}"] = v
    out["darcy_meta"] = np.array([C, Kd])
    # --- correlated noise covariance, uncorrelated prior, MH with the RBF truth restricted to 3 params
    d, m = 3, 3
    Xc3 = rs.standard_normal((120, d))
    W3 = rs.standard_normal((d, m)) * 0.5
    tu3 = ns.rbf.Surrogate_update(d, m, kernel_type=1, solver_type="direct")
    tu3.add_data([Snap(Xc3[j], smooth_map(Xc3[j], W3, np.array([1.0, -2.0, 0.5]))) for j in range(120)])
    (CO3, C3), _ = tu3.update()
    out["corr_truth_coefs"], out["corr_truth_centers"] = CO3, C3
    st = stream_pair(900, 800, d)
    truth = RefSurrogateAsSolver(ns.rbf.Surrogate_apply(d, m, kernel_type=1), [CO3, C3])
    r = run_ref_chain(ns, "MH", d, CORR_NOISE_PROBLEM, truth, None, 800, [0.3, 0.5, 0.2], st, True, None)
    for k, v in r.items():
        out[f"mh_corr_{k}"] = v
    np.savez_compressed(os.path.join(GOLDEN, "chains.npz"), **out)


# --------------------------------------------------------------------------- simple.json, N=4, lockstep schedule
class Turnstile:
    """Runs N reference chains (threads) strictly one step at a time, in chain order.

    Each chain blocks inside its proposal generator stub -- the first call of every
    iteration (classes_SAMPLER.py:150,182) -- until the scheduler lets it take one step.
    """

    def __init__(self, n):
        self.go = [threading.Semaphore(0) for _ in range(n)]
        self.parked = [threading.Semaphore(0) for _ in range(n)]

    def park(self, c):          # chain side: "I am at a step boundary"
        self.parked[c].release()
        self.go[c].acquire()

    def advance(self, c):       # scheduler side: run chain c up to its next boundary (or its end)
        self.go[c].release()
        self.parked[c].acquire()


class ParkingNormal:
    def __init__(self, streams, turnstile, c):
        self.inner, self.t, self.c = streams.normal_stub(), turnstile, c

    def normal(self, mean, std):
        self.t.park(self.c)
        return self.inner.normal(mean, std)


class SharedRef:
    def __init__(self, apply_factory, updater):
        self.apply_factory, self.updater = apply_factory, updater
        self.SOL, self.queue, self.history = None, [], []


class SharedRefProxy:
    def __init__(self, shared):
        self.s, self.a = shared, shared.apply_factory()

    def send_parameters(self, p):
        self.p = p.copy()

    def recv_observations(self):
        return self.a.apply(self.s.SOL, self.p)

    def send_to_data_collector(self, snap):
        self.s.queue.append(snap)


def gen_simple_lockstep(ns, refit_every=50):
    with open(os.path.join(ns.root, "examples", "simple.json")) as f:
        conf = json.load(f)
    N, d, m = 4, conf["no_parameters"], conf["no_observations"]
    pp = conf["problem_parameters"]
    total = sum(s["max_samples"] for s in conf["samplers_list"])
    init = ns.lhs.lhs_normal(d, pp["prior_mean"], pp["prior_std"], N, 0)
    with quiet():
        shared = SharedRef(lambda: ns.poly.Surrogate_apply(d, m),
                           ns.poly.Surrogate_update(d, m, **conf["surr_updater_parameters"]))
    seed0 = [max(1000, N + 2) * c for c in range(N)]
    zs = [np.random.RandomState(s + 1).standard_normal((total, d)) for s in seed0]
    current = [init[c] for c in range(N)]
    out, gstep = {}, 0
    cwd, tmp = os.getcwd(), tempfile.mkdtemp()
    os.chdir(tmp)
    try:
        prob = ns.cS.Problem_Gauss(d, noise_std=pp["noise_std"], prior_mean=pp["prior_mean"], prior_std=pp["prior_std"],
                                   no_observations=m, observations=pp["observations"], saved_samples_name="s")
        props = [ns.cS.Proposal_GaussRandomWalk(d, seed=seed0[c] + 1) for c in range(N)]
        for i, st in enumerate(conf["samplers_list"]):
            K = st["max_samples"]
            gate = Turnstile(N)
            algs, threads, logs = [], [], []
            for c in range(N):
                props[c].set_covariance(proposal_std=st["proposal_std"])
                u = np.random.RandomState(seed0[c] + 2 + i).uniform(size=(K, 2))
                cls = ns.cS.Algorithm_MH if st["type"] == "MH" else ns.cS.Algorithm_DAMH
                alg = cls(prob, props[c], SolverProxy(ns.solvers.Solver_linela2exp_local()),
                          Surrogate=SharedRefProxy(shared), surrogate_is_updated=st["surrogate_is_updated"],
                          initial_sample=current[c], max_samples=K, time_limit=float("inf"), save_raw_data=True,
                          name="alg%03d%s_rank%d" % (i, st["type"], c), seed=seed0[c] + 2 + i)
                streams = R.Streams(zs[c][gstep:gstep + K], u)
                alg._Algorithm_PARENT__generator = streams.uniform_stub()
                props[c]._Proposal_GaussRandomWalk__generator = ParkingNormal(streams, gate, c)
                log = []
                inner = alg.compute_posterior
                alg.compute_posterior = (lambda inner, log: (lambda s, G: (log.append(inner(s, G)), log[-1])[1]))(inner, log)
                algs.append(alg)
                logs.append(log)

                def body(alg=alg, c=c):
                    alg.run()           # stdout is silenced by the main thread for the whole stage
                    gate.parked[c].release()
                threads.append(threading.Thread(target=body))
            with quiet():
                for c in range(N):          # prepare() + run up to the first proposal
                    threads[c].start()
                    gate.parked[c].acquire()
                for k in range(K):
                    for c in range(N):
                        gate.advance(c)
                    if st["surrogate_is_updated"] and ((k + 1) % refit_every == 0 or k == K - 1) and shared.queue:
                        shared.updater.add_data(shared.queue)
                        shared.SOL, n = shared.updater.update()
                        shared.queue = []
                        shared.history.append((gstep + k + 1, n, shared.SOL[1]))
                        out["SOL_%d" % (len(shared.history) - 1)] = shared.SOL[0].copy()
                for t in threads:
                    t.join()
            code = {"prerejected": 0, "accepted": 1, "rejected": 2}
            for c in range(N):
                name = "alg%03d%s_rank%d" % (i, st["type"], c)
                raw = read_rows("saved_samples/s/raw_data/%s.csv" % name)
                data = read_rows("saved_samples/s/data/%s.csv" % name)
                out["st%d_c%d_flags" % (i, c)] = np.array([code[r[0]] for r in raw], dtype=np.uint8)
                out["st%d_c%d_data" % (i, c)] = np.array([[float(x) for x in r] for r in data])
                out["st%d_c%d_counters" % (i, c)] = np.array([algs[c].no_accepted, algs[c].no_rejected, algs[c].no_prerejected])
                out["st%d_c%d_final" % (i, c)] = np.array(algs[c].current_sample)
                current[c] = algs[c].current_sample
            gstep += K
    finally:
        os.chdir(cwd)
    out["history"] = np.array(shared.history)
    out["init"] = init
    out["meta"] = np.array([N, refit_every, total])
    np.savez_compressed(os.path.join(GOLDEN, "simple_lockstep.npz"), **out)


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    ns = R.load()
    gen_poly(ns)
    gen_rbf(ns)
    gen_misc(ns)
    gen_chains(ns)
    gen_simple_lockstep(ns)
    import scipy
    manifest = {"numpy": np.__version__, "scipy": scipy.__version__, "python": sys.version.split()[0],
                "generator": "python -m oracle.make_golden", "reference": "dom0015/surrDAMH @ /root/reference",
                "files": {f: os.path.getsize(os.path.join(GOLDEN, f)) for f in sorted(os.listdir(GOLDEN)) if f.endswith(".npz")}}
    with open(os.path.join(GOLDEN, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print(json.dumps(manifest, indent=1))


if __name__ == "__main__":
    main()
