"""CPU, build container only: the numpy oracle against the LIVE reference imported
from /root/reference (bit-for-bit on the same machine).  Skipped where the reference
tree is absent (the GPU box)."""
import contextlib
import csv
import io
import os
import tempfile

import numpy as np
import pytest

from oracle import chain as OC
from oracle import lhs as OL
from oracle import problem as OPR
from oracle import ref_import as R
from oracle import solvers as OS
from oracle import surrogate_poly as OP
from oracle import surrogate_rbf as ORB
from tests.golden_util import DARCY_PROBLEM, SIMPLE_PROBLEM

pytestmark = pytest.mark.live_reference


@pytest.fixture(scope="module")
def ns():
    with contextlib.redirect_stdout(io.StringIO()):
        return R.load()


class Snap:
    def __init__(self, a, b):
        self.sample, self.G_sample, self.weight = a, b, 1


def test_tables(ns):
    for d in (1, 2, 3, 4, 6):
        for deg in range(0, 7 if d < 6 else 4):
            assert np.array_equal(OP.multi_indices(d, deg), ns.poly.generate_polynomials_degree(d, deg))
    for deg in range(0, 10):
        assert np.array_equal(OP.hermite_table(deg), ns.poly.hermite_poly_normalized(deg))


def test_poly_apply_update_bitwise(ns):
    rs = np.random.RandomState(1)
    with contextlib.redirect_stdout(io.StringIO()):
        for d, m, deg in ((2, 1, 5), (4, 3, 5), (3, 2, 4), (2, 1, 1)):
            P = OP.multi_indices(d, deg).shape[0]
            MAT = rs.standard_normal((P, m))
            for npts in (1, 2, 9):
                pts = rs.standard_normal((npts, d)) * 2 + 1
                assert np.array_equal(ns.poly.Surrogate_apply(d, m).apply([MAT, deg], pts), OP.PolyApply(d, m).apply([MAT, deg], pts))
        for d, m, maxdeg, n1, n2 in ((2, 1, 5, 30, 500), (4, 3, 5, 50, 800), (3, 2, 3, 20, 100)):
            ru, ou = ns.poly.Surrogate_update(d, m, max_degree=maxdeg), OP.PolyUpdate(d, m, max_degree=maxdeg)
            X = rs.standard_normal((n1 + n2, d))
            Y = np.sin(X @ rs.standard_normal((d, m)))
            for lo, hi in ((0, n1), (n1, n1 + n2)):
                batch = [Snap(X[i], Y[i]) for i in range(lo, hi)]
                ru.add_data(batch)
                ou.add_data(batch)
                (rM, rdeg), rn = ru.update()
                (oM, odeg), on = ou.update()
                assert np.array_equal(rM, oM) and rdeg == odeg and rn == on


def test_rbf_apply_update_bitwise(ns):
    rs = np.random.RandomState(2)
    d, m, n = 3, 2, 40
    C, CO, pts = rs.standard_normal((n, d)), rs.standard_normal((n + d + 1, m)), rs.standard_normal((5, d))
    for kt in range(8):
        assert np.array_equal(ns.rbf.Surrogate_apply(d, m, kernel_type=kt).apply([CO, C], pts),
                              ORB.RbfApply(d, m, kernel_type=kt).apply([CO, C], pts))
    d, m = 4, 3
    for kt, solver, keep in ((1, "direct", None), (1, "minres", None), (0, "direct", None), (4, "direct", None),
                             (5, "minres", None), (1, "direct", 25), (2, "minres", 30), (1, "direct", 1000)):
        ru = ns.rbf.Surrogate_update(d, m, kernel_type=kt, solver_type=solver, no_keep=keep)
        ou = ORB.RbfUpdate(d, m, kernel_type=kt, solver_type=solver, no_keep=keep)
        X = rs.standard_normal((90, d))
        Y = np.sin(X @ rs.standard_normal((d, m)))
        for lo, hi in ((0, 40), (40, 90)):
            batch = [Snap(X[i], Y[i]) for i in range(lo, hi)]
            ru.add_data(batch)
            ou.add_data(batch)
            (rCO, rC), rn = ru.update()
            (oCO, oC), on = ou.update()
            assert np.array_equal(rCO, oCO) and np.array_equal(rC, oC) and rn == on


def test_lhs_and_toy_bitwise(ns):
    for n, d, mean, std in ((4, 2, [5.0, 3.0], [[4, -2], [-2, 4]]), (7, 4, 0.0, 1.0), (16, 3, [1., 2., 3.], [1., .5, 2.])):
        assert np.array_equal(ns.lhs.lhs_normal(d, mean, std, n, 0), OL.lhs_normal(d, mean, std, n, 0))
    rs = np.random.RandomState(3)
    a, b = ns.solvers.Solver_linela2exp_local(), OS.LinElast2Exp()
    for _ in range(500):
        p = rs.standard_normal(2) * 2 + np.array([5, 3])
        a.set_parameters(p)
        b.set_parameters(p)
        assert a.get_observations() == b.get_observations()


class _Proxy:
    def __init__(self, solver):
        self.solver = solver

    def send_parameters(self, p):
        self.solver.set_parameters(p)

    def recv_observations(self):
        return self.solver.get_observations()


class _Surr:
    def __init__(self, a, SOL):
        self.a, self.SOL, self.snaps = a, SOL, []

    def send_parameters(self, p):
        self.p = p.copy()

    def recv_observations(self):
        return self.a.apply(self.SOL, self.p)

    def send_to_data_collector(self, s):
        self.snaps.append(s)


def _rows(path):
    with open(path) as f:
        return [r for r in csv.reader(f)]


def _pair(ns, kind, d, kw, solvers, surrs, steps, std, seed, streams=None, updated=True):
    cwd, tmp = os.getcwd(), tempfile.mkdtemp()
    os.chdir(tmp)
    try:
        rprob = ns.cS.Problem_Gauss(d, **kw, saved_samples_name="t")
        rprop = ns.cS.Proposal_GaussRandomWalk(d, proposal_std=std, seed=seed + 1)
        rsurr = None if surrs is None else _Surr(*surrs[0])
        cls = ns.cS.Algorithm_MH if kind == "MH" else ns.cS.Algorithm_DAMH
        alg = cls(rprob, rprop, _Proxy(solvers[0]), Surrogate=rsurr, surrogate_is_updated=updated, max_samples=steps,
                  name="a", seed=seed + 2, save_raw_data=True)
        if streams is not None:
            R.inject_streams(alg, rprop, R.Streams(*streams))
        with contextlib.redirect_stdout(io.StringIO()):
            alg.run()
        gen = None if streams is None else OPR.StreamGenerators(*streams)
        oprop = OPR.RandomWalkProposal(d, proposal_std=std, seed=seed + 1, generator=gen)
        osurr = None if surrs is None else _Surr(*surrs[1])
        ch = OC.Chain(OPR.GaussProblem(d, **kw), oprop, OS.LocalSolverProxy(solvers[1]), kind=kind, seed=seed + 2,
                      surrogate=osurr, surrogate_is_updated=updated, uniform=gen).run(steps)
        assert (alg.no_accepted, alg.no_rejected, alg.no_prerejected) == (ch.no_accepted, ch.no_rejected, ch.no_prerejected)
        assert [[str(x) for x in r] for r in ch.data_rows] == _rows("saved_samples/t/data/a.csv")
        assert [[str(x) for x in r] for r in ch.raw_rows] == _rows("saved_samples/t/raw_data/a.csv")
        if kind == "DAMH":
            for tag in ("accepted", "rejected"):
                mine = [[str(r[1]), str(r[2]), str(r[3])] for r in ch.damh_rows if r[0] == tag]
                assert mine == _rows("saved_samples/t/DAMH_%s/a.csv" % tag)
        if rsurr is not None:
            assert len(rsurr.snaps) == len(osurr.snaps)
            for a, b in zip(rsurr.snaps, osurr.snaps):
                assert np.array_equal(a.sample, b.sample) and np.array_equal(a.G_sample, b.G_sample) and a.weight == b.weight
        assert np.array_equal(alg.current_sample, ch.current)
    finally:
        os.chdir(cwd)


def test_chain_logic_bitwise(ns):
    rs = np.random.RandomState(3)
    with contextlib.redirect_stdout(io.StringIO()):
        ou = OP.PolyUpdate(2, 1, max_degree=5)
        X = rs.standard_normal((800, 2)) * 2 + np.array([5, 3])
        s = OS.LinElast2Exp()
        Y = []
        for x in X:
            s.set_parameters(x)
            Y.append(s.get_observations())
        ou.add_arrays(X, np.array(Y))
        SOLp, _ = ou.update()
        ra = ns.poly.Surrogate_apply(2, 1)
    sol = lambda: (ns.solvers.Solver_linela2exp_local(), OS.LinElast2Exp())
    sur = lambda: ((ra, SOLp), (OP.PolyApply(2, 1), SOLp))
    for seed in (0, 7):
        _pair(ns, "MH", 2, SIMPLE_PROBLEM, sol(), None, 300, 1.0, seed)
        _pair(ns, "MH", 2, SIMPLE_PROBLEM, sol(), sur(), 300, [0.5, 1.0], seed)
        _pair(ns, "DAMH", 2, SIMPLE_PROBLEM, sol(), sur(), 1500, 1.0, seed)
        _pair(ns, "DAMH", 2, SIMPLE_PROBLEM, sol(), sur(), 600, 1.0, seed, updated=False)
        z, u = rs.standard_normal((1000, 2)), rs.uniform(size=(1000, 2))
        _pair(ns, "DAMH", 2, SIMPLE_PROBLEM, sol(), sur(), 1000, 1.0, seed, streams=(z, u))
        _pair(ns, "MH", 2, SIMPLE_PROBLEM, sol(), None, 1000, 1.0, seed, streams=(z, u))
    d, m = 4, 3
    Xc, W = rs.standard_normal((300, d)), rs.standard_normal((d, m))
    f = lambda X: np.array([10, -6, -4]) + 2 * np.sin(X @ W)
    tu = ORB.RbfUpdate(d, m, kernel_type=1, solver_type="direct")
    tu.add_arrays(Xc, f(Xc))
    SOLt, _ = tu.update()
    su = ORB.RbfUpdate(d, m, kernel_type=1, solver_type="direct")
    su.add_arrays(Xc[:80], f(Xc[:80]))
    SOLs, _ = su.update()
    truth = (OS.SurrogateAsSolver(ns.rbf.Surrogate_apply(d, m, kernel_type=1), SOLt), OS.SurrogateAsSolver(ORB.RbfApply(d, m, kernel_type=1), SOLt))
    surr = ((ns.rbf.Surrogate_apply(d, m, kernel_type=1), SOLs), (ORB.RbfApply(d, m, kernel_type=1), SOLs))
    _pair(ns, "DAMH", d, DARCY_PROBLEM, truth, surr, 1000, 0.2, 5)
