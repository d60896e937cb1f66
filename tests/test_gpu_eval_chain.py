"""GPU parity tests (through the C ABI): surrogate evaluation, Philox, the fused and split
MH / DAMH chain kernels -- against the golden vectors of the live reference and the numpy oracle.

Tolerances: surrogate values rtol 1e-10 (north_star; observed ~1e-14); accept / reject flags,
counters, proposals, final samples and snapshot rows bit-exact under the same (z, u) streams;
log-posterior values 1e-10 relative.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import chain as OC
from oracle import problem as OPR
from oracle import solvers as OS
from oracle import surrogate_poly as OP
from oracle import surrogate_rbf as ORB
from tests.golden_util import CORR_NOISE_PROBLEM, DARCY_PROBLEM, SIMPLE_PROBLEM, close, load, stream_pair

pytestmark = pytest.mark.gpu

RTOL = 1e-10


@pytest.fixture(scope="module")
def G():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from tests import gpu_util
    return gpu_util


# ------------------------------------------------------------------------------------------- Philox
def philox_ref(ctr, key):
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c, k = list(ctr), list(key)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
        k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
    return c


def test_philox_reference_kat():
    # Random123 known-answer vectors for philox4x32-10 (CPU check of the test's own reference)
    assert philox_ref([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert philox_ref([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert philox_ref([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_raw_matches_reference(G):
    from surrdamh_b200 import _lib as L
    out = torch.empty((64, 4), dtype=torch.int32, device="cuda")
    seed = 0x299f31d0a4093822
    L.check(L.lib().sdb_philox_raw(seed, 0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 64, L.ptr(out), L.stream_handle()))
    got = out.cpu().numpy().astype(np.uint32)
    for i in (0, 1, 17, 63):
        want = philox_ref([(0x243f6a88 + i) & 0xFFFFFFFF, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])
        assert list(got[i]) == want
    assert list(got[0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_philox_streams_layout_and_moments(G):
    from surrdamh_b200 import _lib as L
    K, d, C_ = 64, 3, 4096
    z = torch.empty((K, d, C_), dtype=torch.float64, device="cuda")
    u = torch.empty((K, 2, C_), dtype=torch.float64, device="cuda")
    L.check(L.lib().sdb_philox_streams(7, 1, 1000, 50, K, d, C_, L.ptr(z), L.ptr(u), L.stream_handle()))
    zn, un = z.cpu().numpy(), u.cpu().numpy()
    # uniforms: 53-bit value from the (x,y) / (z,w) words of block (chain, step, slot 0xFFFF, stream)
    for (k, c) in ((0, 0), (5, 77), (63, 4095)):
        b = philox_ref([1000 + c, 50 + k, 0xFFFF, 1], [7, 0])
        assert un[k, 0, c] == (((b[0] << 32) | b[1]) >> 11) * 2.0 ** -53
        assert un[k, 1, c] == (((b[2] << 32) | b[3]) >> 11) * 2.0 ** -53
        b = philox_ref([1000 + c, 50 + k, 0, 1], [7, 0])
        u1 = ((((b[0] << 32) | b[1]) >> 11) + 1) * 2.0 ** -53
        u2 = (((b[2] << 32) | b[3]) >> 11) * 2.0 ** -53
        r = np.sqrt(-2.0 * np.log(u1))
        assert abs(zn[k, 0, c] - r * np.cos(2 * np.pi * u2)) < 1e-12
        assert abs(zn[k, 1, c] - r * np.sin(2 * np.pi * u2)) < 1e-12
    assert (un >= 0).all() and (un < 1).all()
    assert abs(zn.mean()) < 5e-3 and abs(zn.std() - 1) < 5e-3 and abs(un.mean() - 0.5) < 2e-3
    # sharding independence: chains 2048.. of the same call == a call starting at chain0+2048
    z2 = torch.empty((K, d, 2048), dtype=torch.float64, device="cuda")
    u2_ = torch.empty((K, 2, 2048), dtype=torch.float64, device="cuda")
    L.check(L.lib().sdb_philox_streams(7, 1, 1000 + 2048, 50, K, d, 2048, L.ptr(z2), L.ptr(u2_), L.stream_handle()))
    assert torch.equal(z2, z[:, :, 2048:]) and torch.equal(u2_, u[:, :, 2048:])


# ------------------------------------------------------------------------------ surrogate evaluation
def test_poly_eval_golden(G):
    from surrdamh_b200 import surrogate_poly as SP
    from surrdamh_b200 import tables
    g = load("poly")
    for i, (d, m, deg) in enumerate(g["apply_cases"]):
        assert np.array_equal(tables.multi_indices(int(d), int(deg)), g[f"apply{i}_alpha"])
        assert np.array_equal(tables.hermite_table(int(deg)), g[f"apply{i}_herm"])
        got = SP.Surrogate_apply(int(d), int(m)).apply([g[f"apply{i}_MAT"], int(deg)], g[f"apply{i}_pts"])
        assert got.shape == g[f"apply{i}_GS"].shape
        assert close(got, g[f"apply{i}_GS"], rtol=RTOL), (d, m, deg)


def test_poly_eval_large_vs_oracle(G):
    from surrdamh_b200 import surrogate_poly as SP
    rs = np.random.RandomState(3)
    for d, m, deg, N in ((2, 1, 5, 300001), (4, 3, 5, 50000), (5, 2, 3, 1000), (2, 1, 10, 777)):
        P = OP.multi_indices(d, deg).shape[0]
        MAT = rs.standard_normal((P, m))
        pts = rs.standard_normal((N, d)) * 1.3 + 0.2
        got = SP.Surrogate_apply(d, m).apply([MAT, deg], pts)
        want = OP.PolyApply(d, m).apply([MAT, deg], pts)
        assert close(got, want, rtol=RTOL), (d, m, deg)
    # empty input
    assert SP.Surrogate_apply(2, 1).apply([rs.standard_normal((3, 1)), 1], np.empty((0, 2))).shape == (0, 1)


def test_rbf_eval_golden_all_kernels(G):
    from surrdamh_b200 import surrogate_rbf as SR
    g = load("rbf")
    for kt in range(8):
        got = SR.Surrogate_apply(3, 2, kernel_type=kt).apply([g["apply_coefs"], g["apply_centers"]], g["apply_pts"])
        assert close(got, g[f"apply_GS_kt{kt}"], rtol=RTOL), kt
    for tag, (d, m) in {"cfg3": (2, 1), "cfg5": (4, 3)}.items():
        got = SR.Surrogate_apply(d, m, kernel_type=1).apply([g[f"{tag}_coefs"], g[f"{tag}_centers"]], g[f"{tag}_pts"])
        assert close(got, g[f"{tag}_GS"], rtol=RTOL)


def test_rbf_eval_large_vs_oracle(G):
    """Multi-chunk ring (n > chunk), ragged sizes, several sweeps per CTA, a point on a center."""
    from surrdamh_b200 import surrogate_rbf as SR
    rs = np.random.RandomState(4)
    for d, m, n, N, kt in ((2, 1, 4096, 3000, 1), (2, 1, 5001, 333, 1), (4, 3, 2000, 1025, 1), (7, 2, 901, 40, 0),
                           (3, 5, 1, 9, 1), (2, 1, 300, 400000, 1), (4, 3, 500, 7, 4), (4, 3, 500, 7, 5)):
        centers = rs.standard_normal((n, d))
        coefs = rs.standard_normal((n + d + 1, m)) / np.sqrt(n)
        pts = rs.standard_normal((N, d))
        pts[0] = centers[0]
        got = SR.Surrogate_apply(d, m, kernel_type=kt).apply([coefs, centers], pts)
        ora = ORB.RbfApply(d, m, kernel_type=kt)
        want = np.vstack([ora.apply([coefs, centers], pts[i:i + 20000]) for i in range(0, N, 20000)])
        assert close(got, want, rtol=RTOL), (d, m, n, N, kt)


# ------------------------------------------------------------------------------------------ chains
def _check_against_golden(G, tr, blk, g, prefix, c, init, check_rows=True):
    flags, props, post, pre = G.chain_view(tr, c)
    assert np.array_equal(flags, g[prefix + "flags"]), prefix
    assert np.array_equal(props, g[prefix + "props"]), prefix
    assert close(post, g[prefix + "post"], rtol=RTOL) and close(pre, g[prefix + "pre"], rtol=RTOL), prefix
    assert list(blk.counters_numpy()[c][:3]) == list(g[prefix + "counters"])
    assert np.array_equal(blk.current_numpy()[c], g[prefix + "final"])
    if check_rows:
        rows = G.data_rows(tr, init, float(g[prefix + "post0"]), c)
        want = g[prefix + "data_rows"]
        assert np.array_equal(rows[:, :-1], want[:, :rows.shape[1] - 1])          # weights + samples exact
        assert close(rows[:, -1], want[:, rows.shape[1] - 1], rtol=RTOL)          # posterior of the current sample


@pytest.mark.parametrize("split", [False, True, "generic"])
def test_mh_and_damh_poly_golden(G, split, monkeypatch):
    """split False: fused launch (the lean d = 2 kernel); "generic": the same through the generic fused kernel
    (SDB_NO_LEAN); True: PROPOSE / DECIDE phases with the host evaluating G."""
    from surrdamh_b200.device import DeviceModel
    if split == "generic":
        monkeypatch.setenv("SDB_NO_LEAN", "1")
        split = False
    g = load("chains")
    C_, K = [int(x) for x in g["chains_meta"]]
    sur = DeviceModel.poly(g["poly_MAT"], int(g["poly_deg"]), 2)
    streams = [stream_pair(100 + 10 * c, K, 2) for c in range(C_)]

    def host_G(P):
        s, out = OS.LinElast2Exp(), []
        for p in P:
            s.set_parameters(p)
            out.append([s.get_observations()])
        return np.array(out).reshape(-1, 1)

    blk, tr = G.run_block("MH", 2, 1, SIMPLE_PROBLEM, DeviceModel.toy(), None, K, 1.0, streams, g["simple_init"],
                          split=host_G if split else None)
    for c in range(C_):
        _check_against_golden(G, tr, blk, g, f"mh_simple_c{c}_", c, g["simple_init"][c])
        par, obs, f = G.chain_snapshots(tr, c)
        assert np.array_equal(par, g[f"mh_simple_c{c}_snap_par"])
        assert close(obs, g[f"mh_simple_c{c}_snap_obs"], rtol=RTOL)
        assert np.array_equal(f == 2, g[f"mh_simple_c{c}_snap_wei"] == 0)       # weight 0 <=> a rejected proposal
    blk, tr = G.run_block("DAMH", 2, 1, SIMPLE_PROBLEM, DeviceModel.toy(), sur, K, [0.6, 0.9], streams, g["simple_init"],
                          split=host_G if split else None)
    for c in range(C_):
        _check_against_golden(G, tr, blk, g, f"damh_poly_c{c}_", c, g["simple_init"][c])
        par, obs, f = G.chain_snapshots(tr, c)
        assert np.array_equal(par, g[f"damh_poly_c{c}_snap_par"])
        assert np.array_equal(f == 2, g[f"damh_poly_c{c}_snap_wei"] == 0)


@pytest.mark.parametrize("lanes", [0, 1, 8, 32])
def test_damh_darcy_rbf_golden(G, lanes):
    from surrdamh_b200.device import DeviceModel
    g = load("chains")
    C_, K = [int(x) for x in g["darcy_meta"]]
    truth = DeviceModel.rbf(g["darcy_truth_coefs"], g["darcy_truth_centers"], 1)
    sur = DeviceModel.rbf(g["darcy_surr_coefs"], g["darcy_surr_centers"], 1)
    streams = [stream_pair(500 + 10 * c, K, 4) for c in range(C_)]
    blk, tr = G.run_block("DAMH", 4, 3, DARCY_PROBLEM, truth, sur, K, 0.2, streams, None, snapshots=False, lanes=lanes)
    for c in range(C_):
        _check_against_golden(G, tr, blk, g, f"damh_darcy_c{c}_", c, np.zeros(4))


def test_mh_correlated_noise_golden(G):
    from surrdamh_b200.device import DeviceModel
    g = load("chains")
    truth = DeviceModel.rbf(g["corr_truth_coefs"], g["corr_truth_centers"], 1)
    blk, tr = G.run_block("MH", 3, 3, CORR_NOISE_PROBLEM, truth, None, 800, [0.3, 0.5, 0.2], [stream_pair(900, 800, 3)], None)
    _check_against_golden(G, tr, blk, g, "mh_corr_", 0, np.array(CORR_NOISE_PROBLEM["prior_mean"]))


def test_philox_chains_match_oracle_on_dumped_streams(G):
    """Production RNG: run K steps with Philox inside the kernel, dump the same (z, u) with
    sdb_philox_streams, replay them through the numpy oracle chain: identical decisions."""
    from surrdamh_b200 import _lib as L
    from surrdamh_b200.chains import ChainBlock, Proposal, Trace
    from surrdamh_b200.device import DeviceModel
    g = load("chains")
    K, C_ = 400, 70
    sur = DeviceModel.poly(g["poly_MAT"], int(g["poly_deg"]), 2)
    prob = G.problem_from_kw(2, SIMPLE_PROBLEM)
    init = np.random.RandomState(5).standard_normal((C_, 2)) * 0.5 + np.array([7.0, 3.0])
    blk = ChainBlock(prob, C_, initial=init, chain0=12345, seed=99, stream_id=3)
    prop = Proposal(2, [0.6, 0.9])
    z, u = blk.philox_streams(K)
    tr = Trace(K, C_, 2, 1, blk.device, snapshots=True)
    blk.prepare("DAMH", sur, DeviceModel.toy())
    blk.run("DAMH", K, prop, sur, DeviceModel.toy(), trace=tr)
    torch.cuda.synchronize()
    zn, un = z.cpu().numpy(), u.cpu().numpy()
    SOL = [g["poly_MAT"], int(g["poly_deg"])]
    for c in (0, 1, 33, 69):
        gen = OPR.StreamGenerators(zn[:, :, c], un[:, :, c])
        pr = OPR.GaussProblem(2, **SIMPLE_PROBLEM)
        pp = OPR.RandomWalkProposal(2, proposal_std=[0.6, 0.9], generator=gen)

        class S:
            def send_parameters(self, p):
                self.p = p.copy()

            def recv_observations(self):
                return OP.PolyApply(2, 1).apply(SOL, self.p)

            def send_to_data_collector(self, s):
                pass
        ch = OC.Chain(pr, pp, OS.LocalSolverProxy(OS.LinElast2Exp()), kind="DAMH", initial_sample=init[c], surrogate=S(),
                      surrogate_is_updated=False, uniform=gen).run(K)
        flags, props, post, pre = ch.trace_arrays()
        gf, gp, gpost, gpre = G.chain_view(tr, c)
        assert np.array_equal(flags, gf) and np.array_equal(props, gp)
        assert close(gpost, post, rtol=RTOL) and close(gpre, pre, rtol=RTOL)
        assert list(blk.counters_numpy()[c][:3]) == [ch.no_accepted, ch.no_rejected, ch.no_prerejected]
    # results do not depend on how chains are sharded: rerun chains 32.. as their own block
    blk2 = ChainBlock(prob, C_ - 32, initial=init[32:], chain0=12345 + 32, seed=99, stream_id=3)
    tr2 = Trace(K, C_ - 32, 2, 1, blk.device, snapshots=True)
    blk2.prepare("DAMH", sur, DeviceModel.toy())
    blk2.run("DAMH", K, prop, sur, DeviceModel.toy(), trace=tr2)
    assert torch.equal(tr2.flags, tr.flags[:, 32:].contiguous()) and torch.equal(blk2.u, blk.u[:, 32:].contiguous())
    # and K steps in one launch == K/4 steps in four launches
    blk3 = ChainBlock(prob, C_, initial=init, chain0=12345, seed=99, stream_id=3)
    blk3.prepare("DAMH", sur, DeviceModel.toy())
    for _ in range(4):
        blk3.run("DAMH", K // 4, prop, sur, DeviceModel.toy())
    assert torch.equal(blk3.u, blk.u) and torch.equal(blk3.counters, blk.counters)


def test_snapshot_compaction_order(G):
    from surrdamh_b200.chains import compact_snapshots
    from surrdamh_b200.device import DeviceModel
    g = load("chains")
    C_, K = [int(x) for x in g["chains_meta"]]
    streams = [stream_pair(100 + 10 * c, K, 2) for c in range(C_)]
    blk, tr = G.run_block("MH", 2, 1, SIMPLE_PROBLEM, DeviceModel.toy(), None, K, 1.0, streams, g["simple_init"])
    par, obs, cnt = compact_snapshots(tr, K, K * C_)
    n = int(cnt.item())
    assert n == K * C_                               # MH: every step yields a snapshot
    f = tr.flags.cpu().numpy()
    sp = tr.snap_par.cpu().numpy()
    want = np.array([sp[k, :, c] for k in range(K) for c in range(C_) if f[k, c] != 0])
    assert np.array_equal(par[:n].cpu().numpy(), want)
    # capped
    par2, obs2, cnt2 = compact_snapshots(tr, K, 100)
    assert int(cnt2.item()) == n and np.array_equal(par2.cpu().numpy(), want[:100])


def test_uninformative_posterior_equals_prior(G):
    """Known-answer check (examples/uninformative.json): G == 0 => the chain samples N(0, 3^2 I)."""
    from surrdamh_b200.chains import ChainBlock, Proposal
    from surrdamh_b200.device import DeviceModel, DeviceProblem
    prob = DeviceProblem(2, prior_mean=0.0, prior_std=3.0, noise_std=1.0, no_observations=1, observations=[0.0])
    blk = ChainBlock(prob, 8192, moments=True, seed=1)
    prop = Proposal(2, 3.0)
    blk.prepare("MH", None, DeviceModel.zero())
    blk.run("MH", 200, prop, None, DeviceModel.zero())     # burn-in
    blk.moments.zero_()
    blk.run("MH", 400, prop, None, DeviceModel.zero())
    mom = blk.moments.sum(dim=1).cpu().numpy() / (8192 * 400)
    mean = mom[:2]
    cov = np.array([[mom[2], mom[3]], [mom[3], mom[4]]]) - np.outer(mean, mean)
    assert np.all(np.abs(mean) < 0.05)
    assert abs(cov[0, 0] - 9) < 0.25 and abs(cov[1, 1] - 9) < 0.25 and abs(cov[0, 1]) < 0.2


def test_errors_are_loud(G):
    from surrdamh_b200 import _lib as L
    from surrdamh_b200.chains import ChainBlock, Proposal
    from surrdamh_b200.device import DeviceModel
    prob = G.problem_from_kw(2, SIMPLE_PROBLEM)
    blk = ChainBlock(prob, 4)
    with pytest.raises(L.SdbError):        # DAMH without a surrogate (the reference crashes there too)
        blk.prepare("DAMH", None, DeviceModel.toy())
    with pytest.raises(L.SdbError):        # CPU tensors are refused, there is no CPU path
        L.ptr(torch.zeros(3))


def test_per_chain_and_correlated_proposals(G):
    """proposal_differs (one std vector per chain, process_SAMPLER.py:55-57) and a 2-D proposal_std (covariance):
    proposals follow u' = u + std_c * z, resp. u' = u + L z with L the Cholesky factor; decisions follow the oracle."""
    from surrdamh_b200 import _lib as L
    from surrdamh_b200.chains import ChainBlock, Proposal, Trace
    from surrdamh_b200.device import DeviceModel
    K, C_ = 50, 6
    rs = np.random.RandomState(21)
    stds = [0.2 + 0.3 * c for c in range(C_)]                      # scalar per chain, as illustrative_scaling*.json
    streams = [stream_pair(300 + c, K, 2) for c in range(C_)]
    prob = G.problem_from_kw(2, SIMPLE_PROBLEM)
    z, uni = G.pack_streams(streams)
    init = rs.standard_normal((C_, 2)) * 0.3 + np.array([7.0, 3.0])
    blk = ChainBlock(prob, C_, initial=init)
    prop = Proposal(2, stds, per_chain=True)
    tr = Trace(K, C_, 2, 1, blk.device)
    blk.prepare("MH", None, DeviceModel.toy())
    blk.run("MH", K, prop, None, DeviceModel.toy(), trace=tr, streams=(z, uni))
    torch.cuda.synchronize()
    for c in range(C_):
        gen = OPR.StreamGenerators(*streams[c])
        ch = OC.Chain(OPR.GaussProblem(2, **SIMPLE_PROBLEM), OPR.RandomWalkProposal(2, proposal_std=stds[c], generator=gen),
                      OS.LocalSolverProxy(OS.LinElast2Exp()), kind="MH", initial_sample=init[c], uniform=gen).run(K)
        f, p, post, _ = ch.trace_arrays()
        gf, gp, gpost, _ = G.chain_view(tr, c)
        assert np.array_equal(f, gf) and np.array_equal(p, gp) and close(gpost, post, rtol=RTOL), c
    # covariance proposal: u' = u + L z
    cov = np.array([[0.5, 0.2], [0.2, 0.3]])
    Lc = np.linalg.cholesky(cov)
    blk = ChainBlock(prob, C_, initial=init)
    prop = Proposal(2, cov)
    assert prop.kind == L.PROP_FACTOR
    tr = Trace(1, C_, 2, 1, blk.device)
    blk.prepare("MH", None, DeviceModel.toy())
    blk.run("MH", 1, prop, None, DeviceModel.toy(), trace=tr, streams=(z[:1].contiguous(), uni[:1].contiguous()))
    torch.cuda.synchronize()
    for c in range(C_):
        want = init[c] + Lc @ streams[c][0][0]
        assert np.allclose(tr.prop[0, :, c].cpu().numpy(), want, rtol=1e-14, atol=1e-14)


def test_generic_dimensions_chain(G):
    """Shapes outside the specialised kernels (d = 5, m = 7, correlated prior): the generic kernel against the
    oracle chain, RBF surrogate and RBF truth.  (Correlated NOISE with DAMH is not runnable in the reference: its
    prologue hands a (1, m) surrogate output to np.linalg.solve, classes_SAMPLER.py:178-180,300-303; correlated
    noise is covered with MH by the golden mh_corr case.)"""
    from surrdamh_b200.chains import ChainBlock, Proposal, Trace
    from surrdamh_b200.device import DeviceModel, DeviceProblem
    rs = np.random.RandomState(22)
    d, m, K, C_ = 5, 7, 120, 3
    A = rs.standard_normal((d, d)); prior_cov = A @ A.T + d * np.eye(d)
    noise_std = list(0.2 + 0.1 * rs.uniform(size=m))
    kw = dict(prior_mean=list(rs.standard_normal(d)), prior_std=prior_cov.tolist(), noise_std=noise_std,
              no_observations=m, observations=list(rs.standard_normal(m)))
    cen_t, co_t = rs.standard_normal((90, d)), rs.standard_normal((90 + d + 1, m)) * 0.05
    cen_s, co_s = rs.standard_normal((40, d)), rs.standard_normal((40 + d + 1, m)) * 0.05
    truth, sur = DeviceModel.rbf(co_t, cen_t, 1), DeviceModel.rbf(co_s, cen_s, 0)
    streams = [stream_pair(700 + c, K, d) for c in range(C_)]
    blk, tr = G.run_block("DAMH", d, m, kw, truth, sur, K, 0.4, streams, None, snapshots=False)
    for c in range(C_):
        gen = OPR.StreamGenerators(*streams[c])

        class S:
            def send_parameters(self, p):
                self.p = p.copy()

            def recv_observations(self):
                return ORB.RbfApply(d, m, kernel_type=0).apply([co_s, cen_s], self.p)
        ch = OC.Chain(OPR.GaussProblem(d, **kw), OPR.RandomWalkProposal(d, proposal_std=0.4, generator=gen),
                      OS.LocalSolverProxy(OS.SurrogateAsSolver(ORB.RbfApply(d, m, kernel_type=1), [co_t, cen_t])), kind="DAMH",
                      surrogate=S(), surrogate_is_updated=False, uniform=gen).run(K)
        f, p, post, pre = ch.trace_arrays()
        gf, gp, gpost, gpre = G.chain_view(tr, c)
        assert np.array_equal(f, gf) and np.array_equal(p, gp), c
        assert close(gpost, post, rtol=1e-9) and close(gpre, pre, rtol=1e-9), c


def test_lean_kernel_agrees_with_generic_kernel_over_a_long_run(G):
    """The lean kernel of the toy shape sums PHI.MAT in nested-loop order, the generic kernel in the reference's term
    order: surrogate posteriors agree to rounding (1e-12), so decisions are identical except where a test falls within
    that rounding of a tie -- over 4096 chains x 1500 steps the two kernels must make the same 6.1 million decisions
    (a mismatch would let the chains diverge, which the final samples would show)."""
    import os
    from surrdamh_b200.chains import ChainBlock, Proposal, Trace
    from surrdamh_b200.device import DeviceModel, DeviceProblem
    g = load("chains")
    sur = DeviceModel.poly(g["poly_MAT"], int(g["poly_deg"]), 2)
    prob = DeviceProblem(2, **SIMPLE_PROBLEM)
    out = {}
    for lean in (True, False):
        if lean:
            os.environ.pop("SDB_NO_LEAN", None)
        else:
            os.environ["SDB_NO_LEAN"] = "1"
        try:
            blk = ChainBlock(prob, 4096, seed=17)
            blk.prepare("DAMH", sur, DeviceModel.toy())
            tr = Trace(1500, 4096, 2, 1, blk.device, snapshots=False, full=True)
            blk.run("DAMH", 1500, Proposal(2, 1.0), sur, DeviceModel.toy(), trace=tr)
            torch.cuda.synchronize()
            out[lean] = (blk.u.cpu().numpy(), blk.counters.cpu().numpy(), tr.flags.cpu().numpy(), tr.pre.cpu().numpy())
        finally:
            os.environ.pop("SDB_NO_LEAN", None)
    a, b = out[True], out[False]
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0])
    assert close(a[3], b[3], rtol=1e-12)
