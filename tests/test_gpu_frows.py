"""-m gpu: the rows SURVEY.md section 8(f) marks "next" -- run.py-shaped CLI (f2), the host-G farm overlapped with the
GPU's work (f3), device-side chain statistics on device tensors (f4)."""
import csv
import json
import os
import subprocess
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from tests.test_oracle_golden import SIMPLE_CONF

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _conf(**extra):
    from surrdamh_b200.configuration import Configuration
    c = dict(SIMPLE_CONF, solver_module_path=os.path.join(ROOT, "surrdamh_b200", "solvers.py"), solver_module_name="solvers",
             solver_init_name="Solver_linela2exp_local", solver_parameters={}, no_solvers=1, seed=12, refit_every=20)
    c["samplers_list"] = [
        {"type": "MH", "max_samples": 40, "time_limit": 600, "proposal_std": 1.0, "surrogate_is_updated": True},
        {"type": "DAMH", "max_samples": 60, "time_limit": 600, "proposal_std": 1.0, "surrogate_is_updated": True},
        {"type": "DAMH", "max_samples": 50, "time_limit": 600, "proposal_std": 1.0, "surrogate_is_updated": False}]
    c.update(extra)
    return Configuration(4, conf=c)


# ------------------------------------------------------------------------------------------------------------ f3
def test_host_G_farm_is_pipelined_and_changes_nothing():
    """The true model behind the reference's solver interface (set_parameters / get_observations) on a pool of 3 solver
    processes: the chains are worked in two halves, one half's G batch is queued while the other's is being evaluated
    (process_SOLVER.py:60-106 queues requests for its solvers), and every sample, counter and refit is bit-identical to
    the in-process host path and to the fused device path."""
    from surrdamh_b200.sampler import LockstepSampler
    out = {}
    for tag, kw in {"device": dict(device_solver=True), "host": dict(device_solver=False),
                    "pool": dict(device_solver=False, no_solvers=3)}.items():
        smp = LockstepSampler(_conf(**kw), no_chains=48)
        res = smp.run()
        out[tag] = (smp.block.current_numpy(), smp.block.counters_numpy(), smp.collector.model.coefs.cpu().numpy(),
                    [r.counters.copy() for r in res], smp.no_G_calls_host, dict(smp.overlap_stats))
        if smp.pool is not None:
            smp.pool.close()
    for k in range(3):                                                # pipelined pool == plain host loop, bit for bit
        assert np.array_equal(out["pool"][k], out["host"][k]), k
    # the fused device path evaluates the toy model with CUDA's exp instead of numpy's (last-bit differences in the
    # snapshots' G are possible), so it is compared as a distribution: the same acceptance statistics
    acc_d, acc_h = out["device"][1][:, 0].sum(), out["host"][1][:, 0].sum()
    assert abs(acc_d - acc_h) <= 0.05 * acc_h + 5
    assert out["host"][4] == out["pool"][4] > 0                       # the same number of true-model evaluations
    assert out["pool"][5]["queued_while_busy"] > 50                   # batches were queued behind a batch in flight
    assert out["host"][5]["queued_while_busy"] == 0


# ------------------------------------------------------------------------------------------------------------ f2
def test_run_py_cli_writes_reference_csv_layout(tmp_path):
    """`python run.py <conf> <N> csv` (run.py:12-69 of the reference: configuration name + number of samplers): N lockstep
    chains, the reference's saved_samples/ layout -- data/ rows [weight, u..., posterior, pre_posterior] whose weights sum
    to steps + 1 per chain and stage, raw_data/ one labelled row per step, notes/ one header + one row."""
    conf = json.load(open(os.path.join(ROOT, "examples", "simple_b200.json")))
    conf.pop("no_chains")
    conf["solver_module_path"] = os.path.join(ROOT, "surrdamh_b200", "solvers.py")
    conf["saved_samples_name"] = "cli_test"
    steps = [30, 50, 40]
    for st, n in zip(conf["samplers_list"], steps):
        st["max_samples"] = n
    conf["refit_every"] = 10
    path = tmp_path / "cli_test.json"
    path.write_text(json.dumps(conf))
    N = 6
    r = subprocess.run([sys.executable, os.path.join(ROOT, "run.py"), str(path), str(N), "csv", "--single"], cwd=tmp_path,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    assert r.stdout.count("--- SAMPLER alg") == 3
    base = tmp_path / "saved_samples" / "cli_test"
    kinds = ["MH", "DAMH", "DAMH"]
    for i, (kind, n) in enumerate(zip(kinds, steps)):
        for c in range(N):
            name = "alg%03d%s_rank%d.csv" % (i, kind, c)
            rows = list(csv.reader(open(base / "data" / name)))
            assert all(len(row) == 1 + 2 + 2 for row in rows)
            assert sum(int(float(row[0])) for row in rows) == n + 1       # compressed chain: weights = multiplicities
            raw = list(csv.reader(open(base / "raw_data" / name)))
            assert len(raw) == n and {row[0] for row in raw} <= {"accepted", "rejected", "prerejected"}
            notes = list(csv.reader(open(base / "notes" / name)))
            assert len(notes) == 2 and len(notes[0]) == len(notes[1])
            acc = sum(1 for row in raw if row[0] == "accepted")
            assert len(rows) == acc + 1
            if kind == "DAMH":
                a = list(csv.reader(open(base / "DAMH_accepted" / name)))
                rj = list(csv.reader(open(base / "DAMH_rejected" / name)))
                assert len(a) == acc and len(rj) == sum(1 for row in raw if row[0] == "rejected")


# ------------------------------------------------------------------------------------------------------------ f4
def _acf_numpy(x):                      # emcee.autocorr.function_1d, used at visualization_and_analysis.py:437
    n = 1
    while n < len(x):
        n <<= 1
    f = np.fft.fft(x - np.mean(x), n=2 * n)
    acf = np.fft.ifft(f * np.conjugate(f))[:len(x)].real
    return acf / acf[0]


def _auto_window(taus, c):              # visualization_and_analysis.py:720-724
    m = np.arange(len(taus)) < c * taus
    if np.any(m):
        return np.argmin(m)
    return len(taus) - 1


def test_device_statistics_of_a_gpu_chain_match_reference_formulas():
    """A DAMH run on the GPU with a full trace: the chain of current samples rebuilt ON THE DEVICE from the trace, its
    autocorrelation function / integrated autocorrelation time (Sokal window, visualization_and_analysis.py:720-736),
    weighted mean / covariance from the kernel's running moments (the decompressed-chain moments of :738-747) and the
    cost per uncorrelated sample (:580-594) -- all computed on CUDA tensors -- against numpy restatements of the
    reference's formulas applied to the same chains on the host."""
    from surrdamh_b200 import stats
    from surrdamh_b200.chains import ChainBlock, Proposal, Trace
    from surrdamh_b200.device import DeviceModel, DeviceProblem
    from tests.golden_util import SIMPLE_PROBLEM, load
    g = load("chains")
    sur = DeviceModel.poly(g["poly_MAT"], int(g["poly_deg"]), 2)
    prob = DeviceProblem(2, **SIMPLE_PROBLEM)
    C_, K = 96, 4000
    rs = np.random.RandomState(1)
    init = rs.standard_normal((C_, 2)) * 0.3 + np.array([7.0, 3.0])
    blk = ChainBlock(prob, C_, initial=init, seed=5, moments=True)
    blk.prepare("DAMH", sur, DeviceModel.toy())
    tr = Trace(K, C_, 2, 1, blk.device, snapshots=False, full=True)
    blk.run("DAMH", K, Proposal(2, 0.7), sur, DeviceModel.toy(), trace=tr)
    cur = stats.current_trace(tr.flags, tr.prop, L_dev(init.T))
    assert cur.is_cuda and cur.shape == (K, 2, C_)
    # host restatement of the chain: the reference's decompression (repeat each accepted sample by its weight)
    flags, props = tr.flags.cpu().numpy(), tr.prop.cpu().numpy()
    want = np.empty((K, 2, C_))
    state = init.T.copy()
    for k in range(K):
        acc = flags[k] == 1
        state[:, acc] = props[k][:, acc]
        want[k] = state
    assert np.array_equal(cur.cpu().numpy(), want)
    assert np.array_equal(want[-1], blk.current_numpy().T)
    # moments: running sums of the kernel vs the decompressed chain
    mean, cov = stats.chain_moments(blk)
    flat = want.transpose(1, 0, 2).reshape(2, -1)
    assert np.allclose(mean.cpu().numpy(), flat.mean(axis=1), rtol=1e-12)
    assert np.allclose(cov.cpu().numpy(), np.cov(flat, bias=True), rtol=1e-9, atol=1e-12)
    # autocorrelation time per chain and component
    for comp in range(2):
        acf = stats.autocorr_function(cur[:, comp, :])
        tau = stats.autocorr_time(acf, c=5.0)
        assert acf.is_cuda and tau.is_cuda
        for c in (0, 17, 95):
            ref = _acf_numpy(want[:, comp, c])
            assert np.allclose(acf[:, c].cpu().numpy(), ref, rtol=1e-8, atol=1e-10)
            taus = 2.0 * np.cumsum(ref) - 1.0
            assert abs(tau[c].item() - taus[_auto_window(taus, 5.0)]) < 1e-6 * max(1.0, abs(tau[c].item()))
    counters = blk.counters_numpy()
    got = stats.cpus(counters, tau, 0.05).item()
    full = (counters[:, 0] + counters[:, 1]) / counters[:, :3].sum(axis=1)
    assert abs(got - np.mean((full + 0.05) * tau.cpu().numpy())) < 1e-9 * abs(got)


def L_dev(x):
    from surrdamh_b200 import _lib as L
    return L.dev(np.ascontiguousarray(x))


def test_block_io_delivers_every_block_through_pinned_buffers():
    """surrdamh_b200/host_io.py: the double-buffered host I/O of the bench's end-to-end leg.  Six blocks of a DAMH run with the
    chains' samples uploaded from pinned memory at the start of every block and the product downloaded on the copy stream: the
    delivered samples, counters, moments and flags equal what a plain run that synchronises after every block reads back."""
    torch = pytest.importorskip("torch")
    from surrdamh_b200.chains import ChainBlock, Proposal, Trace
    from surrdamh_b200.device import DeviceModel, DeviceProblem
    from surrdamh_b200.host_io import BlockIO
    from tests.golden_util import SIMPLE_PROBLEM
    C_, K, nblk = 700, 9, 6
    rs = np.random.RandomState(5)
    centers = rs.standard_normal((200, 2)) * 2 + np.array([5.0, 3.0])
    coefs = np.zeros((203, 1))
    coefs[200:, 0] = [1e-4, -5e-5, -1.2e-3]
    sur, truth = DeviceModel.rbf(coefs, centers, 1), DeviceModel.toy()
    prop = Proposal(2, [0.4, 0.2])

    def run(piped):
        # every block restarts from the same uploaded samples (as bench.py's e2e leg does) with fresh random numbers
        blk = ChainBlock(DeviceProblem(2, **SIMPLE_PROBLEM), C_, seed=3, moments=True)
        u0 = blk.u.cpu().clone() + torch.from_numpy(rs0.standard_normal((2, C_))) * 0.1
        tr = [Trace(K, C_, 2, 1, blk.device, snapshots=False, full=False) for _ in range(2)]
        got = []
        if piped:
            io = BlockIO(blk, K)
            io.u_in.copy_(u0)
            io.upload(0)
            for b in range(nblk):
                io.begin(b)
                blk.prepare("DAMH", sur, truth)
                blk.run("DAMH", K, prop, sur, truth, trace=tr[b & 1])
                io.finish(b, tr[b & 1])
                io.upload(b + 1)
                if b >= 1:
                    got.append({k: v.clone() for k, v in io.result(b - 1).items()})
            got.append({k: v.clone() for k, v in io.result(nblk - 1).items()})
            return got
        for b in range(nblk):
            blk.u.copy_(u0.cuda())
            blk.prepare("DAMH", sur, truth)
            blk.run("DAMH", K, prop, sur, truth, trace=tr[b & 1])
            torch.cuda.synchronize()
            got.append({"u": blk.u.cpu().clone(), "counters": blk.counters.cpu().clone(), "moments": blk.moments.cpu().clone(),
                        "flags": tr[b & 1].flags[:K].cpu().clone()})
        return got

    rs0 = np.random.RandomState(9)
    a = run(True)
    rs0 = np.random.RandomState(9)
    b_ = run(False)
    assert len(a) == len(b_) == nblk
    for i, (x, y) in enumerate(zip(a, b_)):
        for k in ("u", "counters", "moments", "flags"):
            assert torch.equal(x[k], y[k]), (i, k)
    assert int(a[-1]["counters"][0].sum()) > 0
