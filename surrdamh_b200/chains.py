"""C lockstep MH / DAMH chains of one GPU: state in HBM as struct-of-arrays, stepped K steps per
kernel launch (sdb_chain_run).

Host-side counterpart of Algorithm_MH / Algorithm_DAMH (classes_SAMPLER.py:147-220) for many
chains at once.  The per-chain semantics (prepare, accept / reject bookkeeping, snapshot rules)
are implemented in csrc/sdb_chain.cu; this class only owns buffers and fills the argument struct.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from .device import DeviceModel, DeviceProblem

ALGORITHMS = {"MH": L.ALG_MH, "DAMH": L.ALG_DAMH}


class Trace:
    """Per-step output of one launch, [K][...][C] (step-major, chains contiguous)."""

    def __init__(self, K, C_, d, m, device, snapshots=False, full=True):
        self.K, self.C, self.d, self.m = K, C_, d, m
        self.flags = torch.empty((K, C_), dtype=torch.uint8, device=device)
        self.prop = torch.empty((K, d, C_), dtype=torch.float64, device=device) if full else None
        self.post = torch.empty((K, C_), dtype=torch.float64, device=device) if full else None
        self.pre = torch.empty((K, C_), dtype=torch.float64, device=device) if full else None
        self.pre_cur = torch.empty((K, C_), dtype=torch.float64, device=device) if full else None
        self.snap_par = torch.empty((K, d, C_), dtype=torch.float64, device=device) if snapshots else None
        self.snap_obs = torch.empty((K, m, C_), dtype=torch.float64, device=device) if snapshots else None

    def bytes_per_step_per_chain(self):
        b = 1
        if self.prop is not None:
            b += 8 * (self.d + 2)
        if self.snap_par is not None:
            b += 8 * (self.d + self.m)
        return b


class Proposal:
    """Gaussian random-walk proposal (Proposal_GaussRandomWalk, classes_SAMPLER.py:222-255).

    proposal_std: scalar or [d] -> per-component std shared by all chains; [C][d] with
    per_chain=True -> one std vector per chain (`proposal_differs`); [d][d] -> covariance, shipped
    as its lower Cholesky factor (u' = u + L z; numpy's multivariate_normal uses an SVD factor, so
    this case is distribution-equivalent, not stream-identical)."""

    def __init__(self, no_parameters, proposal_std=1.0, per_chain=False, device=None):
        d = no_parameters
        a = np.asarray(proposal_std, dtype=float)
        if per_chain:
            a = a.reshape(len(proposal_std), -1)
            if a.shape[1] == 1:
                a = np.repeat(a, d, axis=1)
            self.kind = L.PROP_PER_CHAIN
            self.scale = L.dev(np.ascontiguousarray(a.T), device=device)      # [d][C]
        elif a.ndim == 2:
            self.kind = L.PROP_FACTOR
            self.scale = L.dev(np.linalg.cholesky(a), device=device)
        else:
            if a.ndim == 0:
                a = np.full((d,), float(a))
            self.kind = L.PROP_SHARED
            self.scale = L.dev(a, device=device)


class ChainBlock:
    def __init__(self, problem: DeviceProblem, n_chains, initial=None, chain0=0, seed=0, stream_id=0,
                 moments=False, device=None):
        L.require_cuda()
        self.problem = problem
        self.C = int(n_chains)
        self.d, self.m = problem.d, problem.m
        self.device = torch.device(device or "cuda")
        self.chain0, self.seed, self.stream_id = int(chain0), int(seed), int(stream_id)
        self.step0 = 0
        self.surrogate_f32 = False           # additive: evaluate an RBF surrogate in FP32 (configuration key surrogate_precision)
        d, m, C_ = self.d, self.m, self.C
        if initial is None:                                   # prior mean (classes_SAMPLER.py:23-24)
            init = np.tile(problem.prior_mean_np.reshape(d, 1), (1, C_))
        else:
            init = np.asarray(initial, dtype=float).reshape(-1, d)
            if init.shape[0] == 1:
                init = np.repeat(init, C_, axis=0)
            if init.shape[0] != C_:
                raise L.SdbError(L.ERR_ARG, "initial samples: %d rows for %d chains" % (init.shape[0], C_))
            init = init.T
        self.u = L.dev(np.ascontiguousarray(init), device=self.device)
        self.G_u = torch.zeros((m, C_), dtype=torch.float64, device=self.device)
        self.post_u = torch.zeros((C_,), dtype=torch.float64, device=self.device)
        self.pre_u = torch.zeros((C_,), dtype=torch.float64, device=self.device)
        self.counters = torch.zeros((4, C_), dtype=torch.int32, device=self.device)
        self.n_mom = d + d * (d + 1) // 2
        self.moments = torch.zeros((self.n_mom, C_), dtype=torch.float64, device=self.device) if moments else None
        self.moment_steps = 0

    # ---- helpers -----------------------------------------------------------------------------
    def set_current(self, samples):
        s = np.asarray(samples, dtype=float).reshape(-1, self.d)
        if s.shape[0] == 1:
            s = np.repeat(s, self.C, axis=0)
        self.u.copy_(L.dev(np.ascontiguousarray(s.T), device=self.device))

    def reset_counters(self):
        self.counters.zero_()

    def _args(self, algorithm, phase, surrogate, truth):
        a = L.ChainArgs()
        a.algorithm = ALGORITHMS[algorithm] if isinstance(algorithm, str) else algorithm
        a.phase = phase
        a.problem = self.problem.struct
        a.surrogate = (surrogate or DeviceModel.none()).struct
        a.truth = (truth or DeviceModel.none()).struct
        st = a.state
        st.n_chains = self.C
        st.u, st.G_u, st.post_u, st.pre_u = L.ptr(self.u), L.ptr(self.G_u), L.ptr(self.post_u), L.ptr(self.pre_u)
        st.counters, st.moments = L.ptr(self.counters), L.ptr(self.moments)
        a.seed, a.stream_id, a.chain0, a.step0 = self.seed, self.stream_id, self.chain0, self.step0
        a.rng_kind = L.RNG_PHILOX
        a.surrogate_f32 = int(bool(self.surrogate_f32) and surrogate is not None and surrogate.kind == L.MODEL_RBF)
        self._keep = (surrogate, truth)
        return a

    def _launch(self, a, stream=None):
        L.check(L.lib().sdb_chain_run(C.byref(a), L.stream_handle(stream)))

    def plan(self, algorithm, surrogate, truth, proposal, lanes=0):
        a = self._args(algorithm, L.PHASE_FUSED, surrogate, truth)
        a.n_steps, a.lanes_per_chain = 1, lanes
        a.prop_kind, a.prop_scale = proposal.kind, L.ptr(proposal.scale)
        v = [C.c_int() for _ in range(4)]
        L.check(L.lib().sdb_chain_plan(C.byref(a), *[C.byref(x) for x in v]))
        return dict(lanes_per_chain=v[0].value, threads=v[1].value, blocks=v[2].value, smem=v[3].value)

    # ---- stage prologue ----------------------------------------------------------------------
    def prepare(self, algorithm, surrogate=None, truth=None, G_host=None, stream=None):
        """Algorithm_PARENT.prepare (+ the DAMH prologue): G(u) from the device truth model, or
        supplied by the host (`G_host` [C][m]) when G is a host plugin."""
        a = self._args(algorithm, L.PHASE_PREPARE, surrogate, truth)
        if G_host is not None:
            g = np.asarray(G_host, dtype=float).reshape(self.C, self.m)
            self.G_u.copy_(L.dev(np.ascontiguousarray(g.T), device=self.device))
            a.have_G = 1
        self._launch(a, stream)
        self.moment_steps = 0 if self.moments is None else self.moment_steps

    # ---- K fused steps -------------------------------------------------------------------------
    def run(self, algorithm, n_steps, proposal, surrogate=None, truth=None, trace=None, streams=None,
            refresh_pre=False, lanes=0, stream=None, coresident=False):
        """K fused steps.  The surrogate is evaluated in FP32 when `self.surrogate_f32` is set (RBF surrogates only)."""
        a = self._args(algorithm, L.PHASE_FUSED, surrogate, truth)
        a.coresident = int(coresident)
        a.n_steps, a.lanes_per_chain, a.refresh_pre = int(n_steps), lanes, int(bool(refresh_pre))
        a.prop_kind, a.prop_scale = proposal.kind, L.ptr(proposal.scale)
        self._fill_rng(a, streams)
        self._fill_trace(a, trace, n_steps)
        self._launch(a, stream)
        self.step0 += int(n_steps)
        if self.moments is not None:
            self.moment_steps += int(n_steps)

    # ---- split path: host-evaluated G ----------------------------------------------------------
    def propose(self, algorithm, proposal, surrogate, trace, streams=None, refresh_pre=False, stream=None, chains=None, step0=None):
        """chains = (lo, n): only the chains [lo, lo + n) take part (the pipelined split path); `step` bookkeeping is the
        caller's then (decide() advances step0 only for full launches or when told to)."""
        a = self._args(algorithm, L.PHASE_PROPOSE, surrogate, None)
        if chains is not None:
            a.chain_lo, a.chain_n = int(chains[0]), int(chains[1])
        a.n_steps, a.refresh_pre = 1, int(bool(refresh_pre))
        if step0 is not None:
            a.step0 = int(step0)
        a.prop_kind, a.prop_scale = proposal.kind, L.ptr(proposal.scale)
        self._fill_rng(a, streams)
        self._fill_trace(a, trace, 1)
        self._launch(a, stream)

    def decide(self, algorithm, proposal, surrogate, trace, G_prop, streams=None, stream=None, chains=None, advance=True,
               step0=None):
        a = self._args(algorithm, L.PHASE_DECIDE, surrogate, None)
        if chains is not None:
            a.chain_lo, a.chain_n = int(chains[0]), int(chains[1])
        if step0 is not None:
            a.step0 = int(step0)
        a.n_steps = 1
        a.prop_kind, a.prop_scale = proposal.kind, L.ptr(proposal.scale)
        self._fill_rng(a, streams)
        self._fill_trace(a, trace, 1)
        self._G_prop = G_prop
        a.G_prop = L.ptr(G_prop)
        self._launch(a, stream)
        if advance:
            self.step0 += 1
            if self.moments is not None:
                self.moment_steps += 1

    def _fill_rng(self, a, streams):
        if streams is not None:
            z, uni = streams
            a.rng_kind = L.RNG_STREAMS
            a.z, a.uni = L.ptr(z), L.ptr(uni)
            self._streams = streams

    def _fill_trace(self, a, trace, K):
        if trace is None:
            return
        if trace.K < K or trace.C != self.C:
            raise L.SdbError(L.ERR_ARG, "trace buffer [%d][%d] too small for %d steps of %d chains"
                             % (trace.K, trace.C, K, self.C))
        a.flags, a.prop, a.post, a.pre = L.ptr(trace.flags), L.ptr(trace.prop), L.ptr(trace.post), L.ptr(trace.pre)
        a.snap_par, a.snap_obs = L.ptr(trace.snap_par), L.ptr(trace.snap_obs)
        a.pre_cur = L.ptr(trace.pre_cur)
        self._trace = trace

    # ---- results ---------------------------------------------------------------------------------
    def counters_numpy(self):
        """[C][4]: accepted, rejected, prerejected, rejected-since-last-accept."""
        return self.counters.cpu().numpy().T.copy()

    def current_numpy(self):
        return self.u.cpu().numpy().T.copy()

    def philox_streams(self, K, step0=None, stream=None):
        """The (z [K][d][C], uni [K][2][C]) the kernel's Philox path would draw for steps step0.."""
        z = torch.empty((K, self.d, self.C), dtype=torch.float64, device=self.device)
        uni = torch.empty((K, 2, self.C), dtype=torch.float64, device=self.device)
        L.check(L.lib().sdb_philox_streams(self.seed, self.stream_id, self.chain0,
                                           self.step0 if step0 is None else step0, K, self.d, self.C,
                                           L.ptr(z), L.ptr(uni), L.stream_handle(stream)))
        return z, uni


def compact_snapshots(trace, K, cap, stream=None):
    """Valid snapshot rows of a launch, in (step, chain) order -> (par [n x d], obs [n x m], n)
    (the batch a sampler hands to the collector, classes_communication.py:121-135)."""
    C_, d, m = trace.C, trace.d, trace.m
    dev_ = trace.flags.device
    out_par = torch.empty((cap, d), dtype=torch.float64, device=dev_)
    out_obs = torch.empty((cap, m), dtype=torch.float64, device=dev_)
    count = torch.zeros((1,), dtype=torch.int64, device=dev_)
    scratch = torch.empty(((K * C_ + 1023) // 1024 + 1,), dtype=torch.int64, device=dev_)
    L.check(L.lib().sdb_snapshot_compact(L.ptr(trace.flags), L.ptr(trace.snap_par), L.ptr(trace.snap_obs), K, C_, d, m,
                                         L.ptr(out_par), L.ptr(out_obs), cap, L.ptr(count), L.ptr(scratch),
                                         L.stream_handle(stream)))
    return out_par, out_obs, count
