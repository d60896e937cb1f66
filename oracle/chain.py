"""TEST INFRASTRUCTURE -- numpy restatement of the MH / delayed-acceptance MH chain
logic of the reference, one chain, steppable.

Follows /root/reference/surrDAMH/modules/classes_SAMPLER.py:
  Chain.prepare      <- Algorithm_PARENT.prepare            (:54-75) + DAMH prologue (:178-180)
  Chain._true_model  <- Algorithm_PARENT.request_observations (:77-81)
  Chain._accept      <- if_accepted                         (:83-93)
  Chain._reject      <- if_rejected                         (:95-101)
  Chain._test        <- _acceptance_log_symmetric           (:122-128)
  Chain.step_mh      <- Algorithm_MH.run loop body          (:149-155)
  Chain.step_damh    <- Algorithm_DAMH.run loop body        (:181-214)
  Chain.finish       <- close_files                         (:103-120)
Instead of CSV files the chain appends the same rows to in-memory lists
(`data_rows`, `raw_rows`, `damh_rows`); oracle/trace_csv.py writes them in the
reference's on-disk format.  The per-step record `trace` is what the CUDA path
emits (flag, proposal, posterior of the proposal, surrogate posterior of the
proposal) and is the unit of the parity tests.
"""
import numpy as np

PREREJECTED, ACCEPTED, REJECTED = 0, 1, 2
FLAG_NAMES = {PREREJECTED: "prerejected", ACCEPTED: "accepted", REJECTED: "rejected"}


class Snapshot:                                           # :318-325
    def __init__(self, sample=None, G_sample=None, weight=None):
        self.sample, self.G_sample, self.weight = sample, G_sample, weight


class Chain:
    """One sampler's state machine for one stage.

    solver    : proxy with send_parameters / recv_observations (true model G)
    surrogate : proxy with send_parameters / recv_observations (+ send_to_data_collector)
    uniform   : object with .uniform(0,1); default RandomState(seed) as in :39
    """

    def __init__(self, problem, proposal, solver, kind="MH", seed=0, initial_sample=None,
                 G_initial_sample=None, surrogate=None, surrogate_is_updated=True, uniform=None):
        self.problem, self.proposal, self.solver = problem, proposal, solver
        self.kind = kind
        self.surrogate = surrogate
        self.collects = surrogate is not None and surrogate_is_updated is True   # :30-35
        self.surrogate_is_updated = surrogate_is_updated
        self.current = problem.prior_mean.copy() if initial_sample is None else np.array(initial_sample)  # :23-26
        self.G_current = G_initial_sample
        self.uniform = np.random.RandomState(seed) if uniform is None else uniform
        self.seed = seed
        self.no_accepted = self.no_rejected = self.no_prerejected = 0
        self.data_rows, self.raw_rows, self.damh_rows, self.trace = [], [], [], []
        self.no_G_calls = 0

    # ---- pieces -------------------------------------------------------------
    def _test(self, log_ratio):
        return np.log(self.uniform.uniform(0.0, 1.0)) < log_ratio      # strict <, :123-126

    def _true_model(self, sample):
        self.solver.send_parameters(sample)
        self.no_G_calls += 1
        return self.solver.recv_observations()

    def _emit_data_row(self):                              # :130-134
        self.data_rows.append([1 + self.rej_current] + list(self.current)
                              + [self.post_current, self.pre_post_current])

    def _collect(self, sample, G, weight):                 # :136-138
        if self.collects:
            self.surrogate.send_to_data_collector(Snapshot(sample.copy(), np.array(G).copy(), weight))

    def _accept(self):
        self._emit_data_row()
        self._collect(self.current, self.G_current, self.rej_current + 1)   # displaced sample
        self.no_accepted += 1
        self.rej_current = 0
        self.current, self.G_current, self.post_current = self.proposed, self.G_proposed, self.post_proposed
        self.raw_rows.append(["accepted"] + list(self.proposed))

    def _reject(self):
        self.no_rejected += 1
        self.rej_current += 1
        self._collect(self.proposed, self.G_proposed, 0)
        self.raw_rows.append(["rejected"] + list(self.proposed))

    # ---- stage protocol -----------------------------------------------------
    def prepare(self):
        if self.G_current is None:                         # :70-72, one extra G call per stage
            self.G_current = self._true_model(self.current)
        self.post_current = self.problem.log_posterior(self.current, self.G_current)
        self.rej_current = 0
        self.pre_post_current = 0                          # :75
        if self.kind == "DAMH":                            # :178-180
            self.surrogate.send_parameters(self.current)
            GS = self.surrogate.recv_observations()
            self.pre_post_current = self.problem.log_posterior(self.current, GS)

    def step_mh(self, i):
        self.proposed = self.proposal.propose(self.current)
        self.G_proposed = self._true_model(self.proposed)
        self.post_proposed = self.problem.log_posterior(self.proposed, self.G_proposed)
        if self._test(self.post_proposed - self.post_current):
            self._accept()
            flag = ACCEPTED
        else:
            self._reject()
            flag = REJECTED
        self.trace.append((flag, self.proposed.copy(), self.post_proposed, np.nan))

    def step_damh(self, i):
        self.proposed = self.proposal.propose(self.current)
        # surrogate at BOTH points every step: it may have been swapped (:183-187)
        self.surrogate.send_parameters(np.array([self.current, self.proposed]))
        GS = self.surrogate.recv_observations()
        self.pre_post_current = self.problem.log_posterior(self.current, GS[0, :])
        pre_post_proposed = self.problem.log_posterior(self.proposed, GS[1, :])
        pre_log_ratio = pre_post_proposed - self.pre_post_current
        if self._test(pre_log_ratio):                      # stage 1
            self.G_proposed = self._true_model(self.proposed)
            self.post_proposed = self.problem.log_posterior(self.proposed, self.G_proposed)
            log_ratio = self.post_proposed - self.post_current
            if self._test(log_ratio - pre_log_ratio):      # stage 2, second uniform
                self.damh_rows.append(("accepted", i, self.post_proposed, pre_post_proposed))
                self._accept()
                flag = ACCEPTED
            else:
                self.damh_rows.append(("rejected", i, self.post_proposed, pre_post_proposed))
                self._reject()
                flag = REJECTED
            self.trace.append((flag, self.proposed.copy(), self.post_proposed, pre_post_proposed))
        else:
            self.no_prerejected += 1
            self.rej_current += 1
            self.raw_rows.append(["prerejected"] + list(self.proposed))
            self.trace.append((PREREJECTED, self.proposed.copy(), np.nan, pre_post_proposed))

    def step(self, i):
        (self.step_damh if self.kind == "DAMH" else self.step_mh)(i)

    def finish(self):
        self._emit_data_row()                              # :104

    def run(self, max_samples):
        self.prepare()
        for i in range(max_samples):
            self.step(i)
        self.finish()
        return self

    # ---- views used by the tests ---------------------------------------------
    def trace_arrays(self):
        flags = np.array([t[0] for t in self.trace], dtype=np.uint8)
        props = np.array([t[1] for t in self.trace], dtype=float).reshape(len(self.trace), -1)
        post = np.array([t[2] for t in self.trace], dtype=float)
        pre = np.array([t[3] for t in self.trace], dtype=float)
        return flags, props, post, pre
