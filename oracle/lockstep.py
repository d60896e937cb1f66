"""TEST INFRASTRUCTURE -- the sampler stage loop and the collector's refit policy,
restated for C chains stepping in LOCKSTEP under a deterministic schedule.

Follows
  /root/reference/surrDAMH/process_SAMPLER.py:25,46-88   (seeds, stage loop, hand-over
      of current_sample to the next stage, per-stage proposal_std / initial_sample)
  /root/reference/surrDAMH/process_COLLECTOR.py:100-113  (whenever snapshots arrived:
      add_data + a full update(), then push SOL to the samplers)
  /root/reference/surrDAMH/modules/classes_communication.py:83-150
      (Solver_local_collector_MPI: local surrogate evaluation with the sampler's
       current SOL; snapshots queued for the collector; SOL swapped when one arrives)

What is different, on purpose (SURVEY.md section 7 trap 7): the reference's refit timing
is a race between busy-polling MPI ranks.  Here -- and identically in the CUDA
driver -- it is a schedule: all chains take step k, their snapshots are queued in
(step, chain) order, and after every `refit_every` steps (and after the last step
of a stage) the queue, if non-empty, is handed to the updater and the new SOL is
seen by every chain from the next step on.  A DAMH stage with no SOL yet is an
error (the reference crashes there too, trap 10).
"""
import numpy as np

from .chain import Chain
from .problem import GaussProblem, RandomWalkProposal
from .solvers import LocalSolverProxy


class SharedSurrogate:
    """The state all lockstep chains share: current SOL + the snapshot queue."""

    def __init__(self, apply_obj, updater):
        self.apply_obj, self.updater = apply_obj, updater
        self.SOL = None
        self.queue = []
        self.no_refits = 0
        self.history = []          # (global_step, no_snapshots) per refit

    def refit_if_pending(self, global_step):
        if not self.queue:
            return False
        self.updater.add_data(self.queue)
        self.SOL, n = self.updater.update()
        self.queue = []
        self.no_refits += 1
        self.history.append((global_step, n))
        return True


class SurrogateProxy:
    """Per-chain view with the proxy interface the chain calls."""

    def __init__(self, shared):
        self.shared = shared

    def send_parameters(self, parameters):
        self.parameters = parameters.copy()

    def recv_observations(self):
        if self.shared.SOL is None:
            raise RuntimeError("DAMH step before any surrogate was fitted")
        return self.shared.apply_obj.apply(self.shared.SOL, self.parameters)

    def send_to_data_collector(self, snapshot):
        self.shared.queue.append(snapshot)


def sampler_seed0(rank, no_chains):
    # process_SAMPLER.py:25 with size_world = samplers + pool + collector
    return max(1000, no_chains + 2) * rank


def run_lockstep(conf, no_chains, solver_factory, apply_obj=None, updater=None, refit_every=64,
                 initial_samples=None, generators=None, keep_chains=True, initial_SOL=None):
    """Run every stage of conf['samplers_list'] for `no_chains` chains.

    conf           : dict with the reference JSON keys (no_parameters, no_observations,
                     problem_parameters, samplers_list, ...)
    solver_factory : () -> plugin object with set_parameters / get_observations
    generators     : optional list (per chain) of objects providing .normal and .uniform
                     driven by explicit streams indexed by GLOBAL step; default numpy
                     RandomState seeded as in process_SAMPLER.py:25,46,64
    Returns (stages, shared): stages[i] = list of finished Chain objects of stage i.
    """
    d, m = conf["no_parameters"], conf["no_observations"]
    pp = conf["problem_parameters"]
    problem = GaussProblem(d, prior_mean=pp["prior_mean"], prior_std=pp["prior_std"],
                           noise_std=pp["noise_std"], no_observations=m,
                           observations=np.array(pp["observations"], dtype=float).reshape(-1))
    shared = SharedSurrogate(apply_obj, updater) if apply_obj is not None else None
    if shared is not None and initial_SOL is not None:
        shared.SOL = initial_SOL
    proposals, solvers = [], []
    for c in range(no_chains):
        seed0 = sampler_seed0(c, no_chains)
        gen = None if generators is None else generators[c]
        proposals.append(RandomWalkProposal(d, seed=seed0 + 1, generator=gen))
        solvers.append(LocalSolverProxy(solver_factory()))
    current = [None] * no_chains if initial_samples is None else [np.array(x) for x in initial_samples]
    stages, global_step = [], 0
    for i, st in enumerate(conf["samplers_list"]):
        chains = []
        for c in range(no_chains):
            std = st["proposal_std"][c] if st.get("proposal_differs", False) else st["proposal_std"]
            proposals[c].set_covariance(std)
            init = st["initial_sample"] if "initial_sample" in st else current[c]
            uni = None if generators is None else generators[c]
            chains.append(Chain(problem, proposals[c], solvers[c], kind=st["type"],
                                seed=sampler_seed0(c, no_chains) + 2 + i, initial_sample=init,
                                surrogate=None if shared is None else SurrogateProxy(shared),
                                surrogate_is_updated=st["surrogate_is_updated"], uniform=uni))
        for ch in chains:
            ch.prepare()
        n_steps = st["max_samples"]
        for k in range(n_steps):
            for ch in chains:
                ch.step(k)
            global_step += 1
            if shared is not None and st["surrogate_is_updated"] and ((k + 1) % refit_every == 0 or k == n_steps - 1):
                shared.refit_if_pending(global_step)
        for c, ch in enumerate(chains):
            ch.finish()
            current[c] = ch.current
        stages.append(chains if keep_chains else None)
    return stages, shared
