"""JSON configuration with the reference's schema (surrDAMH/configuration.py:16-92).

Same keys, same defaults, same attribute names, so a reference config file loads unchanged:
  required : no_parameters, no_observations, problem_parameters{prior_mean, prior_std, observations,
             noise_std}, solver_module_path, solver_module_name, solver_init_name, solver_parameters,
             no_solvers, samplers_list[{type, max_samples, time_limit, proposal_std,
             surrogate_is_updated, ?proposal_differs, ?initial_sample}]
  optional : surrogate_type ('rbf' | anything else -> poly; absent -> no surrogate),
             surr_solver_parameters, surr_updater_parameters, solver_parent_parameters,
             paths_to_append, saved_samples_name, initial_sample_type, debug
ADDITIVE keys of this implementation (ignored by the reference, never renames):
  no_chains            chains per GPU (default: the no_samplers argument)
  refit_every          lockstep steps between two surrogate refits (default 64)
  snapshots_per_refit  at most this many new snapshots enter one refit: every rank contributes
                       floor(q / world) of its own, evenly spaced over their (step, chain) order within the
                       block; the gathered set is ordered (rank, step, chain) (default: all, as the reference)
  refit_mode           'broadcast' (rank 0 refits, the packed state is broadcast; default) | 'rotate' (the
                       refit duty moves round-robin over the ranks) | 'replicated' (every rank refits)
  collector_async      true: the refit of block b runs on a side stream while the next blocks step; its
                       surrogate is used from block b + refit_lag + 1 on (default false: used from block b + 1)
  refit_lag            blocks of lag of the asynchronous collector (default 1)
  collector_pipelined  (with collector_async, default true) the round runs as two pipelined stages -- thinning on every
                       rank, the dense solve on the rank on duty -- so consecutive rounds overlap (PipelinedCollector)
  reserve_sms          SMs the chain kernel leaves free for the asynchronous collector's kernels (default 15; the
                       chain CTAs own their SMs, so the refit needs SMs of its own to run beside them).  One GPU must
                       solve every round within a block period: ~15 SMs at 4096 centers.  With refit_mode "rotate" on N
                       GPUs a rank solves every N-th round: 11 SMs with refit_lag 2 on two GPUs, 6 SMs with refit_lag 3
                       from four on leave the chains more SMs (what bench.py uses)
  surrogate_precision  'f64' (default, the reference's arithmetic) | 'f32': an RBF surrogate is evaluated on the FP32
                       pipe inside the chain kernel (stage 1 only screens; stage 2 keeps the chain exact)
  drop_coincident      remove exactly coincident snapshots before an RBF solve (default true: lockstep chains
                       leaving a common start point emit identical snapshots, a singular system in the reference)
  seed                 Philox key of the chain RNG (default 0)
  device_solver        false forces the host plugin path even if the solver offers device_model()
"""
import importlib.util as iu
import json
import os
import sys

from . import surrogate_poly, surrogate_rbf


class Configuration:
    def __init__(self, no_samplers, conf_name=None, conf=None):
        """conf_name: a path to a JSON file, or a bare name resolved as examples/<name>.json relative
        to the working directory (the reference's rule, configuration.py:20-24); conf: an already
        parsed dict (takes precedence)."""
        if conf is None:
            if conf_name is None:
                conf_name = "simple_MPI"
            conf_path = conf_name if os.path.isfile(conf_name) else "examples/" + conf_name + ".json"
            with open(conf_path) as f:
                conf = json.load(f)
            conf_name = os.path.splitext(os.path.basename(conf_path))[0]
        elif conf_name is None:
            conf_name = conf.get("saved_samples_name", "conf")
        self.conf = conf
        self.conf_name = conf_name

        # problem
        self.saved_samples_name = conf.get("saved_samples_name", conf_name)
        self.no_parameters = conf["no_parameters"]
        self.no_observations = conf["no_observations"]
        self.problem_parameters = conf["problem_parameters"]

        # solver plugin
        for path in conf.get("paths_to_append", []):
            sys.path.append(path)
        self.child_solver_parameters = conf.get("solver_parameters", {})
        self.child_solver_init = None
        if "solver_module_path" in conf:
            spec = iu.spec_from_file_location(conf["solver_module_name"], conf["solver_module_path"])
            module = iu.module_from_spec(spec)
            spec.loader.exec_module(module)
            self.child_solver_init = getattr(module, conf["solver_init_name"])

        # sampling
        self.no_full_solvers = conf.get("no_solvers", 1)
        self.no_samplers = no_samplers
        self.list_alg = conf["samplers_list"]
        self.initial_sample_type = conf.get("initial_sample_type", "lhs")

        # surrogate
        if "surrogate_type" in conf:
            self.use_surrogate = True
            self.rank_surr_collector = self.no_samplers + 1
            surr = surrogate_rbf if conf["surrogate_type"] == "rbf" else surrogate_poly
            self.surrogate_type = "rbf" if conf["surrogate_type"] == "rbf" else "poly"
            self.surr_solver_init = surr.Surrogate_apply
            self.surr_updater_init = surr.Surrogate_update
            base = {"no_parameters": self.no_parameters, "no_observations": self.no_observations}
            self.surr_solver_parameters = dict(base, **conf.get("surr_solver_parameters", {}))
            self.surr_updater_parameters = dict(base, **conf.get("surr_updater_parameters", {}))
        else:
            self.use_surrogate = False
            self.surrogate_type = None

        # other settings kept for attribute compatibility
        self.solver_parent_parameters = [dict({"no_parameters": self.no_parameters, "no_observations": self.no_observations,
                                               "problem_name": conf_name, "no_samplers": no_samplers},
                                              **conf.get("solver_parent_parameters", {}))] * self.no_full_solvers
        self.max_buffer_size = 1 << 30
        self.solver_parent_rank = self.no_samplers
        self.debug = conf.get("debug", False)

        # additive keys
        self.no_chains = int(conf.get("no_chains", no_samplers))
        self.refit_every = int(conf.get("refit_every", 64))
        self.snapshots_per_refit = conf.get("snapshots_per_refit", None)
        self.refit_mode = conf.get("refit_mode", "broadcast")
        self.collector_async = bool(conf.get("collector_async", False))
        self.refit_lag = int(conf.get("refit_lag", 1))
        self.collector_pipelined = bool(conf.get("collector_pipelined", True))
        self.reserve_sms = int(conf.get("reserve_sms", 15))
        self.drop_coincident = bool(conf.get("drop_coincident", True))
        self.surrogate_precision = conf.get("surrogate_precision", "f64")
        if self.surrogate_precision not in ("f64", "f32"):
            raise ValueError("surrogate_precision must be 'f64' or 'f32', got %r" % (self.surrogate_precision,))
        self.seed = int(conf.get("seed", 0))
        self.device_solver = bool(conf.get("device_solver", True))
