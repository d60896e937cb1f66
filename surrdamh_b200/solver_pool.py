"""A pool of host solver workers for the true model G when it is NOT device-evaluable (a user's FEM code).

Counterpart of the reference's solver pool: surrDAMH/process_SOLVER.py:60-106 queues the samplers' requests and
dispatches them to `no_solvers` spawned solver processes (process_CHILD.py:37-50), each of which owns one instance of
the user's solver class and answers set_parameters / get_observations.  With lockstep chains the requests of a step
arrive together (the proposals that passed the surrogate test): a batch is split over the workers, every worker
calls the reference plugin interface once per parameter vector.  Batches are queued asynchronously (`submit` /
`collect`), so the sampler can keep the pool busy with one half of the chains while the GPU decides and re-proposes the
other half (sampler.LockstepSampler._run_split).

Workers are separate processes (the plugin may not be thread-safe and usually is not GIL-friendly); each imports the
solver module from `solver_module_path` itself, exactly as process_CHILD.py:13-25 does.
"""
import importlib.util as iu
import multiprocessing as mp

import numpy as np

_solver = None


def _init_worker(module_path, module_name, init_name, parameters, paths_to_append):
    global _solver
    import sys
    for p in paths_to_append or []:
        sys.path.append(p)
    spec = iu.spec_from_file_location(module_name, module_path)
    module = iu.module_from_spec(spec)
    spec.loader.exec_module(module)
    _solver = getattr(module, init_name)(**(parameters or {}))


def _evaluate_chunk(chunk):
    out = []
    for p in chunk:
        _solver.set_parameters(p)
        out.append(np.asarray(_solver.get_observations(), dtype=float).reshape(-1))
    return np.array(out)


class HostSolverPool:
    def __init__(self, no_solvers, solver_module_path, solver_module_name, solver_init_name, solver_parameters=None,
                 paths_to_append=None):
        self.no_solvers = int(no_solvers)
        ctx = mp.get_context("spawn")          # never fork a process that holds a CUDA context
        self.pool = ctx.Pool(self.no_solvers, initializer=_init_worker,
                             initargs=(solver_module_path, solver_module_name, solver_init_name, solver_parameters,
                                       paths_to_append))
        self.no_calls = 0

    @classmethod
    def from_conf(cls, conf):
        c = conf.conf
        return cls(conf.no_full_solvers, c["solver_module_path"], c["solver_module_name"], c["solver_init_name"],
                   c.get("solver_parameters", {}), c.get("paths_to_append", []))

    def submit(self, params):
        """Queue params [n x d] on the workers and return at once (process_SOLVER.py:60-106 queues the samplers'
        requests the same way); `collect` waits for the answers.  Several batches may be in flight."""
        params = np.asarray(params, dtype=float)
        n = params.shape[0]
        if n == 0:
            return None
        k = min(self.no_solvers, n)
        bounds = np.linspace(0, n, k + 1).astype(int)
        chunks = [params[bounds[i]:bounds[i + 1]] for i in range(k)]
        self.no_calls += n
        return self.pool.map_async(_evaluate_chunk, chunks)

    @staticmethod
    def collect(handle):
        """-> G [n x m], rows in the order submitted."""
        if handle is None:
            return np.empty((0, 0))
        return np.vstack(handle.get())

    def evaluate(self, params):
        """params [n x d] -> G [n x m], rows in the order given."""
        return self.collect(self.submit(params))

    def close(self):
        self.pool.close()
        self.pool.join()
