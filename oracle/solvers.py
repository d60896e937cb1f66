"""TEST INFRASTRUCTURE -- numpy restatement of the toy forward models ("solvers").

Follows /root/reference/examples/solvers/solver_examples.py:
  LinElast2Exp  <- Solver_linela2exp_local (:12-29)
  Uninformative <- Solver_uninformative    (:66-75)
Both expose the reference plugin interface set_parameters / get_observations
(README.md:61-64).
"""
import numpy as np


class LinElast2Exp:
    """1-D two-material bar: displacement at x=L for log-stiffness parameters (u1,u2).

    The chain of expressions below keeps the reference's order of operations
    (:24-29) so the value is bit-identical; algebraically it reduces to
    -0.0375*exp(-u1) - 0.0125*exp(-u2) for the default f, L, M.
    """

    def __init__(self, f=-0.1, L=1.0, M=0.5):
        self.f, self.L, self.M = f, L, M
        self.no_parameters, self.no_observations = 2, 1

    def set_parameters(self, data_par):
        self.k1 = np.exp(data_par[0])
        self.k2 = np.exp(data_par[1])

    def get_observations(self):
        f, L, M, k1, k2 = self.f, self.L, self.M, self.k1, self.k2
        D1 = (f * L) / k2
        C1 = D1 * k2 / k1
        D2 = -f / (2 * k1) * (M * M) + C1 * M + f / (2 * k2) * (M * M) - D1 * M
        return -f / (2 * k2) * (L * L) + D1 * L + D2


class Uninformative:
    """G == 0: the posterior equals the prior up to a constant (known-answer check)."""

    def __init__(self, no_parameters, no_observations):
        self.no_parameters, self.no_observations = no_parameters, no_observations

    def set_parameters(self, data_par):
        self.data_par = data_par

    def get_observations(self):
        return np.full((self.no_observations,), 0.0)


class LocalSolverProxy:
    """In-process stand-in for Solver_MPI_collector_MPI
    (classes_communication.py:48-81): send_parameters / recv_observations around a
    plugin instance, one request in flight, no MPI hop (which favours the CPU side
    when this is timed as the baseline)."""

    def __init__(self, solver):
        self.solver = solver
        self.no_calls = 0

    def send_parameters(self, parameters):
        self.solver.set_parameters(parameters)

    def recv_observations(self):
        self.no_calls += 1
        return self.solver.get_observations()


class SurrogateAsSolver:
    """A frozen surrogate (apply object + SOL) behind the plugin interface; used as
    the "true G" of the Darcy-shaped configuration (SURVEY.md section 8d cfg 5)."""

    def __init__(self, apply_obj, SOL):
        self.apply_obj, self.SOL = apply_obj, SOL

    def set_parameters(self, data_par):
        self.p = np.array(data_par)

    def get_observations(self):
        return self.apply_obj.apply(self.SOL, self.p)[0, :]
