"""Solver plugins with the reference's interface (README.md:61-64 of the reference:
`set_parameters(parameters)`, `get_observations()`), plus the OPTIONAL device protocol the lockstep
sampler looks for:

    device_model() -> surrdamh_b200.device.DeviceModel

A solver that offers device_model() is evaluated inside the fused chain kernel (no host round
trip per step); any other solver -- e.g. a user's FEM code -- is called on the host for the
proposals that survive the surrogate test, exactly once per true-model evaluation, through
set_parameters / get_observations (the split path of surrdamh_b200.sampler).

Counterparts of examples/solvers/solver_examples.py:12-29 (Solver_linela2exp_local) and :66-75
(Solver_uninformative) of the reference.
"""
import numpy as np


class Solver_linela2exp_local:
    """Two-material 1-D elastic bar: displacement at x = L for log-stiffnesses (u1, u2)."""

    def __init__(self, f=-0.1, L=1.0, M=0.5):
        self.f, self.L, self.M = f, L, M
        self.no_parameters, self.no_observations = 2, 1

    def set_parameters(self, data_par):
        self.k1, self.k2 = np.exp(data_par[0]), np.exp(data_par[1])   # numpy's exp, as the reference

    def get_observations(self):
        f, L, M = self.f, self.L, self.M
        uL_slope = f * L / self.k2                 # integration constants of the right / left segment
        left_slope = uL_slope * self.k2 / self.k1
        offset = -f / (2 * self.k1) * (M * M) + left_slope * M + f / (2 * self.k2) * (M * M) - uL_slope * M
        return np.float64(-f / (2 * self.k2) * (L * L) + uL_slope * L + offset)

    def device_model(self):
        from surrdamh_b200.device import DeviceModel
        return DeviceModel.toy(self.f, self.L, self.M)


class Solver_uninformative:
    """G == 0: the posterior equals the prior (known-answer configuration)."""

    def __init__(self, no_parameters, no_observations):
        self.no_parameters, self.no_observations = no_parameters, no_observations

    def set_parameters(self, data_par):
        self.data_par = data_par

    def get_observations(self):
        return np.zeros((self.no_observations,))

    def device_model(self):
        from surrdamh_b200.device import DeviceModel
        return DeviceModel.zero(self.no_parameters, self.no_observations)


class Solver_frozen_surrogate:
    """A frozen polynomial / RBF model behind the solver interface: the stand-in "true G" of the
    Darcy-shaped surrogate-only configuration (the external FEM solver is out of scope).

    kind 'rbf': SOL = [COEFS, centers], kernel_type;  kind 'poly': SOL = [MAT, degree]."""

    def __init__(self, no_parameters, no_observations, kind="rbf", SOL=None, kernel_type=1, npz=None):
        self.no_parameters, self.no_observations = no_parameters, no_observations
        self.kind, self.kernel_type = kind, kernel_type
        if npz is not None:
            with np.load(npz) as z:
                SOL = [z["coefs"], z["centers"]] if kind == "rbf" else [z["MAT"], int(z["degree"])]
        self.SOL = SOL
        self._model = None
        self._apply = None

    def device_model(self):
        from surrdamh_b200.device import DeviceModel
        if self._model is None:
            if self.kind == "rbf":
                self._model = DeviceModel.rbf(self.SOL[0], self.SOL[1], self.kernel_type)
            else:
                self._model = DeviceModel.poly(self.SOL[0], self.SOL[1], self.no_parameters)
        return self._model

    def set_parameters(self, data_par):
        self.data_par = np.asarray(data_par, dtype=float)

    def get_observations(self):
        # host interface kept for completeness; evaluates on the GPU through the surrogate classes
        if self._apply is None:
            from surrdamh_b200 import surrogate_poly, surrogate_rbf
            self._apply = (surrogate_rbf.Surrogate_apply(self.no_parameters, self.no_observations, self.kernel_type)
                           if self.kind == "rbf" else surrogate_poly.Surrogate_apply(self.no_parameters, self.no_observations))
        return self._apply.apply(self.device_model(), self.data_par)[0, :]
