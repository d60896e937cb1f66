"""Device-resident descriptions of forward models and of the Gaussian problem.

DeviceModel   -> the `sdb_model` struct (poly / rbf surrogate, analytic toy model, zero model)
DeviceProblem -> the `sdb_problem` struct (Problem_Gauss, classes_SAMPLER.py:257-316)

Both keep their torch tensors alive next to the ctypes struct that points into them.
"""
import numpy as np
import scipy.linalg
import torch

from . import _lib as L
from . import tables


class DeviceModel:
    def __init__(self, kind, **kw):
        self.kind = kind
        self.kernel_type = int(kw.get("kernel_type", 0))
        self.degree = int(kw.get("degree", 0))
        self.coefs = kw.get("coefs")
        self.centers = kw.get("centers")
        self.alpha = kw.get("alpha")
        self.herm = kw.get("herm")
        self.toy = kw.get("toy", (0.0, 0.0, 0.0))
        self.d, self.m = kw.get("d"), kw.get("m")
        s = L.Model()
        s.kind, s.kernel_type, s.degree = kind, self.kernel_type, self.degree
        s.n_terms = int(self.alpha.shape[0]) if self.alpha is not None else 0
        s.n_centers = int(self.centers.shape[0]) if self.centers is not None else 0
        s.coefs = L.ptr(self.coefs)
        s.centers = L.ptr(self.centers)
        s.alpha = L.ptr(self.alpha)
        s.herm = L.ptr(self.herm)
        s.toy_f, s.toy_L, s.toy_M = self.toy
        self.struct = s

    # ---- constructors -------------------------------------------------------------------------
    @staticmethod
    def none():
        return DeviceModel(L.MODEL_NONE)

    @staticmethod
    def zero(d=None, m=None):
        return DeviceModel(L.MODEL_ZERO, d=d, m=m)

    @staticmethod
    def toy(f=-0.1, L_=1.0, M=0.5):
        return DeviceModel(L.MODEL_TOY, toy=(float(f), float(L_), float(M)), d=2, m=1)

    @staticmethod
    def poly(MAT, degree, no_parameters, device=None):
        """SOL = [MAT, degree] of the polynomial surrogate."""
        MAT = L.dev(MAT, device=device)
        if MAT.dim() == 1:
            MAT = MAT.reshape(-1, 1)
        alpha = tables.multi_indices(no_parameters, int(degree))
        if alpha.shape[0] != MAT.shape[0]:
            raise L.SdbError(L.ERR_ARG, "MAT has %d rows, degree %d in %d parameters has %d terms"
                             % (MAT.shape[0], degree, no_parameters, alpha.shape[0]))
        return DeviceModel(L.MODEL_POLY, degree=degree, coefs=MAT, d=no_parameters, m=MAT.shape[1],
                           alpha=L.dev(alpha, torch.int32, device), herm=L.dev(tables.hermite_table(int(degree)), device=device))

    @staticmethod
    def rbf(COEFS, centers, kernel_type=0, device=None):
        """SOL = [COEFS, centers] of the RBF surrogate."""
        centers = L.dev(centers, device=device)
        COEFS = L.dev(COEFS, device=device)
        if COEFS.dim() == 1:
            COEFS = COEFS.reshape(-1, 1)
        n, d = centers.shape
        if COEFS.shape[0] != n + d + 1:
            raise L.SdbError(L.ERR_ARG, "COEFS has %d rows, expected n+d+1 = %d" % (COEFS.shape[0], n + d + 1))
        return DeviceModel(L.MODEL_RBF, kernel_type=kernel_type, coefs=COEFS, centers=centers, d=d, m=COEFS.shape[1])

    def sol_numpy(self):
        """The reference-shaped SOL list on the host."""
        if self.kind == L.MODEL_POLY:
            return [self.coefs.cpu().numpy(), self.degree]
        if self.kind == L.MODEL_RBF:
            return [self.coefs.cpu().numpy(), self.centers.cpu().numpy()]
        raise L.SdbError(L.ERR_ARG, "model kind %d has no SOL" % self.kind)


def _spec(x, n):
    """scalar -> filled vector; list -> array (1-D: standard deviations, 2-D: covariance)."""
    a = np.asarray(x, dtype=float)
    if a.ndim == 0:
        a = np.full((n,), float(a))
    return a


class DeviceProblem:
    """Gaussian prior + Gaussian noise.  1-D specs are standard deviations (the kernel gets the
    variances std**2, computed here exactly as numpy does at classes_SAMPLER.py:296,306); 2-D specs
    are covariance matrices, shipped as the LAPACK LU factors np.linalg.solve (dgesv) would compute,
    so the device solve follows the same elimination order (:300-303,:310-313)."""

    def __init__(self, no_parameters, prior_mean=0.0, prior_std=1.0, noise_std=1.0, no_observations=None,
                 observations=None, device=None):
        d = int(no_parameters)
        obs = np.asarray(observations, dtype=float).reshape(-1)
        m = int(no_observations) if no_observations is not None else obs.size
        if obs.size == 1 and m > 1:
            obs = np.full((m,), obs[0])
        self.d, self.m = d, m
        self.prior_mean_np = _spec(prior_mean, d)
        self.prior_spec_np = _spec(prior_std, d)
        self.noise_spec_np = _spec(noise_std, m)
        self.obs_np = obs
        self.t = {}
        s = L.Problem()
        s.d, s.m = d, m
        self.t["prior_mean"] = L.dev(self.prior_mean_np, device=device)
        self.t["obs"] = L.dev(obs, device=device)
        s.prior_mean, s.obs = L.ptr(self.t["prior_mean"]), L.ptr(self.t["obs"])
        for name, spec in (("prior", self.prior_spec_np), ("noise", self.noise_spec_np)):
            if spec.ndim == 2:
                lu, piv = scipy.linalg.lu_factor(spec)      # dgetrf, as inside dgesv
                self.t[name + "_scale"] = L.dev(lu, device=device)
                self.t[name + "_piv"] = L.dev(piv.astype(np.int32), torch.int32, device)
                full = 1
            else:
                self.t[name + "_scale"] = L.dev(spec ** 2, device=device)
                self.t[name + "_piv"] = None
                full = 0
            setattr(s, name + "_full", full)
            setattr(s, name + "_scale", L.ptr(self.t[name + "_scale"]))
            setattr(s, name + "_piv", L.ptr(self.t[name + "_piv"]))
        self.struct = s

    @property
    def prior_std_vector(self):
        """Marginal standard deviations (what the LHS initialiser uses, lhs_normal.py:35-37)."""
        sp = self.prior_spec_np
        return np.sqrt(np.diag(sp)) if sp.ndim == 2 else sp
