"""Polynomial (normalised Hermite) surrogate on the GPU, with the reference's interface.

Drop-in for surrDAMH/modules/surrogate_poly.py:
  Surrogate_apply(no_parameters, no_observations).apply(SOL, datapoints) -> [N, m]   (:10-35)
  Surrogate_update(no_parameters, no_observations, max_degree=5)                     (:37-110)
      .add_data(list_of_Snapshot) / .update() -> (SOL, no_snapshots),  SOL = [MAT, degree]
numpy in -> numpy out (reference behaviour); CUDA tensors in -> CUDA tensors out (what the
lockstep driver uses, no host round trip).  All arithmetic runs in csrc/sdb_eval.cu (apply)
and csrc/sdb_fit.cu (update); there is no CPU path.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from . import tables
from .device import DeviceModel


def _points(datapoints, d):
    """-> (contiguous CUDA [N x d], was_numpy)"""
    if isinstance(datapoints, torch.Tensor) and datapoints.is_cuda:
        return datapoints.reshape(-1, d).to(torch.float64).contiguous(), False
    return L.dev(np.asarray(datapoints, dtype=float).reshape(-1, d)), True


class Surrogate_apply:
    def __init__(self, no_parameters, no_observations):
        self.no_parameters = no_parameters
        self.no_observations = no_observations
        self.current_degree = 1            # only grows, as in the reference (:22-24)
        self._cache = (None, None)

    def device_model(self, SOL):
        """SOL ([MAT, degree] or a DeviceModel) -> DeviceModel; the last upload is cached."""
        if isinstance(SOL, DeviceModel):
            return SOL
        if SOL is None:
            raise TypeError("cannot unpack non-iterable NoneType object (no surrogate has been fitted yet)")
        if self._cache[0] is SOL:
            return self._cache[1]
        MAT, degree = SOL
        mo = DeviceModel.poly(MAT, int(degree), self.no_parameters)
        self._cache = (SOL, mo)
        return mo

    def apply(self, SOL, datapoints, stream=None):
        mo = self.device_model(SOL)
        if mo.degree > self.current_degree:
            self.current_degree = mo.degree
        pts, was_numpy = _points(datapoints, self.no_parameters)
        N = pts.shape[0]
        out = torch.empty((N, self.no_observations), dtype=torch.float64, device=pts.device)
        if N == 0:
            return out.cpu().numpy() if was_numpy else out
        L.check(L.lib().sdb_poly_eval(L.ptr(pts), N, self.no_parameters, self.no_observations, C.byref(mo.struct),
                                      L.ptr(out), L.stream_handle(stream)))
        return out.cpu().numpy() if was_numpy else out


class Surrogate_update:
    def __init__(self, no_parameters, no_observations, max_degree=5):
        L.require_cuda()
        self.no_parameters = no_parameters
        self.no_observations = no_observations
        self.max_degree = max_degree
        self.current_degree = 1
        self.no_processed = 0
        self._par = torch.empty((0, no_parameters), dtype=torch.float64, device="cuda")   # processed_par
        self._obs = torch.empty((0, no_observations), dtype=torch.float64, device="cuda")
        self._new_par, self._new_obs = [], []
        self.last_info = None

    # the reference exposes the processed samples as an attribute (process_COLLECTOR.py:115)
    @property
    def processed_par(self):
        return self._par.cpu().numpy()

    @property
    def processed_obs(self):
        return self._obs.cpu().numpy()

    def add_data(self, snapshots):
        if len(snapshots) == 0:
            return
        par = np.array([np.asarray(s.sample, dtype=float).reshape(-1) for s in snapshots])
        obs = np.array([np.asarray(s.G_sample, dtype=float).reshape(-1) for s in snapshots])
        self.add_arrays(par, obs)

    def add_arrays(self, par, obs):
        """Snapshots as arrays / CUDA tensors [k x d], [k x m] (weights are ignored, :63)."""
        p, _ = _points(par, self.no_parameters)
        o, _ = _points(obs, self.no_observations)
        self._new_par.append(p)
        self._new_obs.append(o)

    def update_device(self, stream=None):
        """-> (DeviceModel, no_snapshots); everything stays in HBM."""
        d, m = self.no_parameters, self.no_observations
        self._par = torch.cat([self._par] + self._new_par, dim=0)
        self._obs = torch.cat([self._obs] + self._new_obs, dim=0)
        self._new_par, self._new_obs = [], []
        n = self.no_processed = self._par.shape[0]
        self.current_degree = tables.fit_degree(n, d, self.max_degree, self.current_degree)
        deg = self.current_degree
        alpha = tables.multi_indices(d, deg)
        P = alpha.shape[0]
        MAT = torch.empty((P, m), dtype=torch.float64, device="cuda")
        mo = DeviceModel(L.MODEL_POLY, degree=deg, coefs=MAT, d=d, m=m, alpha=L.dev(alpha, torch.int32),
                         herm=L.dev(tables.hermite_table(deg)))
        work = torch.empty((L.lib().sdb_poly_fit_work(n, P, m),), dtype=torch.float64, device="cuda")
        info = torch.zeros((4,), dtype=torch.int32, device="cuda")
        L.check(L.lib().sdb_poly_fit(L.ptr(self._par), L.ptr(self._obs), n, d, m, C.byref(mo.struct), L.ptr(MAT),
                                     L.ptr(work), L.ptr(info), L.stream_handle(stream)))
        self._info = info     # [kept directions, sweeps, converged, -]
        return mo, n

    def update(self):
        mo, n = self.update_device()
        self.last_info = self._info.cpu().numpy()
        return [mo.coefs.cpu().numpy(), mo.degree], n
