"""Radial-basis-function surrogate (RBF weights + linear tail) on the GPU, reference interface.

Drop-in for surrDAMH/modules/surrogate_rbf.py:
  Surrogate_apply(no_parameters, no_observations, kernel_type=0).apply(SOL, datapoints)   (:13-40)
  Surrogate_update(no_parameters, no_observations, initial_iteration=None, no_keep=None,
                   expensive=False, kernel_type=0, solver_tol_exp=-6, solver_type='minres')  (:42-137)
      .add_data(list_of_Snapshot) / .update() -> (SOL, no_snapshots),  SOL = [COEFS, centers]
numpy in -> numpy out; CUDA tensors in -> CUDA tensors out.  Kernels: csrc/sdb_eval.cu
(tiled distance + kernel + K.W contraction), csrc/sdb_fit.cu (thinning, saddle-point system build),
csrc/sdb_lu.cu (LU with partial pivoting = solver_type 'direct'), csrc/sdb_minres.cu (the
reference's default iterative solver).  No CPU path.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L
from .device import DeviceModel
from .surrogate_poly import _points


class Surrogate_apply:
    def __init__(self, no_parameters, no_observations, kernel_type=0):
        self.no_parameters = no_parameters
        self.no_observations = no_observations
        self.kernel_type = kernel_type
        self._cache = (None, None)

    def device_model(self, SOL):
        if isinstance(SOL, DeviceModel):
            return SOL
        if SOL is None:
            raise TypeError("cannot unpack non-iterable NoneType object (no surrogate has been fitted yet)")
        if self._cache[0] is SOL:
            return self._cache[1]
        COEFS, centers = SOL
        mo = DeviceModel.rbf(COEFS, centers, self.kernel_type)
        self._cache = (SOL, mo)
        return mo

    def apply(self, SOL, datapoints, stream=None):
        mo = self.device_model(SOL)
        if mo.kernel_type != self.kernel_type:      # the apply side's own kernel_type rules (:36)
            mo = DeviceModel(L.MODEL_RBF, kernel_type=self.kernel_type, coefs=mo.coefs, centers=mo.centers, d=mo.d, m=mo.m)
        pts, was_numpy = _points(datapoints, self.no_parameters)
        N = pts.shape[0]
        out = torch.empty((N, self.no_observations), dtype=torch.float64, device=pts.device)
        if N == 0:
            return out.cpu().numpy() if was_numpy else out
        L.check(L.lib().sdb_rbf_eval(L.ptr(pts), N, self.no_parameters, self.no_observations, C.byref(mo.struct),
                                     L.ptr(out), L.stream_handle(stream)))
        return out.cpu().numpy() if was_numpy else out


class Surrogate_update:
    def __init__(self, no_parameters, no_observations, initial_iteration=None, no_keep=None, expensive=False,
                 kernel_type=0, solver_tol_exp=-6, solver_type="minres", max_iterations=None):
        L.require_cuda()
        self.no_parameters = no_parameters
        self.no_observations = no_observations
        self.initial_iteration = initial_iteration
        self.no_keep = no_keep
        self.expensive = expensive          # has no effect in the reference either (:104-110)
        self.kernel_type = kernel_type
        self.solver_tol_exp = solver_tol_exp
        self.solver_type = solver_type
        self.max_iterations = max_iterations
        self.no_processed = 0
        self._par = torch.empty((0, no_parameters), dtype=torch.float64, device="cuda")
        self._obs = torch.empty((0, no_observations), dtype=torch.float64, device="cuda")
        self._new_par, self._new_obs = [], []
        self.last_info = None
        self._work = None                   # persistent scratch: stable addresses let the LU graph be replayed
        self._io = None
        self._info_buf = torch.zeros((4,), dtype=torch.int32, device="cuda")

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
        p, _ = _points(par, self.no_parameters)
        o, _ = _points(obs, self.no_observations)
        self._new_par.append(p)
        self._new_obs.append(o)

    # ---- pieces (each one kernel family) ---------------------------------------------------------
    def _thin(self, stream):
        n, d = self._par.shape
        keep = torch.empty((n,), dtype=torch.uint8, device="cuda")
        work = torch.empty((L.lib().sdb_rbf_thin_work(n),), dtype=torch.float64, device="cuda")
        L.check(L.lib().sdb_rbf_thin(L.ptr(self._par), n, d, int(self.no_keep), L.ptr(keep), L.ptr(work),
                                     L.stream_handle(stream)))
        mask = keep.bool()
        self._par = self._par[mask].contiguous()
        self._obs = self._obs[mask].contiguous()

    def update_device(self, stream=None):
        """-> (DeviceModel, no_snapshots); nothing leaves HBM."""
        d, m = self.no_parameters, self.no_observations
        self._par = torch.cat([self._par] + self._new_par, dim=0).contiguous()
        self._obs = torch.cat([self._obs] + self._new_obs, dim=0).contiguous()
        self._new_par, self._new_obs = [], []
        if self.no_keep is not None and self._par.shape[0] > self.no_keep:
            self._thin(stream)
        n = self.no_processed = self._par.shape[0]
        N = n + d + 1
        info = self._info_buf
        info.zero_()
        # persistent input / output / scratch buffers: stable addresses let the solvers replay their CUDA graphs
        if self._io is None or self._io[0] != (n, d, m):
            self._io = ((n, d, m), torch.empty((n, d), dtype=torch.float64, device="cuda"),
                        torch.empty((n, m), dtype=torch.float64, device="cuda"),
                        torch.empty((N, m), dtype=torch.float64, device="cuda"))
        _, Xb, Yb, Cb = self._io
        Xb.copy_(self._par)
        Yb.copy_(self._obs)

        def scratch(need):
            if self._work is None or self._work.numel() < need:
                self._work = torch.empty((need,), dtype=torch.float64, device="cuda")
            return self._work

        done = False
        if self.solver_type == "nullspace" and int(self.kernel_type) in (0, 1, 4, 5, 6) and n > d + 1:
            # additive solver: null-space method + Cholesky (no pivoting); falls back to the pivoted LU below
            # if the projected kernel matrix is numerically not definite
            work = scratch(L.lib().sdb_rbf_fit_nullspace_work(n, d, m))
            L.check(L.lib().sdb_rbf_fit_nullspace(L.ptr(Xb), L.ptr(Yb), n, d, m, int(self.kernel_type), L.ptr(Cb), L.ptr(work),
                                                  L.ptr(info), L.stream_handle(stream)))
            done = int(info[0].item()) == 0
            self.fallbacks = getattr(self, "fallbacks", 0) + (0 if done else 1)
        if done:
            pass
        elif self.solver_type in ("direct", "nullspace"):
            work = scratch(L.lib().sdb_rbf_fit_work(n, d, m))
            L.check(L.lib().sdb_rbf_fit(L.ptr(Xb), L.ptr(Yb), n, d, m, int(self.kernel_type), L.ptr(Cb), L.ptr(work), L.ptr(info),
                                        L.stream_handle(stream)))
        else:
            S = torch.empty((N, N), dtype=torch.float64, device="cuda")
            rhs = torch.empty((m, N), dtype=torch.float64, device="cuda")       # one contiguous column per observation
            L.check(L.lib().sdb_rbf_system(L.ptr(Xb), L.ptr(Yb), n, d, m, int(self.kernel_type), L.ptr(S), L.ptr(rhs),
                                           L.stream_handle(stream)))
            x = torch.zeros((m, N), dtype=torch.float64, device="cuda")
            x0 = None
            if self.initial_iteration is not None:
                x0 = L.dev(np.asarray(self.initial_iteration, dtype=float).reshape(-1))
            work = torch.empty((8 * N + 64,), dtype=torch.float64, device="cuda")
            maxiter = int(self.max_iterations) if self.max_iterations is not None else 5 * N   # scipy's default
            for k in range(m):
                L.check(L.lib().sdb_minres(L.ptr(S), N, L.ptr(rhs[k]), L.ptr(x0), L.ptr(x[k]), float(10.0 ** self.solver_tol_exp),
                                           maxiter, L.ptr(work), L.ptr(info), L.stream_handle(stream)))
            Cb.copy_(x.t())
        COEFS = Cb.clone()
        self._info = info
        mo = DeviceModel(L.MODEL_RBF, kernel_type=self.kernel_type, coefs=COEFS, centers=self._par.clone(), d=d, m=m)
        return mo, n

    def update(self):
        mo, n = self.update_device()
        self.last_info = self._info.cpu().numpy()
        if self.solver_type in ("direct", "nullspace") and self.last_info[0] != 0:
            raise np.linalg.LinAlgError("Singular matrix")          # what np.linalg.solve raises
        return [mo.coefs.cpu().numpy(), mo.centers.cpu().numpy()], n
