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
    def __init__(self, no_parameters, no_observations, kernel_type=0, precision="f64"):
        self.no_parameters = no_parameters
        self.no_observations = no_observations
        self.kernel_type = kernel_type
        if precision not in ("f64", "f32"):
            raise ValueError("precision must be 'f64' (the reference's arithmetic) or 'f32'")
        self.precision = precision          # additive: 'f32' evaluates on the FP32 pipe (csrc/sdb_eval.cu, rbf_eval32_kernel)
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
        fn = L.lib().sdb_rbf_eval_f32 if self.precision == "f32" else L.lib().sdb_rbf_eval
        L.check(fn(L.ptr(pts), N, self.no_parameters, self.no_observations, C.byref(mo.struct), L.ptr(out), L.stream_handle(stream)))
        return out.cpu().numpy() if was_numpy else out


class Surrogate_update:
    def __init__(self, no_parameters, no_observations, initial_iteration=None, no_keep=None, expensive=False,
                 kernel_type=0, solver_tol_exp=-6, solver_type="minres", max_iterations=None, drop_coincident=False):
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
        # additive: remove exactly coincident snapshots before the solve (the lockstep sampler switches it on: all chains
        # leaving one start point emit the same displaced snapshot, which makes the saddle matrix singular)
        self.drop_coincident = bool(drop_coincident)
        self.fallbacks = 0                  # null-space Cholesky rounds that fell back to the pivoted LU
        self.round_failures = 0             # fixed-shape rounds whose solve failed (previous surrogate kept)
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
    def _thin(self):
        n, d = self._par.shape
        keep = torch.empty((n,), dtype=torch.uint8, device="cuda")
        work = torch.empty((L.lib().sdb_rbf_thin_work(n),), dtype=torch.float64, device="cuda")
        no_keep = n if self.no_keep is None else int(self.no_keep)
        L.check(L.lib().sdb_rbf_thin_masked(L.ptr(self._par), n, d, no_keep, None, int(self.drop_coincident), L.ptr(keep),
                                            L.ptr(work), L.stream_handle()))
        mask = keep.bool()
        self._par = self._par[mask].contiguous()
        self._obs = self._obs[mask].contiguous()

    def _scratch(self, need):
        if self._work is None or self._work.numel() < need:
            self._work = torch.empty((need,), dtype=torch.float64, device="cuda")
        return self._work

    def uses_cluster_kernels(self):
        """True when the dense solve is the LU with partial pivoting (solver_type 'direct', or a kernel type the null-space
        Cholesky does not cover): its panel kernel is a 16-CTA thread-block cluster, which cannot be placed on the scattered SMs a
        chain kernel leaves to an asynchronous collector -- such a collector must not ask the chains to reserve SMs (its round then
        runs when a chain block ends, like the synchronous one, but still off the host's critical path)."""
        return not (self.solver_type == "nullspace" and int(self.kernel_type) in (0, 1, 4, 5, 6))

    def _solve_dense(self, Xb, Yb, Cb, n, info):
        """Dense solve of the saddle system for the n rows of (Xb, Yb) into Cb; launches only (info stays on the device)."""
        d, m = self.no_parameters, self.no_observations
        if self.solver_type == "nullspace" and int(self.kernel_type) in (0, 1, 4, 5, 6) and n > d + 1:
            work = self._scratch(L.lib().sdb_rbf_fit_nullspace_work(n, d, m))
            L.check(L.lib().sdb_rbf_fit_nullspace(L.ptr(Xb), L.ptr(Yb), n, d, m, int(self.kernel_type), L.ptr(Cb), L.ptr(work),
                                                  L.ptr(info), L.stream_handle()))
            return "nullspace"
        work = self._scratch(L.lib().sdb_rbf_fit_work(n, d, m))
        L.check(L.lib().sdb_rbf_fit(L.ptr(Xb), L.ptr(Yb), n, d, m, int(self.kernel_type), L.ptr(Cb), L.ptr(work), L.ptr(info),
                                    L.stream_handle()))
        return "lu"

    def update_device(self, stream=None):
        """-> (DeviceModel, no_snapshots); nothing leaves HBM.  A singular system raises np.linalg.LinAlgError, as
        np.linalg.solve does at surrogate_rbf.py:130 (the updater's state then holds the offending rows, as there)."""
        if stream is not None and stream != torch.cuda.current_stream():
            with torch.cuda.stream(stream):          # every torch op and every kernel below goes to ONE stream
                return self.update_device()
        d, m = self.no_parameters, self.no_observations
        self._par = torch.cat([self._par] + self._new_par, dim=0).contiguous()
        self._obs = torch.cat([self._obs] + self._new_obs, dim=0).contiguous()
        self._new_par, self._new_obs = [], []
        if (self.no_keep is not None and self._par.shape[0] > self.no_keep) or (self.drop_coincident and self._par.shape[0] > 1):
            self._thin()
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
        if self.solver_type in ("direct", "nullspace"):
            how = self._solve_dense(Xb, Yb, Cb, n, info)
            failed = int(info[0].item()) != 0
            if failed and how == "nullspace":
                # the projected kernel matrix is numerically not definite: the pivoted LU of the reference's 'direct' solver
                self.fallbacks += 1
                info.zero_()
                work = self._scratch(L.lib().sdb_rbf_fit_work(n, d, m))
                L.check(L.lib().sdb_rbf_fit(L.ptr(Xb), L.ptr(Yb), n, d, m, int(self.kernel_type), L.ptr(Cb), L.ptr(work),
                                            L.ptr(info), L.stream_handle()))
                failed = int(info[0].item()) != 0
            if failed or not bool(torch.isfinite(Cb).all().item()):
                self.last_info = info.cpu().numpy()
                raise np.linalg.LinAlgError("Singular matrix")          # what np.linalg.solve raises
        else:
            S = torch.empty((N, N), dtype=torch.float64, device="cuda")
            rhs = torch.empty((m, N), dtype=torch.float64, device="cuda")       # one contiguous column per observation
            L.check(L.lib().sdb_rbf_system(L.ptr(Xb), L.ptr(Yb), n, d, m, int(self.kernel_type), L.ptr(S), L.ptr(rhs),
                                           L.stream_handle()))
            x = torch.zeros((m, N), dtype=torch.float64, device="cuda")
            x0 = None
            if self.initial_iteration is not None:
                x0 = L.dev(np.asarray(self.initial_iteration, dtype=float).reshape(-1))
            work = torch.empty((8 * N + 64,), dtype=torch.float64, device="cuda")
            maxiter = int(self.max_iterations) if self.max_iterations is not None else 5 * N   # scipy's default
            for k in range(m):
                L.check(L.lib().sdb_minres(L.ptr(S), N, L.ptr(rhs[k]), L.ptr(x0), L.ptr(x[k]), float(10.0 ** self.solver_tol_exp),
                                           maxiter, L.ptr(work), L.ptr(info), L.stream_handle()))
            Cb.copy_(x.t())
        COEFS = Cb.clone()
        self._info = info
        mo = DeviceModel(L.MODEL_RBF, kernel_type=self.kernel_type, coefs=COEFS, centers=self._par.clone(), d=d, m=m)
        return mo, n

    # ---- fixed-shape round: no host synchronisation ----------------------------------------------------------------
    # Once the updater holds exactly no_keep rows, a refit has static shapes: candidates = the held rows + world x quota
    # gathered snapshots, masked greedy thinning back to no_keep, dense solve.  State and result travel as one packed
    # buffer [COEFS | centers | their observations | status] (sections padded to 128 bytes), which is also what the
    # collector broadcasts, so every rank can take the next round's refit duty.
    @staticmethod
    def pack_layout(n, d, m):
        pad = lambda v: (v + 15) // 16 * 16
        o_c = 0
        o_x = o_c + pad((n + d + 1) * m)
        o_y = o_x + pad(n * d)
        o_s = o_y + pad(n * m)
        return o_c, o_x, o_y, o_s, o_s + 16

    def round_ready(self):
        return (self.no_keep is not None and self._par.shape[0] == int(self.no_keep) and not self._new_par
                and self.solver_type in ("direct", "nullspace"))

    def pack_views(self, pack, n=None):
        n = int(self.no_keep) if n is None else n
        d, m = self.no_parameters, self.no_observations
        o_c, o_x, o_y, o_s, _ = self.pack_layout(n, d, m)
        return (pack[o_c:o_c + (n + d + 1) * m].view(n + d + 1, m), pack[o_x:o_x + n * d].view(n, d),
                pack[o_y:o_y + n * m].view(n, m), pack[o_s:o_s + 16])

    def adopt_pack(self, pack):
        """Make a packed state (own refit or another rank's broadcast) this updater's state; -> DeviceModel."""
        coefs, X, Y, _ = self.pack_views(pack)
        self._par, self._obs = X, Y
        self.no_processed = X.shape[0]
        self._pack_cur = pack
        return DeviceModel(L.MODEL_RBF, kernel_type=self.kernel_type, coefs=coefs, centers=X, d=self.no_parameters,
                           m=self.no_observations)

    def new_pack(self):
        n, d, m = int(self.no_keep), self.no_parameters, self.no_observations
        return torch.zeros((self.pack_layout(n, d, m)[4],), dtype=torch.float64, device="cuda")

    def pack_from_state(self, model):
        """The packed form of the current (general-path) state and model: entry into the fixed-shape rounds."""
        pack = self.new_pack()
        coefs, X, Y, _ = self.pack_views(pack)
        coefs.copy_(model.coefs)
        X.copy_(self._par)
        Y.copy_(self._obs)
        return pack

    def update_round_device(self, gathered, world, quota, pack_cur, pack_out):
        """One fixed-shape refit on the current stream.  gathered: [world x (quota + 1) x (d + m)] blocks of
        sdb_snapshot_select.  pack_cur: the state to start from; pack_out receives the new state -- or a copy of pack_cur
        when the solve failed (status[0] != 0: singular / not definite / non-finite coefficients / too few distinct rows),
        so the chains never see a broken surrogate; the host reads the status later (`poll_status`)."""
        n0, d, m = int(self.no_keep), self.no_parameters, self.no_observations
        ncand = n0 + world * quota
        rb = getattr(self, "_round_bufs", None)
        if rb is None or rb["key"] != (ncand, n0):
            rb = dict(key=(ncand, n0), X=torch.empty((ncand, d), dtype=torch.float64, device="cuda"),
                      Y=torch.empty((ncand, m), dtype=torch.float64, device="cuda"),
                      alive=torch.empty((ncand,), dtype=torch.uint8, device="cuda"),
                      keep=torch.empty((ncand,), dtype=torch.uint8, device="cuda"),
                      work=torch.empty((L.lib().sdb_rbf_thin_work(ncand),), dtype=torch.float64, device="cuda"),
                      total=torch.zeros((1,), dtype=torch.int64, device="cuda"),
                      count=torch.zeros((1,), dtype=torch.int32, device="cuda"),
                      info=torch.zeros((4,), dtype=torch.int32, device="cuda"), new=self.new_pack())
            self._round_bufs = rb
        _, X0, Y0, _ = self.pack_views(pack_cur)
        coefs, Xn, Yn, status = self.pack_views(rb["new"])
        h, st = L.lib(), L.stream_handle()
        L.check(h.sdb_round_candidates(L.ptr(X0), L.ptr(Y0), n0, L.ptr(gathered), world, quota, d, m, L.ptr(rb["X"]), L.ptr(rb["Y"]),
                                       L.ptr(rb["alive"]), L.ptr(rb["total"]), st))
        L.check(h.sdb_rbf_thin_masked(L.ptr(rb["X"]), ncand, d, n0, L.ptr(rb["alive"]), 0, L.ptr(rb["keep"]), L.ptr(rb["work"]), st))
        L.check(h.sdb_rows_compact(L.ptr(rb["X"]), L.ptr(rb["Y"]), L.ptr(rb["keep"]), ncand, d, m, n0, L.ptr(Xn), L.ptr(Yn),
                                   L.ptr(rb["count"]), st))
        rb["info"].zero_()
        self._solve_dense(Xn, Yn, coefs, n0, rb["info"])
        bad = (rb["info"][0] != 0) | (rb["count"][0] != n0) | ~torch.isfinite(coefs).all()
        status.zero_()
        status[0] = bad.to(torch.float64)
        status[1] = rb["info"][0].to(torch.float64)
        status[2] = rb["count"][0].to(torch.float64)
        status[3] = rb["total"][0].to(torch.float64)
        torch.where(bad, pack_cur, rb["new"], out=pack_out)
        pack_out[self.pack_layout(n0, d, m)[3]:].copy_(status)      # the status always describes THIS round
        return pack_out

    # ---- the same round in two stages (PipelinedCollector) -------------------------------------------------------------
    # thin_stage: the cheap, strictly sequential part (candidates, greedy thinning, compaction -- one SM for ~1 ms);
    # solve_stage: the dense solve, which depends on the thinned set of ITS round only.  Round b+1 can therefore be
    # thinned while round b is still being solved, and the solves of consecutive rounds can run on different ranks.
    def result_layout(self):
        n, d, m = int(self.no_keep), self.no_parameters, self.no_observations
        o_s = ((n + d + 1) * m + 15) // 16 * 16
        return o_s, o_s + 16

    def new_result(self):
        """[COEFS | status]: what the rank on solve duty broadcasts."""
        return torch.zeros((self.result_layout()[1],), dtype=torch.float64, device="cuda")

    def new_thin_state(self):
        """(X [no_keep x d], Y [no_keep x m], meta [4] = rows kept, snapshots offered): the center set after a round."""
        n, d, m = int(self.no_keep), self.no_parameters, self.no_observations
        return (torch.zeros((n, d), dtype=torch.float64, device="cuda"), torch.zeros((n, m), dtype=torch.float64, device="cuda"),
                torch.zeros((4,), dtype=torch.float64, device="cuda"))

    def thin_stage(self, X0, Y0, gathered, world, quota, Xn, Yn, meta):
        n0, d, m = int(self.no_keep), self.no_parameters, self.no_observations
        ncand = n0 + world * quota
        tb = getattr(self, "_thin_bufs", None)
        if tb is None or tb["key"] != (ncand, n0):
            tb = dict(key=(ncand, n0), X=torch.empty((ncand, d), dtype=torch.float64, device="cuda"),
                      Y=torch.empty((ncand, m), dtype=torch.float64, device="cuda"),
                      alive=torch.empty((ncand,), dtype=torch.uint8, device="cuda"),
                      keep=torch.empty((ncand,), dtype=torch.uint8, device="cuda"),
                      work=torch.empty((L.lib().sdb_rbf_thin_work(ncand),), dtype=torch.float64, device="cuda"),
                      total=torch.zeros((1,), dtype=torch.int64, device="cuda"),
                      count=torch.zeros((1,), dtype=torch.int32, device="cuda"))
            self._thin_bufs = tb
        h, st = L.lib(), L.stream_handle()
        L.check(h.sdb_round_candidates(L.ptr(X0), L.ptr(Y0), n0, L.ptr(gathered), world, quota, d, m, L.ptr(tb["X"]), L.ptr(tb["Y"]),
                                       L.ptr(tb["alive"]), L.ptr(tb["total"]), st))
        L.check(h.sdb_rbf_thin_masked(L.ptr(tb["X"]), ncand, d, n0, L.ptr(tb["alive"]), 0, L.ptr(tb["keep"]), L.ptr(tb["work"]), st))
        L.check(h.sdb_rows_compact(L.ptr(tb["X"]), L.ptr(tb["Y"]), L.ptr(tb["keep"]), ncand, d, m, n0, L.ptr(Xn), L.ptr(Yn),
                                   L.ptr(tb["count"]), st))
        meta.zero_()
        meta[0] = tb["count"][0].to(torch.float64)
        meta[1] = tb["total"][0].to(torch.float64)

    def solve_stage(self, X, Y, meta, result):
        """Dense solve of the rows (X, Y) into result = [COEFS | status]; status[0] != 0: keep the previous surrogate."""
        n0, d, m = int(self.no_keep), self.no_parameters, self.no_observations
        sb = getattr(self, "_solve_bufs", None)
        if sb is None or sb["key"] != (n0, d, m):
            # fixed staging buffers: one CUDA graph of the solver serves every slot of the collector's rings
            sb = dict(key=(n0, d, m), X=torch.empty((n0, d), dtype=torch.float64, device="cuda"),
                      Y=torch.empty((n0, m), dtype=torch.float64, device="cuda"),
                      C=torch.empty((n0 + d + 1, m), dtype=torch.float64, device="cuda"),
                      info=torch.zeros((4,), dtype=torch.int32, device="cuda"))
            self._solve_bufs = sb
        sb["X"].copy_(X)
        sb["Y"].copy_(Y)
        sb["info"].zero_()
        self._solve_dense(sb["X"], sb["Y"], sb["C"], n0, sb["info"])
        o_s, _ = self.result_layout()
        bad = (sb["info"][0] != 0) | (meta[0] != n0) | ~torch.isfinite(sb["C"]).all()
        result[:(n0 + d + 1) * m].copy_(sb["C"].reshape(-1))
        status = result[o_s:o_s + 16]
        status.zero_()
        status[0] = bad.to(torch.float64)
        status[1] = sb["info"][0].to(torch.float64)
        status[2:4] = meta[0:2]
        status[4] = X.sum() + Y.sum()          # checksum of the rows these coefficients belong to (see assemble)

    def assemble(self, result, X, Y, pack_prev, pack_out):
        """pack_out = [COEFS of result | X | Y | status], or a copy of pack_prev when the round's solve failed -- or when
        the coefficients were fitted to OTHER rows than this rank's thinned set (every rank thins for itself; the checksum
        the solving rank sends along catches a rank whose thinning diverged, status[5] = 1)."""
        n0, d, m = int(self.no_keep), self.no_parameters, self.no_observations
        o_c, o_x, o_y, o_s, _ = self.pack_layout(n0, d, m)
        r_s, _ = self.result_layout()
        status = result[r_s:r_s + 16]
        mismatch = status[4] != (X.sum() + Y.sum())
        status[5] = mismatch.to(torch.float64)
        status[0] = torch.maximum(status[0], status[5])
        pack_out[o_c:o_c + (n0 + d + 1) * m].copy_(result[:(n0 + d + 1) * m])
        pack_out[o_x:o_x + n0 * d].copy_(X.reshape(-1))
        pack_out[o_y:o_y + n0 * m].copy_(Y.reshape(-1))
        torch.where(status[0] != 0, pack_prev, pack_out, out=pack_out)
        pack_out[o_s:].copy_(status)                                   # the status always describes THIS round

    def model_of_pack(self, pack):
        coefs, X, _, _ = self.pack_views(pack)
        return DeviceModel(L.MODEL_RBF, kernel_type=self.kernel_type, coefs=coefs, centers=X, d=self.no_parameters,
                           m=self.no_observations)

    def adopt_thin_state(self, X, Y):
        """The updater's rows := a thinned center set (the end of a pipelined stage)."""
        self._par, self._obs = X, Y
        self.no_processed = X.shape[0]

    def update(self):
        mo, n = self.update_device()
        self.last_info = self._info.cpu().numpy()
        return [mo.coefs.cpu().numpy(), mo.centers.cpu().numpy()], n
