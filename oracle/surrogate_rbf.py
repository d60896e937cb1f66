"""TEST INFRASTRUCTURE -- numpy restatement of the radial-basis-function surrogate
of the reference (RBF weights + linear polynomial tail).

Follows /root/reference/surrDAMH/modules/surrogate_rbf.py:
  radial_kernel        <- kernel                   (:146-173)
  pairwise_distance    <- the repmat loops         (:26-34, :91-96)
  RbfApply.apply       <- Surrogate_apply.apply    (:20-40)
  greedy_thinning      <- the no_keep loop         (:97-120)
  RbfUpdate.add_data / .update <- Surrogate_update (:62-137)
"""
import numpy as np
import scipy.sparse.linalg as splin


def radial_kernel(r, kernel_type):
    """phi(r), returned (the reference works in place).  Types: 0 r, 1 r^3, 2 r^5,
    3 r^7, 7 r^9, 4 exp(-r^2), 5 (r^2+1)^(-1/2).  Type 6 rebinds a local name and
    therefore leaves the distances untouched, i.e. behaves as type 0 (:167-170)."""
    odd_power = {1: 3, 2: 5, 3: 7, 7: 9}
    if kernel_type in odd_power:
        return np.power(r, odd_power[kernel_type])
    if kernel_type == 4:
        return np.exp(-np.power(r, 2))
    if kernel_type == 5:
        return np.power(np.power(r, 2) + 1, -0.5)
    return r


def pairwise_distance(A, B):
    """R[a, b] = sqrt(sum_k (B[b,k] - A[a,k])^2), accumulated k = 0..d-1 (no
    |a|^2+|b|^2-2ab trick: the reference sums squared differences directly)."""
    acc = np.zeros((A.shape[0], B.shape[0]))
    for k in range(A.shape[1]):
        diff = B[None, :, k] - A[:, None, k]
        acc = acc + np.power(diff, 2)
    return np.sqrt(acc)


class RbfApply:
    """SOL = [COEFS ((n+d+1) x m), centers (n x d)]: rows 0..n-1 RBF weights, rows
    n..n+d-1 the linear terms, row n+d the constant."""

    def __init__(self, no_parameters, no_observations, kernel_type=0):
        self.no_parameters, self.no_observations = no_parameters, no_observations
        self.kernel_type = kernel_type

    def apply(self, SOL, datapoints):
        COEFS, centers = SOL
        d = self.no_parameters
        pts = datapoints.reshape(-1, d)
        K = radial_kernel(pairwise_distance(pts, centers), self.kernel_type)
        tail = d + 1
        radial = np.matmul(K, COEFS[:-tail])
        linear = COEFS[-1] + np.matmul(pts, COEFS[-d - 1:-1])
        shape = (pts.shape[0], self.no_observations)
        return np.reshape(radial, shape) + np.reshape(linear, shape)


def greedy_thinning(D, no_keep):
    """Boolean keep-mask: repeatedly find the globally closest pair (first flat
    argmin of the off-diagonal distance matrix) and drop the point owning the ROW
    of that entry -- i.e. the lower-indexed point of the pair -- until no_keep
    remain (:98-113).  The reference's `expensive` flag changes nothing (:104-110
    only rebinds a local), so it is not a parameter here."""
    n = D.shape[0]
    big = 2 * np.max(D)
    M = D + big * np.eye(n)
    keep = np.ones(n, dtype=bool)
    for _ in range(n - no_keep):
        row = np.argmin(M) // n
        M[row, :] = big
        M[:, row] = big
        keep[row] = False
    return keep


class RbfUpdate:
    """Interpolation system [[phi(D), P], [P^T, 0]] COEFS = [Y; 0], P = [X, 1]
    (:121-128), solved by np.linalg.solve ('direct') or per-column MINRES with
    tolerance 10**solver_tol_exp (the default, :129-135)."""

    def __init__(self, no_parameters, no_observations, initial_iteration=None, no_keep=None,
                 expensive=False, kernel_type=0, solver_tol_exp=-6, solver_type="minres"):
        self.no_parameters, self.no_observations = no_parameters, no_observations
        self.initial_iteration = initial_iteration
        self.no_keep = no_keep
        self.expensive = expensive
        self.kernel_type = kernel_type
        self.solver_tol_exp = solver_tol_exp
        self.solver_type = solver_type
        self.processed_par = np.empty((0, no_parameters))
        self.processed_obs = np.empty((0, no_observations))
        self.pending_par = np.empty((0, no_parameters))
        self.pending_obs = np.empty((0, no_observations))
        self.no_processed = 0

    def add_data(self, snapshots):
        if len(snapshots) == 0:
            return
        par = np.array([np.asarray(s.sample, dtype=float).reshape(-1) for s in snapshots])
        obs = np.array([np.asarray(s.G_sample, dtype=float).reshape(-1) for s in snapshots])
        self.add_arrays(par, obs)

    def add_arrays(self, par, obs):
        self.pending_par = np.vstack((self.pending_par, par.reshape(-1, self.no_parameters)))
        self.pending_obs = np.vstack((self.pending_obs, obs.reshape(-1, self.no_observations)))

    def update(self):
        self.processed_par = np.vstack((self.processed_par, self.pending_par))
        self.processed_obs = np.vstack((self.processed_obs, self.pending_obs))
        self.pending_par = np.empty((0, self.no_parameters))
        self.pending_obs = np.empty((0, self.no_observations))
        D = pairwise_distance(self.processed_par, self.processed_par)
        if self.no_keep is not None:
            keep = greedy_thinning(D, self.no_keep)
            D = D[keep, :][:, keep]
            self.processed_par = self.processed_par[keep, :]
            self.processed_obs = self.processed_obs[keep, :]
        n = self.no_processed = self.processed_par.shape[0]
        tail = self.no_parameters + 1
        P = np.append(self.processed_par, np.ones((n, 1)), axis=1)
        top = np.append(radial_kernel(D, self.kernel_type), P, axis=1)
        bottom = np.append(P.transpose(), np.zeros((tail, tail)), axis=1)
        system = np.vstack([top, bottom])
        rhs = np.vstack([self.processed_obs, np.zeros((tail, self.no_observations))])
        if self.solver_type == "direct":
            COEFS = np.linalg.solve(system, rhs)
        else:
            COEFS = np.empty((n + tail, self.no_observations))
            for k in range(self.no_observations):
                COEFS[:, k] = splin.minres(system, rhs[:, k], x0=self.initial_iteration,
                                           rtol=pow(10, self.solver_tol_exp), show=False)[0]
        self.last_system, self.last_rhs = system, rhs   # for residual / conditioning checks in tests
        return [COEFS.copy(), self.processed_par.copy()], n
