"""TEST INFRASTRUCTURE -- numpy restatement of the polynomial (normalised Hermite)
surrogate of the reference.

Follows /root/reference/surrDAMH/modules/surrogate_poly.py:
  multi_indices         <- generate_polynomials_degree (:112-133)
  hermite_table         <- hermite_poly_normalized     (:135-148)
  univariate_values     <- poly_eval_multi             (:162-175)
  basis_matrix          <- the PHI product loop        (:26-28, :76-79, :104-107)
  PolyApply.apply       <- Surrogate_apply.apply       (:18-30)
  PolyUpdate.add_data / .update <- Surrogate_update    (:54-97)
The scalar operation order of the reference is kept wherever it determines the
bits (monomial accumulation, left-to-right basis product, distinct arrays in the
normal-equation products so numpy picks gemm, not syrk).
"""
import itertools
import math

import numpy as np


def multi_indices(dim, degree):
    """All exponent vectors of total degree <= `degree`, in the reference's order:
    the zero vector, the d unit vectors in identity order, then each higher total
    degree as a block sorted ascending lexicographically (np.unique(axis=0) at :131)."""
    blocks = [np.zeros((1, dim), dtype=int)]
    if degree >= 1:
        blocks.append(np.eye(dim, dtype=int))
    for g in range(2, degree + 1):
        level = [a for a in itertools.product(range(g + 1), repeat=dim) if sum(a) == g]
        blocks.append(np.array(sorted(level), dtype=int).reshape(-1, dim))
    return np.vstack(blocks)


def hermite_table(degree):
    """Row k = ascending-power coefficients of He_k / sqrt(k!) (probabilists' Hermite).

    He_k = x*He_{k-1} - He'_{k-1}; integer coefficients are exact, one division by
    np.sqrt(k!) per row as at :146-147."""
    n = degree + 1
    H = np.zeros((n, n))
    H[0, 0] = 1
    if degree == 0:
        return H
    H[1, 1] = 1
    k = np.arange(1, n)
    for i in range(2, n):
        H[i, 1:] += H[i - 1, :-1]            # x * He_{i-1}
        H[i, :-1] -= k * H[i - 1, 1:]        # - d/dx He_{i-1}
    for i in range(n):
        H[i, :] = np.divide(H[i, :], np.sqrt(math.factorial(i)))
    return H


def univariate_values(table, points):
    """values[a, j, k] = sum_i table[k, i] * points[a, j]**i, accumulated i = 0, 1, ...
    with powers formed by repeated multiplication (:170-174)."""
    n_poly, n_coef = table.shape
    values = np.empty(points.shape + (n_poly,))
    values[...] = table[:, 0]
    power = np.ones(points.shape)
    for i in range(1, n_coef):
        power = power * points
        values += table[:, i] * power[..., None]
    return values


def basis_matrix(table, alpha, points):
    """PHI[a, t] = prod_j h_{alpha[t, j]}(points[a, j]), multiplied j = 0..d-1 from 1."""
    vals = univariate_values(table, points)
    PHI = np.ones((points.shape[0], alpha.shape[0]))
    for j in range(points.shape[1]):
        PHI *= vals[:, j, alpha[:, j]]
    return PHI


class PolyApply:
    """SOL = [MAT (no_poly x m), degree]; the apply-side degree only grows (:22-24)."""

    def __init__(self, no_parameters, no_observations):
        self.no_parameters, self.no_observations = no_parameters, no_observations
        self.current_degree = 1
        self._rebuild()

    def _rebuild(self):
        self.poly = multi_indices(self.no_parameters, self.current_degree)
        self.no_poly = self.poly.shape[0]
        self.hermite = hermite_table(self.current_degree)

    def apply(self, SOL, datapoints):
        MAT, degree = SOL
        pts = datapoints.reshape(-1, self.no_parameters)
        if degree > self.current_degree:
            self.current_degree = degree
            self._rebuild()
        return np.matmul(basis_matrix(self.hermite, self.poly, pts), MAT)


class PolyUpdate:
    """Least squares over ALL snapshots so far: MAT = pinv(PHI^T PHI) PHI^T Y (:93-95).

    degree = min(floor(ln n / ln d), max_degree), never decreasing (:71-75);
    snapshot weights are ignored (set to 1 at :63)."""

    def __init__(self, no_parameters, no_observations, max_degree=5):
        self.no_parameters, self.no_observations = no_parameters, no_observations
        self.max_degree = max_degree
        self.processed_par = np.empty((0, no_parameters))
        self.processed_obs = np.empty((0, no_observations))
        self.pending_par = np.empty((0, no_parameters))
        self.pending_obs = np.empty((0, no_observations))
        self.no_processed = 0
        self.current_degree = 1
        self._rebuild()

    def _rebuild(self):
        self.poly = multi_indices(self.no_parameters, self.current_degree)
        self.no_poly = self.poly.shape[0]
        self.hermite = hermite_table(self.current_degree)

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
        n = self.no_processed + self.pending_par.shape[0]
        degree = int(np.floor(np.log(n) / np.log(self.no_parameters)))
        degree = min(degree, self.max_degree)
        if degree > self.current_degree:
            self.current_degree = degree
            self._rebuild()
        self.processed_par = np.vstack((self.processed_par, self.pending_par))
        self.processed_obs = np.vstack((self.processed_obs, self.pending_obs))
        self.no_processed = n
        self.pending_par = np.empty((0, self.no_parameters))
        self.pending_obs = np.empty((0, self.no_observations))
        PHI = basis_matrix(self.hermite, self.poly, self.processed_par)
        PHIw = PHI * np.ones((n, 1))            # weights == 1; a distinct array, as at :80,:109
        A = np.matmul(PHIw.transpose(), PHI)
        RHS = np.matmul(PHIw.transpose(), self.processed_obs)
        MAT = np.matmul(np.linalg.pinv(A), RHS)
        self.last_A, self.last_RHS = A, RHS      # kept for conditioning-aware tolerances in tests
        return [MAT, self.current_degree], n
