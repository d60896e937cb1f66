"""Initial samples: Latin hypercube in the unit cube mapped through the prior's normal quantiles.

Host-side (runs once, before any chain step); same design as surrDAMH/modules/lhs_normal.py:13-39 so
that chain c of a lockstep run starts where sampler rank c of the reference starts:
  * RandomState(seed); five candidate designs, each = one block of uniforms followed by one
    permutation per parameter, cell = (permutation + uniform) / n
  * a candidate is accepted when no two of its points coincide; the reference intends to keep the
    candidate with the largest minimal distance but never updates its running maximum, so the LAST
    accepted candidate is the one returned (:31-33) -- reproduced
  * a 2-D prior_std is a covariance: only its marginal standard deviations are used (:35-37)
Memory is O(n^2) in the reference; here the coincidence check is done on sorted coordinates, O(n log n).
"""
import numpy as np
from scipy.stats import norm


def _has_coincident_points(unit):
    """True when two rows are identical (the only way the reference's min squared distance is 0)."""
    order = np.lexsort(unit.T[::-1])
    s = unit[order]
    return bool(np.any(np.all(s[1:] == s[:-1], axis=1)))


def lhs_normal(no_parameters, prior_mean, prior_std, n, seed):
    rs = np.random.RandomState(seed)
    design = np.zeros((n, no_parameters))
    for _ in range(5):
        jitter = rs.uniform(size=(n, no_parameters))
        cells = np.empty((n, no_parameters))
        for j in range(no_parameters):
            cells[:, j] = rs.permutation(n)
        cand = (cells + jitter) / n
        if not _has_coincident_points(cand):
            design = cand
    std = np.asarray(prior_std, dtype=float)
    if std.ndim == 2:
        std = np.sqrt(np.diag(std))
    return norm.ppf(design, loc=prior_mean, scale=std)
