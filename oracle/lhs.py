"""TEST INFRASTRUCTURE -- numpy restatement of the Latin-hypercube initial samples.

Follows /root/reference/surrDAMH/modules/lhs_normal.py:13-39.  The random stream
is consumed exactly as there (per try: one uniform block, then one permutation
per parameter).  The reference never updates its `maxmin`, so every try whose
minimum pairwise squared distance is > 0 overwrites the result and the LAST such
try wins (:31-33); a try with two coincident points is skipped.  O(n^2) memory,
small n only (SURVEY.md section 7 trap 11).
"""
import numpy as np
from scipy.stats import norm


def lhs_normal(no_parameters, prior_mean, prior_std, n, seed):
    rs = np.random.RandomState(seed)
    chosen = np.zeros([n, n])
    for _ in range(5):
        jitter = rs.uniform(size=[n, no_parameters])
        strata = np.zeros([n, no_parameters])
        for j in range(no_parameters):
            strata[:, j] = rs.permutation(n)
        unit = (strata + jitter) / n
        sq = np.zeros([n, n])
        for j in range(no_parameters):
            gap = unit[None, :, j] - unit[:, None, j]
            sq = sq + np.multiply(gap, gap)
        if np.min(sq + np.eye(n) * 1000) > 0:
            chosen = unit
    if np.array(prior_std).ndim == 2:
        prior_std = np.sqrt(np.diag(prior_std))
    return norm.ppf(chosen, loc=prior_mean, scale=prior_std)
