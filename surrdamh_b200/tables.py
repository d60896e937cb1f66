"""Host-side constant tables of the polynomial surrogate (tiny, built once per degree).

The kernels take the multi-index table `alpha [P x d]` and the normalised Hermite
coefficient table `herm [(deg+1) x (deg+1)]` as inputs; their ORDER and BITS define the
meaning of a coefficient matrix MAT, so they follow the reference exactly:
  multi_indices  -> surrDAMH/modules/surrogate_poly.py:112-133 (generate_polynomials_degree)
  hermite_table  -> surrDAMH/modules/surrogate_poly.py:135-148 (hermite_poly_normalized)
  fit_degree     -> surrDAMH/modules/surrogate_poly.py:71-75   (degree rule of update())
"""
import math

import numpy as np


def _compositions(total, parts):
    """All non-negative integer vectors of length `parts` summing to `total`, ascending
    lexicographically (what np.unique(axis=0) leaves at :131)."""
    if parts == 1:
        yield (total,)
        return
    for first in range(total + 1):
        for rest in _compositions(total - first, parts - 1):
            yield (first,) + rest


def multi_indices(dim, degree):
    """Exponent vectors of total degree <= degree: zero vector, unit vectors in identity order
    (e_0, e_1, ...: NOT lexicographic), then every higher degree as one ascending-lex block."""
    rows = [(0,) * dim]
    if degree >= 1:
        rows += [tuple(1 if j == i else 0 for j in range(dim)) for i in range(dim)]
    for g in range(2, degree + 1):
        rows += list(_compositions(g, dim))
    return np.array(rows, dtype=np.int32).reshape(-1, dim)


def n_terms(dim, degree):
    return math.comb(dim + degree, dim)


def hermite_table(degree):
    """Row k: ascending-power coefficients of He_k / sqrt(k!) (probabilists' Hermite).

    Integer coefficients by He_k = x He_{k-1} - (k-1) He_{k-2} (exact in float64 for the degrees
    the kernels accept), then ONE division of each entry by np.sqrt(float(k!)), which is the
    rounding the reference applies (:146-147)."""
    n = degree + 1
    He = [[0] * n for _ in range(n)]
    He[0][0] = 1
    if degree >= 1:
        He[1][1] = 1
    for k in range(2, n):
        for i in range(n):
            v = -(k - 1) * He[k - 2][i]
            if i >= 1:
                v += He[k - 1][i - 1]
            He[k][i] = v
    out = np.zeros((n, n))
    for k in range(n):
        out[k, :] = np.divide(np.array(He[k], dtype=float), np.sqrt(math.factorial(k)))
    return out


def fit_degree(no_snapshots, no_parameters, max_degree, current_degree):
    """min(floor(ln n / ln d), max_degree), never decreasing.  d == 1 divides by zero in the
    reference (log 1 = 0 -> inf -> int() raises OverflowError); reproduced as an error."""
    if no_parameters == 1:
        raise OverflowError("cannot convert float infinity to integer (polynomial degree rule with "
                            "no_parameters == 1, as in the reference)")
    deg = int(np.floor(np.log(no_snapshots) / np.log(no_parameters)))
    return max(min(deg, max_degree), current_degree)
