"""
Finite-difference weights and the CFL time step.

Host-side mirror of simwave/kernel/frontend/fd.py.  The reference obtains the
central weights from the third-party ``findiff`` package
(``findiff.coefficients(deriv, acc)['center']['coefficients']``, fd.py:50,
requirements.txt:6 ``findiff>=0.8.9``), which is not vendored in the reference
tree and not installed here.  ``findiff`` solves the Taylor (Vandermonde)
system  sum_j c_j * j**i = i! * delta(i, deriv)  over the offsets
j = -p..p with p = (2*floor((deriv+1)/2) - 1 + acc) // 2.  We solve the same
system exactly in rational arithmetic and round once to float64, so the
float32 weights are the correctly rounded ones (pinned for orders 2/4/8 by the
reference's tests/test_space_model.py:133-155).
"""
from fractions import Fraction
from functools import lru_cache
from math import factorial

import numpy as np


@lru_cache(maxsize=None)
def _central_weights(derivative_order, space_order):
    """Exact rational weights on offsets -p..p."""
    width = 2 * ((derivative_order + 1) // 2) - 1 + space_order
    p = width // 2
    offsets = list(range(-p, p + 1))
    m = len(offsets)

    # augmented Vandermonde system, Gauss-Jordan over the rationals
    rows = []
    for i in range(m):
        rhs = Fraction(factorial(derivative_order)) if i == derivative_order \
            else Fraction(0)
        rows.append([Fraction(o) ** i for o in offsets] + [rhs])

    for col in range(m):
        piv = next(r for r in range(col, m) if rows[r][col] != 0)
        rows[col], rows[piv] = rows[piv], rows[col]
        inv = 1 / rows[col][col]
        rows[col] = [v * inv for v in rows[col]]
        for r in range(m):
            if r != col and rows[r][col] != 0:
                f = rows[r][col]
                rows[r] = [a - f * b for a, b in zip(rows[r], rows[col])]

    return tuple(rows[i][m] for i in range(m))


def coefficients(derivative_order, space_order):
    """
    Full symmetric list of central FD weights (float64 ndarray) for the given
    derivative at accuracy ``space_order`` (reference fd.py:31-52).
    """
    w = _central_weights(int(derivative_order), int(space_order))
    return np.array([float(v) for v in w], dtype=np.float64)


def half_coefficients(derivative_order, space_order):
    """Centre weight followed by the right-hand half (reference fd.py:5-28)."""
    full = coefficients(derivative_order, space_order)
    return full[len(full) // 2:]


def calculate_dt(dimension, space_order, grid_spacing, velocity_model):
    """
    CFL-limited time step for the 2nd-order-in-time acoustic scheme
    (reference fd.py:55-95):  dt = sqrt(4 / (ndim * sum|c|)) * min(h) / max(v).
    The operand types follow the reference so the value rounds identically.
    """
    a1 = 4
    a2 = dimension * np.sum(np.abs(coefficients(2, space_order)))
    limit = np.sqrt(a1 / a2)
    return limit * np.min(grid_spacing) / np.max(velocity_model)
