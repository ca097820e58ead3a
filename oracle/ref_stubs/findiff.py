"""
Stand-in for the third-party `findiff` package (not installed, no network) so
that the reference front end can be imported in this container to generate
golden vectors (tests/golden/make_golden.py).  TEST INFRASTRUCTURE ONLY.

findiff.coefficients(deriv, acc) returns, for the central scheme, the solution
of the Taylor system on offsets -p..p, p = (2*floor((deriv+1)/2) - 1 + acc)//2.
findiff itself solves it numerically in float64; this stand-in solves it
exactly with sympy rationals (an implementation independent of
simwave_b200.kernel.frontend.fd, which uses fractions.Fraction).
"""
import numpy as np
import sympy


def coefficients(deriv, acc):
    num_central = 2 * ((deriv + 1) // 2) - 1 + acc
    p = num_central // 2
    offsets = list(range(-p, p + 1))
    n = len(offsets)
    A = sympy.Matrix(n, n, lambda i, j: sympy.Integer(offsets[j]) ** i)
    b = sympy.zeros(n, 1)
    b[deriv] = sympy.factorial(deriv)
    sol = A.LUsolve(b)
    return {
        'center': {
            'coefficients': np.array([float(v) for v in sol], dtype=np.float64),
            'offsets': np.array(offsets),
        }
    }
