"""
Kaiser-windowed-sinc interpolation tables (Hicks 2002) for sources/receivers.

Host-side mirror of simwave/kernel/frontend/kws.py.  The reference evaluates
the window over a whole grid axis and then walks it in a Python loop to find
the non-NaN entries (kws.py:44-135).  Here only the handful of candidate
indices around the position are evaluated, with the *same* NumPy/SciPy
element-wise operations on the same dtypes, so indices and float32 weights
come out bit-identical to the reference run in the same environment
(checked against tests/golden/tables_*.npz).
"""
import warnings

import numpy as np
from scipy.special import i0

# Optimal Kaiser b per window half-width (reference kws.py:28-39, Hicks 2002).
_KAISER_B = {1: 1.24, 2: 2.94, 3: 4.53, 4: 6.31, 5: 7.91,
             6: 9.42, 7: 10.95, 8: 12.53, 9: 14.09, 10: 14.18}


def get_kaiser_half_width(half_width):
    """b parameter of the Kaiser window for ``half_width`` (1..10)."""
    if half_width not in _KAISER_B:
        raise Exception(
            "Kaiser windowing half-width {} not supported".format(half_width)
        )
    return _KAISER_B[half_width]


def _windowed_sinc(index, source_point, half_width):
    """
    Kaiser-windowed sinc at float32 grid indices ``index`` for a source at
    ``source_point`` (grid units).  NaN outside the window.  Operation order
    and dtypes follow reference kws.py:71-94.
    """
    x = index - source_point
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        root = np.sqrt(1 - (x / half_width) ** 2)
    b = get_kaiser_half_width(half_width)
    kaiser = i0(b * root) / i0(b)
    return kaiser * np.sinc(x)


def kaiser_windowing_sinc(num_points, source_point, half_width):
    """
    Full-axis version kept for API compatibility (reference kws.py:44-94):
    1D array over the whole axis, NaN outside the window.
    """
    index = np.linspace(start=0, stop=num_points - 1, num=num_points,
                        dtype=np.float32)
    return _windowed_sinc(index, source_point, half_width)


def get_kws_valid_points(kaiser_windowed_array):
    """First/last non-NaN index and the float32 values between them
    (reference kws.py:97-135)."""
    arr = np.asarray(kaiser_windowed_array)
    valid = np.flatnonzero(~np.isnan(arr))
    if valid.size == 0:
        raise Exception(
            "There is no valid point in the source/receiver location"
        )
    return int(valid[0]), int(valid[-1]), arr[valid].astype(np.float32)


def axis_window(num_points, source_point, half_width):
    """
    Window of one axis: (begin, end, float32 values).  Equivalent to
    ``get_kws_valid_points(kaiser_windowing_sinc(...))`` but evaluates only
    the indices that can be inside the window.
    """
    centre = float(source_point)
    lo = max(int(np.floor(centre)) - half_width - 1, 0)
    hi = min(int(np.ceil(centre)) + half_width + 1, num_points - 1)
    if hi < lo:
        raise Exception(
            "There is no valid point in the source/receiver location"
        )
    # float32 indices are exact integers, as in the reference's linspace
    index = np.arange(lo, hi + 1, dtype=np.float32)
    begin, end, values = get_kws_valid_points(
        _windowed_sinc(index, source_point, half_width)
    )
    return begin + lo, end + lo, values


def get_source_points(grid_shape, source_location, half_width):
    """
    Point interval and weights of one source/receiver over all axes
    (reference kws.py:138-183).

    Returns ``[b_axis1, e_axis1, ..]`` (uint64) and the concatenated float32
    weights ``[axis1.., axis2.., ..]``.
    """
    if len(grid_shape) != len(source_location):
        raise Exception(
            "Grid and source/receiver location must have the same dimension."
        )

    bounds = []
    weights = []
    for num_points, position in zip(grid_shape, source_location):
        begin, end, values = axis_window(num_points, position, half_width)
        bounds += [begin, end]
        weights.append(values)

    return (np.array(bounds, dtype=np.uint),
            np.concatenate(weights).astype(np.float32, copy=False))
