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


def batch_axis_windows(num_points, positions, half_width):
    """
    ``axis_window`` for many positions on one axis at once.

    Every position gets the same 2*half_width + 4 candidate indices around it
    (a superset of what ``axis_window`` evaluates); the window function is
    applied to the whole (count, candidates) matrix with the same element-wise
    operations on the same dtypes, so each row holds the same float32 weights
    as the one-by-one path.  Candidates outside the axis are discarded.

    Returns ``begin``, ``end`` (int64 arrays, inclusive) and a list-free pair
    ``(values, counts)``: the float32 weights of all windows concatenated in
    position order, and the number of weights per window.
    """
    positions = np.asarray(positions)
    count = positions.shape[0]
    first = np.floor(positions.astype(np.float64)).astype(np.int64) - half_width - 1
    candidates = first[:, None] + np.arange(2 * half_width + 4, dtype=np.int64)[None, :]
    inside = (candidates >= 0) & (candidates <= num_points - 1)
    # float32 indices are exact integers, as in the reference's linspace
    index = candidates.astype(np.float32)
    kws_values = _windowed_sinc(index, positions[:, None], half_width)
    valid = inside & ~np.isnan(kws_values)
    counts = valid.sum(axis=1)
    if count and counts.min() == 0:
        raise Exception(
            "There is no valid point in the source/receiver location"
        )
    # valid candidates of a row are contiguous: the window is an interval
    begin = np.where(valid, candidates, np.iinfo(np.int64).max).min(axis=1)
    end = np.where(valid, candidates, np.iinfo(np.int64).min).max(axis=1)
    return begin, end, kws_values[valid].astype(np.float32), counts


def get_source_points_batch(grid_shape, locations, half_width):
    """
    ``get_source_points`` for an array of locations, shape (count, ndim):
    the three tables the kernel consumes -- uint64 intervals
    ``[b_axis1, e_axis1, ..]`` per location, the concatenated float32 weights
    ``[axis1.., axis2.., ..]`` per location, and the uint64 running offsets of
    each location's weights (length count + 1).  Bit-identical to calling
    ``get_source_points`` location by location.
    """
    locations = np.asarray(locations)
    if locations.ndim != 2 or locations.shape[1] != len(grid_shape):
        raise Exception(
            "Grid and source/receiver location must have the same dimension."
        )
    count, ndim = locations.shape
    intervals = np.empty((count, 2 * ndim), dtype=np.uint)
    axis_values, axis_counts = [], []
    for axis, num_points in enumerate(grid_shape):
        begin, end, values, counts = batch_axis_windows(
            num_points, locations[:, axis], half_width)
        intervals[:, 2 * axis] = begin
        intervals[:, 2 * axis + 1] = end
        axis_values.append(values)
        axis_counts.append(counts)

    per_location = np.sum(axis_counts, axis=0) if count else np.zeros(0, np.int64)
    offsets = np.zeros(count + 1, dtype=np.uint)
    offsets[1:] = np.cumsum(per_location)
    weights = np.empty(int(offsets[-1]), dtype=np.float32)
    # scatter every axis' weights behind the previous axes' of the same location
    start = offsets[:-1].astype(np.int64)
    for values, counts in zip(axis_values, axis_counts):
        owner = np.repeat(np.arange(count), counts)
        within = np.arange(values.size) - np.repeat(np.cumsum(counts) - counts, counts)
        weights[start[owner] + within] = values
        start = start + counts
    return intervals.reshape(-1), weights, offsets


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
