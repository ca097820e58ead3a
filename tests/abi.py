"""
The `forward` C-ABI as ctypes sees it: argument types and the flattening of a
problem dict (tests/problems.py, workloads.py) into the argument tuple.

Argument order = simwave's Middleware._keys_in_order
(simwave/kernel/backend/middleware.py:166-202; include/simwave_cuda.h).  The
CUDA backend's callers (tests/cuda_abi.py, bench.py, smoke) use this module;
the CPU checkers under oracle/ keep their own statement of the same ABI, so
nothing on the product arm imports from oracle/.
"""
import ctypes

import numpy as np
from numpy.ctypeslib import ndpointer


def forward_argtypes(ndim, density, dtype):
    ct = ctypes.c_float if np.dtype(dtype) == np.float32 else ctypes.c_double
    fp = ndpointer(ct, flags="C_CONTIGUOUS")
    up = ndpointer(ctypes.c_size_t, flags="C_CONTIGUOUS")
    sz = ctypes.c_size_t
    a = [fp, fp]                       # u, velocity
    if density:
        a.append(fp)                   # density
    a += [fp, fp, sz, sz]              # damp, wavelet, wavelet_size, wavelet_count
    a += [fp, fp] if density else [fp]  # coeff_order2[, coeff_order1]
    a += [up]                          # boundary_conditions
    a += [up, sz, fp, sz, up]          # source tables
    a += [up, sz, fp, sz, up]          # receiver tables
    a += [fp, sz, sz]                  # receivers, num_sources, num_receivers
    a += [sz] * ndim                   # nz, nx[, ny]
    a += [ct] * ndim                   # dz, dx[, dy]
    a += [sz, ct, sz, sz, sz, sz]      # saving_stride, dt, begin, end, space_order, num_snapshots
    return a


def forward_args(p):
    """Problem dict -> positional arguments of `forward`."""
    f = p["velocity"].dtype.type
    density = p.get("density") is not None
    args = [p["u"], p["velocity"]]
    if density:
        args.append(p["density"])
    args += [p["damp"], p["wavelet"], p["wavelet"].shape[0],
             1 if p["wavelet"].ndim == 1 else p["wavelet"].shape[1]]
    args += [p["coeff2"], p["coeff1"]] if density else [p["coeff2"]]
    args += [p["bc"]]
    args += [p["src_intervals"], len(p["src_intervals"]), p["src_values"],
             len(p["src_values"]), p["src_offsets"]]
    args += [p["rec_intervals"], len(p["rec_intervals"]), p["rec_values"],
             len(p["rec_values"]), p["rec_offsets"]]
    args += [p["receivers"], len(p["src_offsets"]) - 1, len(p["rec_offsets"]) - 1]
    args += list(p["velocity"].shape)
    args += [f(h) for h in p["spacing"]]
    args += [p["saving_stride"], f(p["dt"]), p.get("begin_timestep", 1),
             p["end_timestep"], p["space_order"], p["u"].shape[0]]
    return args
