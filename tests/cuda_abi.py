"""
Calls the CUDA backend the way simwave's Middleware does: ctypes on the
drop-in shim library's `forward`, argtypes built from the values
(simwave/kernel/backend/middleware.py:107-158).  Used by the GPU parity tests,
smoke() and bench.py so that everything goes through the C ABI.
"""
import ctypes
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (REPO, os.path.dirname(os.path.abspath(__file__))):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from simwave_b200.kernel.backend.compiler import prebuilt_library  # noqa: E402
from abi import forward_args, forward_argtypes  # noqa: E402

_libs = {}


def shim(ndim, density, dtype):
    key = (ndim, bool(density), np.dtype(dtype).name)
    if key not in _libs:
        path = prebuilt_library(
            ndim, "variable_density" if density else "constant_density",
            "-DFLOAT" if np.dtype(dtype) == np.float32 else "-DDOUBLE")
        lib = ctypes.CDLL(path)
        for fn in (lib.forward, lib.adjoint):       # same argument list
            fn.restype = ctypes.c_double
            fn.argtypes = forward_argtypes(ndim, density, dtype)
        _libs[key] = lib
    return _libs[key]


def core():
    path = os.path.join(os.path.dirname(prebuilt_library(
        3, "constant_density", "-DFLOAT")), "libsimwave_b200.so")
    lib = ctypes.CDLL(path)
    lib.simwave_cuda_last_error.restype = ctypes.c_char_p
    lib.simwave_cuda_version.restype = ctypes.c_char_p
    lib.simwave_cuda_last_launch_count.restype = ctypes.c_ulonglong
    lib.simwave_cuda_last_timing.argtypes = [ctypes.POINTER(ctypes.c_double)] * 4
    return lib


def last_timing():
    vals = (ctypes.c_double * 6)()
    core().simwave_cuda_last_timing_ex(vals, 6)
    return dict(zip(("loop", "h2d", "d2h", "total", "run_wall", "teardown"),
                    [float(v) for v in vals]))


def cuda_forward(p):
    """Run problem dict ``p`` in place through the shim's `forward`."""
    lib = shim(p["velocity"].ndim, p.get("density") is not None,
               p["velocity"].dtype)
    seconds = lib.forward(*forward_args(p))
    if seconds < 0:
        raise RuntimeError(core().simwave_cuda_last_error().decode())
    return seconds


def cuda_adjoint(p):
    """g = F^T d through the shim's `adjoint`: reads p['receivers'], writes
    p['wavelet'] and p['u'] in place (include/simwave_cuda.h section 1b)."""
    lib = shim(p["velocity"].ndim, p.get("density") is not None,
               p["velocity"].dtype)
    seconds = lib.adjoint(*forward_args(p))
    if seconds < 0:
        raise RuntimeError(core().simwave_cuda_last_error().decode())
    return seconds
