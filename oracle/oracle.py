"""
ctypes loader for the CPU checkers.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product (simwave_b200/) never does.

Two kinds of library, both exporting the reference ABI ``double forward(...)``
(simwave/kernel/backend/middleware.py:149-154, argument order of
``Middleware._keys_in_order`` :166-202):

* ``kind='ref'``  -- the reference's own wave.c, compiled by oracle/Makefile
  from /root/reference into oracle/_ref/ (prebuilt files travel to the GPU box).
* ``kind='port'`` -- oracle/wave_oracle.c, our restatement, compiled on demand
  into oracle/_build/ (needs only gcc).
"""
import ctypes
import os
import subprocess

import numpy as np
from numpy.ctypeslib import ndpointer

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("SIMWAVE_REFERENCE", "/root/reference")

_DTYPE_TAG = {"float32": "f32", "float64": "f64"}


def lib_path(kind, ndim, density, dtype, variant=""):
    dens = "variable" if density else "constant"
    tag = _DTYPE_TAG[np.dtype(dtype).name]
    sub = {"ref": "_ref", "port": "_build"}[kind]
    name = "{}_{}d_{}_{}{}.so".format(kind, ndim, dens, tag,
                                      "_" + variant if variant else "")
    return os.path.join(HERE, sub, name)


def ensure_built(kind):
    """Build the requested family with make (no-op when up to date)."""
    if kind == "ref" and not os.path.isdir(REF_ROOT):
        return  # GPU box: only prebuilt files are available
    subprocess.run(["make", "-s", "-C", HERE, kind, "REF=" + REF_ROOT],
                   check=True)


def available(kind, ndim=3, density=False, dtype=np.float32, variant=""):
    return os.path.exists(lib_path(kind, ndim, density, dtype, variant))


def best_kind():
    """'ref' when the compiled reference is present, else 'port'."""
    return "ref" if available("ref") else "port"


def _argtypes(ndim, density, dtype):
    ct = ctypes.c_float if np.dtype(dtype) == np.float32 else ctypes.c_double
    fp = ndpointer(ct, flags="C_CONTIGUOUS")
    up = ndpointer(ctypes.c_size_t, flags="C_CONTIGUOUS")
    sz = ctypes.c_size_t
    a = [fp, fp]                       # u, velocity
    if density:
        a.append(fp)                   # density
    a += [fp, fp, sz, sz]              # damp, wavelet, wavelet_size, wavelet_count
    a += [fp, fp] if density else [fp]  # coeff(s)
    a += [up]                          # boundary_conditions
    a += [up, sz, fp, sz, up]          # src tables
    a += [up, sz, fp, sz, up]          # rec tables
    a += [fp, sz, sz]                  # receivers, num_sources, num_receivers
    a += [sz] * ndim                   # nz, nx[, ny]
    a += [ct] * ndim                   # dz, dx[, dy]
    a += [sz, ct, sz, sz, sz, sz]      # saving_stride, dt, begin, end, space_order, num_snapshots
    return a


_cache = {}


def load(kind, ndim, density, dtype, variant=""):
    key = (kind, ndim, bool(density), np.dtype(dtype).name, variant)
    if key not in _cache:
        path = lib_path(kind, ndim, density, dtype, variant)
        if not os.path.exists(path):
            ensure_built(kind)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        lib = ctypes.CDLL(path)
        fn = lib.forward
        fn.restype = ctypes.c_double
        fn.argtypes = _argtypes(ndim, density, dtype)
        _cache[key] = fn
    return _cache[key]


def abi_args(p):
    """Flatten a problem dict (see tests/problems.py) into the ABI tuple."""
    ndim = p["velocity"].ndim
    dtype = p["velocity"].dtype
    f = dtype.type
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
    assert ndim in (2, 3)
    return args


def forward(p, kind=None, variant=""):
    """Run the CPU checker in place on problem dict ``p`` (mutates p['u'] and
    p['receivers']); returns the kernel's own wall-clock seconds."""
    kind = kind or best_kind()
    fn = load(kind, p["velocity"].ndim, p.get("density") is not None,
              p["velocity"].dtype, variant)
    return fn(*abi_args(p))
