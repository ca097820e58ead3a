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


# ---------------------------------------------------------------------------
# Adjoint operator (test infrastructure, like everything in this module).
#
# The reference has no adjoint kernel to compile, so this checker restates the
# operator the product defines (include/simwave_cuda.h section 1b) with the
# reference's own forward kernel as its engine: g = F^T d equals the forward
# run with the tables exchanged -- the traces injected at the receiver windows,
# the field sampled at the source windows -- and time reversed, with the
# per-axis weights of a point on a Neumann face plane doubled on the injection
# side and halved on the sampling side, windows clipped to interior points.
# Pinned by the one property that defines an adjoint: <F w, d> = <w, F^T d>
# for random w, d (tests/test_adjoint.py, float64, <= 1e-10), with F the
# compiled reference.
# ---------------------------------------------------------------------------
def adjoint_tables(p, intervals, values, offsets, inject):
    ndim = p["velocity"].ndim
    shape = p["velocity"].shape
    r = p["space_order"] // 2
    bc = [int(b) for b in p["bc"]]
    count = len(offsets) - 1
    iv = np.asarray(intervals).reshape(count, 2 * ndim).copy()
    out_values, out_offsets = [], [0]
    for i in range(count):
        v = values[int(offsets[i]):int(offsets[i + 1])]
        pos, kept = 0, 0
        for ax in range(ndim):
            b, e = int(iv[i, 2 * ax]), int(iv[i, 2 * ax + 1])
            w = v[pos:pos + e - b + 1]
            pos += e - b + 1
            lo, hi = r, shape[ax] - r - 1
            cb, ce = max(b, lo), min(e, hi)
            if cb > ce:
                cb = ce = min(max(b, lo), hi)
                w = np.zeros(1, dtype=values.dtype)
            else:
                w = w[cb - b:ce - b + 1].copy()
                for face, plane in ((2 * ax, lo), (2 * ax + 1, hi)):
                    if bc[face] == 2 and cb <= plane <= ce:
                        w[plane - cb] *= values.dtype.type(2.0 if inject else 0.5)
            iv[i, 2 * ax], iv[i, 2 * ax + 1] = cb, ce
            out_values.append(w)
            kept += w.size
        out_offsets.append(out_offsets[-1] + kept)
    return (np.ascontiguousarray(iv.reshape(-1).astype(np.uint64)),
            np.ascontiguousarray(np.concatenate(out_values).astype(values.dtype)),
            np.asarray(out_offsets, dtype=np.uint64))


def adjoint(p, kind=None, variant=""):
    """g = F^T d for problem dict ``p``: reads p['receivers'] (the data d),
    writes p['wavelet'] (rows begin-1 .. end-1) and p['u'] in place, like the
    product's `adjoint` entry points."""
    if p.get("density") is not None:
        raise ValueError("adjoint: constant density only")
    begin, end = p.get("begin_timestep", 1), p["end_timestep"]
    steps = end - begin + 1
    nsrc, nrec = len(p["src_offsets"]) - 1, len(p["rec_offsets"]) - 1
    q = dict(p)
    q["src_intervals"], q["src_values"], q["src_offsets"] = adjoint_tables(
        p, p["rec_intervals"], p["rec_values"], p["rec_offsets"], True)
    q["rec_intervals"], q["rec_values"], q["rec_offsets"] = adjoint_tables(
        p, p["src_intervals"], p["src_values"], p["src_offsets"], False)
    reversed_traces = np.ascontiguousarray(p["receivers"][begin - 1:end][::-1])
    q["wavelet"] = reversed_traces if nrec > 1 else reversed_traces.reshape(steps)
    q["receivers"] = np.zeros((steps, nsrc), dtype=p["receivers"].dtype)
    q["begin_timestep"], q["end_timestep"] = 1, steps
    seconds = forward(q, kind=kind, variant=variant)
    g = q["receivers"][::-1]
    if p["wavelet"].ndim == 1:
        p["wavelet"][begin - 1:end] = g.sum(axis=1) if nsrc > 1 else g[:, 0]
    else:
        p["wavelet"][begin - 1:end] = g
    return seconds
