"""
Slab domain decomposition of one 3D forward run across GPUs (one process per
GPU).  New functionality: the reference has no multi-device code at all
(SURVEY.md section 2.2); the kernel ABI it defines is kept, each rank simply
runs the same time loop on its z-slab.

Decomposition: the interior planes r .. nz-r-1 of the extended grid are split
into contiguous z-ranges, one per rank.  A rank's local arrays are the planes
it owns plus r planes on either side: the grid's own halo at the outer faces,
GHOST planes (owned by the neighbour) at the inner faces.  The device keeps the
ghost planes of the wavefield current (include/simwave_cuda.h, "Slab
decomposition"): after every step the outermost owned planes are written into
the neighbour's ghost planes through a CUDA IPC peer mapping and a per-step
flag is published; no host synchronisation inside the time loop.

Sources and receivers: every rank keeps every source / receiver (so that
per-source wavelets and trace columns keep their indices) with its window
clipped to the planes the rank owns; a window that lies entirely elsewhere
becomes a single zero-weight point.  Receiver traces are therefore partial
sums, added over ranks at the end (exact when a window lies inside one slab).

This module holds the host-side logic only (pure NumPy) plus a thin ctypes
wrapper over the plan API; `torch.distributed` (NCCL or gloo) is used by the
callers for plumbing: exchanging the 512-byte descriptors, barriers, and the
final reduction of the traces.
"""
import ctypes
import os

import numpy as np

from simwave_b200.kernel.backend.compiler import LIB_DIR

DESC_BYTES = 512


def split_planes(nz, radius, world):
    """Owned interior plane ranges [(lo, hi), ...] (global indices, hi
    exclusive), as even as possible.  Every slab owns at least 2*radius planes:
    the planes it hands to its upper and to its lower neighbour must not
    overlap."""
    interior = nz - 2 * radius
    if world < 1 or interior // world < 2 * radius:
        raise ValueError("too many slabs for %d interior planes" % interior)
    base, extra = divmod(interior, world)
    ranges, lo = [], radius
    for k in range(world):
        hi = lo + base + (1 if k < extra else 0)
        ranges.append((lo, hi))
        lo = hi
    return ranges


def _clip_tables(intervals, values, offsets, ndim, z_lo, z_hi, z_shift):
    """Clip every window's z-interval to [z_lo, z_hi) (global planes), shift
    to local plane indices; windows outside become one zero-weight point."""
    count = len(offsets) - 1
    iv = intervals.reshape(count, 2 * ndim).copy()
    new_vals, new_off = [], [0]
    for i in range(count):
        v = values[int(offsets[i]):int(offsets[i + 1])]
        zb, ze = int(iv[i, 0]), int(iv[i, 1])
        nzw = ze - zb + 1
        wz, rest = v[:nzw], v[nzw:]
        cb, ce = max(zb, z_lo), min(ze, z_hi - 1)
        if cb > ce:
            cb = ce = min(max(zb, z_lo), z_hi - 1)
            wz_new = np.zeros(1, dtype=values.dtype)
        else:
            wz_new = wz[cb - zb:ce - zb + 1]
        iv[i, 0], iv[i, 1] = cb - z_shift, ce - z_shift
        new_vals.append(np.concatenate([wz_new, rest]))
        new_off.append(new_off[-1] + wz_new.size + rest.size)
    return (np.ascontiguousarray(iv.reshape(-1)),
            np.ascontiguousarray(np.concatenate(new_vals).astype(values.dtype)),
            np.asarray(new_off, dtype=np.uint64))


def partition(p, rank, world):
    """Local problem of ``rank`` for the global ABI-level problem dict ``p``
    (layout of tests/problems.py / workloads.py).  Returns (local, info)."""
    if p["velocity"].ndim != 3:
        raise ValueError("slab decomposition is for 3D problems")
    if p["saving_stride"] != 0:
        raise ValueError("slab decomposition needs saving_stride == 0")
    r = p["space_order"] // 2
    nz = p["velocity"].shape[0]
    lo, hi = split_planes(nz, r, world)[rank]
    a, b = lo - r, hi + r                      # local planes [a, b)
    up, down = rank > 0, rank < world - 1

    q = dict(p)
    for key in ("velocity", "density", "damp"):
        if p.get(key) is not None:
            q[key] = np.ascontiguousarray(p[key][a:b])
    q["u"] = np.ascontiguousarray(p["u"][:, a:b])
    q["receivers"] = np.zeros_like(p["receivers"])
    bc = p["bc"].copy()
    if up:
        bc[0] = 0
    if down:
        bc[1] = 0
    q["bc"] = bc
    # planes this rank answers for: its interior range, plus the grid's own
    # halo at an outer face
    own_lo = lo if up else 0
    own_hi = hi if down else nz
    for kind in ("src", "rec"):
        iv, val, off = _clip_tables(p[kind + "_intervals"], p[kind + "_values"],
                                    p[kind + "_offsets"], 3, own_lo, own_hi, a)
        q[kind + "_intervals"], q[kind + "_values"], q[kind + "_offsets"] = iv, val, off
    q["slab_up"], q["slab_down"] = int(up), int(down)
    info = {"planes": (a, b), "interior": (lo, hi), "owned": (own_lo, own_hi),
            "up": up, "down": down, "radius": r}
    return q, info


def assemble_wavefield(parts, infos, nz):
    """Global u (slots, nz, nx, ny) from the per-rank local arrays."""
    first = parts[0]
    out = np.zeros((first.shape[0], nz) + first.shape[2:], dtype=first.dtype)
    for u, info in zip(parts, infos):
        a, _ = info["planes"]
        lo, hi = info["owned"]
        out[:, lo:hi] = u[:, lo - a:hi - a]
    return out


# ---------------------------------------------------------------------------
# ctypes wrapper over the plan API
# ---------------------------------------------------------------------------
class _Problem(ctypes.Structure):
    _fields_ = [
        ("ndim", ctypes.c_int), ("dtype_bytes", ctypes.c_int),
        ("u", ctypes.c_void_p), ("velocity", ctypes.c_void_p),
        ("density", ctypes.c_void_p), ("damp", ctypes.c_void_p),
        ("wavelet", ctypes.c_void_p), ("wavelet_size", ctypes.c_size_t),
        ("wavelet_count", ctypes.c_size_t),
        ("coeff_order2", ctypes.c_void_p), ("coeff_order1", ctypes.c_void_p),
        ("boundary_conditions", ctypes.c_void_p),
        ("src_points_interval", ctypes.c_void_p),
        ("src_points_values", ctypes.c_void_p),
        ("src_points_values_size", ctypes.c_size_t),
        ("src_points_values_offset", ctypes.c_void_p),
        ("rec_points_interval", ctypes.c_void_p),
        ("rec_points_values", ctypes.c_void_p),
        ("rec_points_values_size", ctypes.c_size_t),
        ("rec_points_values_offset", ctypes.c_void_p),
        ("receivers", ctypes.c_void_p),
        ("num_sources", ctypes.c_size_t), ("num_receivers", ctypes.c_size_t),
        ("nz", ctypes.c_size_t), ("nx", ctypes.c_size_t), ("ny", ctypes.c_size_t),
        ("dz", ctypes.c_double), ("dx", ctypes.c_double), ("dy", ctypes.c_double),
        ("saving_stride", ctypes.c_size_t), ("dt", ctypes.c_double),
        ("space_order", ctypes.c_size_t), ("num_snapshots", ctypes.c_size_t),
        ("slab_up", ctypes.c_int), ("slab_down", ctypes.c_int),
        ("u_slot_stride", ctypes.c_size_t),
        ("out_plane_begin", ctypes.c_size_t), ("out_plane_end", ctypes.c_size_t),
    ]


def problem_struct(p, keep):
    """``simwave_problem`` (include/simwave_cuda.h) for problem dict ``p``;
    arrays referenced by the struct are appended to ``keep``."""
    def ptr(a):
        if a is None:
            return None
        keep.append(a)
        return a.ctypes.data

    shape = p["velocity"].shape
    ndim = len(shape)
    f = p["velocity"].dtype.type
    h = [float(f(x)) for x in p["spacing"]]
    pb = _Problem()
    pb.ndim = ndim
    pb.dtype_bytes = p["velocity"].dtype.itemsize
    pb.u = ptr(p["u"])
    pb.velocity = ptr(p["velocity"])
    pb.density = ptr(p.get("density"))
    pb.damp = ptr(p["damp"])
    pb.wavelet = ptr(p["wavelet"])
    pb.wavelet_size = p["wavelet"].shape[0]
    pb.wavelet_count = 1 if p["wavelet"].ndim == 1 else p["wavelet"].shape[1]
    pb.coeff_order2 = ptr(p["coeff2"])
    pb.coeff_order1 = ptr(p["coeff1"]) if p.get("density") is not None else None
    pb.boundary_conditions = ptr(p["bc"])
    pb.src_points_interval = ptr(p["src_intervals"])
    pb.src_points_values = ptr(p["src_values"])
    pb.src_points_values_size = len(p["src_values"])
    pb.src_points_values_offset = ptr(p["src_offsets"])
    pb.rec_points_interval = ptr(p["rec_intervals"])
    pb.rec_points_values = ptr(p["rec_values"])
    pb.rec_points_values_size = len(p["rec_values"])
    pb.rec_points_values_offset = ptr(p["rec_offsets"])
    pb.receivers = ptr(p["receivers"])
    pb.num_sources = len(p["src_offsets"]) - 1
    pb.num_receivers = len(p["rec_offsets"]) - 1
    pb.nz, pb.nx = shape[0], shape[1]
    pb.ny = shape[2] if ndim == 3 else 0
    pb.dz, pb.dx = h[0], h[1]
    pb.dy = h[2] if ndim == 3 else 0.0
    pb.saving_stride = p["saving_stride"]
    pb.dt = float(f(p["dt"]))
    pb.space_order = p["space_order"]
    pb.num_snapshots = p["u"].shape[0]
    pb.slab_up = int(p.get("slab_up", 0))
    pb.slab_down = int(p.get("slab_down", 0))
    return pb


_core = None


def core_library():
    """libsimwave_b200.so with the plan API prototyped."""
    global _core
    if _core is None:
        path = os.path.join(LIB_DIR, "libsimwave_b200.so")
        if not os.path.exists(path):
            raise FileNotFoundError(
                path + " is missing; build it with __graft_entry__.build(). "
                "There is no CPU fallback.")
        lib = ctypes.CDLL(path)
        lib.simwave_cuda_last_error.restype = ctypes.c_char_p
        lib.simwave_plan_create.restype = ctypes.c_void_p
        lib.simwave_plan_create.argtypes = [ctypes.c_void_p]
        lib.simwave_plan_run.argtypes = [ctypes.c_void_p, ctypes.c_size_t,
                                         ctypes.c_size_t,
                                         ctypes.POINTER(ctypes.c_double)]
        lib.simwave_plan_reset.argtypes = [ctypes.c_void_p]
        lib.simwave_plan_download.argtypes = [ctypes.c_void_p, ctypes.c_void_p,
                                              ctypes.c_void_p]
        lib.simwave_plan_destroy.argtypes = [ctypes.c_void_p]
        lib.simwave_plan_slab_export.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        lib.simwave_plan_slab_connect.argtypes = [ctypes.c_void_p, ctypes.c_void_p,
                                                  ctypes.c_void_p]
        lib.simwave_cuda_last_launch_count.restype = ctypes.c_ulonglong
        _core = lib
    return _core


class Plan:
    """A problem resident on the current CUDA device (plan API)."""

    def __init__(self, p):
        self.lib = core_library()
        self._keep = []
        self.problem = p
        pb = problem_struct(p, self._keep)
        self.handle = self.lib.simwave_plan_create(ctypes.byref(pb))
        if not self.handle:
            raise RuntimeError("plan_create failed: " + self._error())

    def _error(self):
        return self.lib.simwave_cuda_last_error().decode(errors="replace")

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError("%s failed: %s" % (what, self._error()))

    def run(self, begin=None, end=None):
        """Advance the time loop; returns the device time of the loop (s)."""
        begin = 1 if begin is None else begin
        end = self.problem["end_timestep"] if end is None else end
        loop = ctypes.c_double()
        self._check(self.lib.simwave_plan_run(self.handle, begin, end,
                                              ctypes.byref(loop)), "plan_run")
        return loop.value

    def reset(self):
        self._check(self.lib.simwave_plan_reset(self.handle), "plan_reset")

    def download(self):
        """Write wavefield slots and traces into the problem's own arrays."""
        self._check(self.lib.simwave_plan_download(self.handle, None, None),
                    "plan_download")

    def launches(self):
        return int(self.lib.simwave_cuda_last_launch_count())

    def slab_export(self):
        buf = ctypes.create_string_buffer(DESC_BYTES)
        self._check(self.lib.simwave_plan_slab_export(self.handle, buf),
                    "slab_export")
        return buf.raw

    def slab_connect(self, up_desc, down_desc):
        up = ctypes.create_string_buffer(up_desc, DESC_BYTES) if up_desc else None
        down = ctypes.create_string_buffer(down_desc, DESC_BYTES) if down_desc else None
        self._check(self.lib.simwave_plan_slab_connect(self.handle, up, down),
                    "slab_connect")

    def destroy(self):
        if self.handle:
            self.lib.simwave_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def connect_neighbours(plan, rank, world, all_gather_bytes):
    """Exchange slab descriptors and connect ``plan`` to its neighbours.
    ``all_gather_bytes(b)`` returns the list of every rank's ``b`` (e.g. built
    on torch.distributed.all_gather_object)."""
    descs = all_gather_bytes(plan.slab_export())
    plan.slab_connect(descs[rank - 1] if rank > 0 else None,
                      descs[rank + 1] if rank < world - 1 else None)
