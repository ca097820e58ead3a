"""
Synthetic versions of the five BASELINE.json configurations, built as the
arrays the `forward` C-ABI receives.  The big 3D models
are built directly in their extended (halo + damping layer) form: pushing a
10^8..10^9-point model through SpaceModel.interpolate needs tens of GB of
float64 temporaries, and the kernel only ever sees the
extended arrays anyway.  Geometry, boundary conditions, source / receiver
layout and wavelets follow the reference's benchmark scripts.
"""
import numpy as np

from simwave_b200.kernel.frontend import fd, kws

BC = {"none": 0, "null_dirichlet": 1, "null_neumann": 2}


def _tables(shape, positions, radius, dtype):
    iv, values, offsets = kws.get_source_points_batch(
        shape, np.asarray(positions, dtype=dtype), radius)
    return iv.astype(np.uint64), values.astype(dtype), offsets.astype(np.uint64)


def _damping(inner_shape, nbl, halo, alpha, degree, dtype):
    mask = np.pad(np.zeros(inner_shape, dtype=dtype), nbl, mode="linear_ramp",
                  end_values=nbl)
    mask = (mask ** degree) * alpha
    return np.pad(mask, [(halo, halo)] * len(inner_shape)).astype(dtype)


def _extend(model, nbl, halo):
    pad = [(b + halo, a + halo) for b, a in nbl]
    return np.ascontiguousarray(np.pad(model, pad, mode="edge"))


def _ricker(f0, t0, tf, timesteps, dtype):
    t = np.linspace(dtype(t0), dtype(tf), timesteps, dtype=dtype)
    arg = np.pi * f0 * (t - 1 / f0)
    return (1 * (1 - 2.0 * arg ** 2) * np.exp(-arg ** 2)).astype(dtype)


def _assemble(velocity, density, nbl, spacing, space_order, bc, src_coords,
              rec_coords, radius, f0, tf, timesteps, alpha=0.001, degree=3,
              dtype=np.float32, saving_stride=0, name=""):
    """velocity/density: physical-domain models; coords in metres."""
    dtype = np.dtype(dtype).type
    ndim = velocity.ndim
    r = space_order // 2
    inner = tuple(n + b + a for n, (b, a) in zip(velocity.shape, nbl))
    ext_v = _extend(velocity.astype(dtype), nbl, r)
    ext_d = None if density is None else _extend(density.astype(dtype), nbl, r)
    damp = _damping(velocity.shape, nbl, r, alpha, degree, dtype)
    shape = ext_v.shape
    h = [dtype(x) for x in spacing]
    dt = dtype(fd.calculate_dt(ndim, space_order, h, velocity.astype(dtype)))
    full_steps = int(np.ceil((dtype(tf) - dtype(0) + dt) / dt))
    if timesteps is None:
        timesteps = full_steps

    def to_grid(coords):
        origin = np.array([b + r for b, _ in nbl], dtype=dtype)
        return np.asarray(coords, dtype=dtype) / np.array(h, dtype=dtype) + origin

    src_iv, src_val, src_off = _tables(shape, to_grid(src_coords), radius, dtype)
    rec_iv, rec_val, rec_off = _tables(shape, to_grid(rec_coords), radius, dtype)
    wavelet = _ricker(f0, 0.0, tf, full_steps, dtype)[:timesteps].copy()
    slots = 3 if saving_stride == 0 else \
        len(range(0, timesteps, saving_stride)) + 2
    del inner
    return {
        "name": name,
        "u": np.zeros((slots,) + shape, dtype=dtype),
        "velocity": ext_v, "density": ext_d, "damp": damp,
        "wavelet": wavelet,
        "coeff2": dtype(fd.half_coefficients(2, space_order)),
        "coeff1": dtype(fd.half_coefficients(1, space_order)),
        "bc": np.asarray([BC[b] for b in bc], dtype=np.uint64),
        "src_intervals": src_iv, "src_values": src_val, "src_offsets": src_off,
        "rec_intervals": rec_iv, "rec_values": rec_val, "rec_offsets": rec_off,
        "receivers": np.zeros((timesteps, len(rec_off) - 1), dtype=dtype),
        "spacing": spacing, "saving_stride": saving_stride, "dt": dt,
        "end_timestep": timesteps, "space_order": space_order,
        "full_timesteps": full_steps,
    }


def interior_points(p):
    r = p["space_order"] // 2
    return int(np.prod([n - 2 * r for n in p["velocity"].shape]))


def bytes_per_point(p):
    """Algorithmic bytes per grid-point update:
    read u_prev, u_cur, velocity, damp (+density), write u_next."""
    fields = 5 + (1 if p.get("density") is not None else 0)
    return fields * p["velocity"].dtype.itemsize


# ---------------------------------------------------------------------------
def readme_2d_velocity():
    vel = np.full((513, 513), 1500.0, dtype=np.float32)
    vel[257:] = 2000.0
    return vel


def readme_2d(timesteps=None):
    """C1: README 2D two-layer model (reference README.md:54-137)."""
    vel = readme_2d_velocity()
    return _assemble(
        vel, None, ((0, 0), (0, 0)), (10.0, 10.0), 4,
        ("null_neumann", "null_dirichlet", "none", "null_dirichlet"),
        [(2560.0, 2560.0)], [(2560.0, 10.0 * i) for i in range(512)],
        1, 10.0, 1.0, timesteps, name="readme_2d")


def marmousi_2d_velocity(seed=1):
    nz, nx = 351, 1701
    rng = np.random.default_rng(seed)
    z = np.arange(nz, dtype=np.float64)[:, None]
    x = np.arange(nx, dtype=np.float64)[None, :]
    vel = 1500.0 + (4700.0 - 1500.0) * (z / (nz - 1)) + 0.0 * x
    for k in range(8):                       # dipping reflector steps
        depth = 60 + 35 * k + 0.02 * (k + 1) * x
        vel = vel + 120.0 * (z > depth)
    coarse = rng.standard_normal((12, 40))
    smooth = np.kron(coarse, np.ones((30, 43)))[:nz, :nx]
    vel = vel * (1.0 + 0.03 * np.tanh(smooth))
    vel[:20] = 1500.0                        # water layer
    return np.clip(vel, 1028.0, 4700.0).astype(np.float32)


def marmousi_2d(space_order=8, timesteps=None, seed=1, dtype=np.float32):
    """C2: Marmousi-shaped 351 x 1701 model (benchmark/marmousi_2D.py:71-113;
    the script itself builds it in float64, ``dtype=np.float64``)."""
    vel = marmousi_2d_velocity(seed)
    return _assemble(
        vel, None, ((0, 70), (70, 70)), (10.0, 10.0), space_order,
        ("null_neumann", "null_dirichlet", "null_dirichlet", "null_dirichlet"),
        [(20.0, 8500.0)], [(20.0, 10.0 * i) for i in range(1700)],
        1, 10.0, 2.0, timesteps, dtype=dtype, name="marmousi_2d")


def _layered_3d(shape, vmin, vmax, seed, step_axis=2):
    nz = shape[0]
    rng = np.random.default_rng(seed)
    z = np.arange(nz, dtype=np.float32)[:, None, None]
    vel = np.empty(shape, dtype=np.float32)
    vel[:] = vmin + (vmax - vmin) * (z / (nz - 1))
    # thrust-like lateral step: layers shifted upwards on one side
    half = shape[step_axis] // 2
    shift = max(1, nz // 10)
    idx = [slice(None)] * 3
    idx[step_axis] = slice(half, None)
    shifted = np.roll(vel[tuple(idx)], -shift, axis=0)
    shifted[-shift:] = vmax
    vel[tuple(idx)] = shifted
    coarse = rng.standard_normal(tuple((n + 31) // 32 for n in shape)).astype(np.float32)
    pert = np.kron(coarse, np.ones((32, 32, 32), dtype=np.float32))
    pert = pert[:shape[0], :shape[1], :shape[2]]
    vel *= (1.0 + 0.03 * np.tanh(pert))
    return np.clip(vel, vmin, vmax).astype(np.float32)


def overthrust_3d(space_order=8, timesteps=None, seed=2, dtype=np.float32):
    """C3: Overthrust-shaped 207 x 801 x 801 model, no damping layer
    (benchmark/overthrust_3D.py:77-114).  The bench line is float32 (what
    north_star asks for); the reference's script itself builds a float64 model
    (overthrust_3D.py:82), which ``dtype=np.float64`` reproduces."""
    vel = _layered_3d((207, 801, 801), 2179.0, 6000.0, seed)
    return _assemble(
        vel, None, ((0, 0),) * 3, (20.0, 20.0, 20.0), space_order,
        ("null_neumann", "null_dirichlet", "null_dirichlet", "null_dirichlet",
         "null_dirichlet", "null_dirichlet"),
        [(20.0, 8000.0, 8000.0)],
        [(20.0, 8000.0, 20.0 * i) for i in range(800)],
        1, 8.0, 4.0, timesteps, dtype=dtype, name="overthrust_3d")


def variable_density_3d(n=1024, space_order=16, timesteps=200, nz=None):
    """C4: n^3 variable density, order 16, 40-point cubic damping layers on
    every side but the top.  Built directly in extended form.  ``nz`` selects
    a slab height (weak-scaling runs stack slabs along z)."""
    r = space_order // 2
    nbl = ((0, 40), (40, 40), (40, 40))
    nz = n if nz is None else nz
    phys = tuple(m - b - a for m, (b, a) in zip((nz, n, n), nbl))
    rng3, rng4 = np.random.default_rng(3), np.random.default_rng(4)

    def smooth(rng):
        coarse = rng.random(tuple((m + 63) // 64 + 1 for m in phys)).astype(np.float32)
        up = np.kron(coarse, np.ones((64, 64, 64), dtype=np.float32))
        return up[:phys[0], :phys[1], :phys[2]]

    vel = (1500.0 + 3000.0 * smooth(rng3)).astype(np.float32)
    den = (1000.0 + 1500.0 * smooth(rng4)).astype(np.float32)
    h = (10.0, 10.0, 10.0)
    size = [(m - 1) * s for m, s in zip(phys, h)]
    src = [(20.0, size[1] / 2, size[2] / 2)]
    rec = [(20.0, size[1] / 2, size[2] * i / 1023.0) for i in range(1024)]
    p = _assemble(
        vel, den, nbl, h, space_order,
        ("null_neumann", "null_dirichlet", "null_dirichlet", "null_dirichlet",
         "null_dirichlet", "null_dirichlet"),
        src, rec, 4, 8.0, 4.0, timesteps, name="variable_density_3d")
    del r
    return p


def shot_3d(shot=0, n=512, space_order=8, timesteps=300, seed=5):
    """C5: one shot of the 64-shot survey over a shared n^3 model."""
    vel = _layered_3d((n, n, n), 1500.0, 4500.0, seed)
    h = (10.0, 10.0, 10.0)
    x = 80.0 * shot + 40.0
    rec = [(20.0, x, 10.0 * i) for i in range(n)]
    return _assemble(
        vel, None, ((0, 0),) * 3, h, space_order,
        ("null_neumann", "null_dirichlet", "null_dirichlet", "null_dirichlet",
         "null_dirichlet", "null_dirichlet"),
        [(20.0, x, (n - 1) * 5.0)], rec, 4, 10.0, 2.0, timesteps,
        name="shot_3d")


def reshoot(base, shot):
    """Problem of another shot of the C5 survey over the SAME model arrays as
    ``base`` (a ``shot_3d`` problem): only the source / receiver tables move
    (the shot line sits at x = 80 m * shot + 40 m); ``u`` and ``receivers`` are
    fresh.  A survey driver keeps one model on the host and calls ``forward``
    once per shot with these."""
    dtype = base["velocity"].dtype.type
    n = base["velocity"].shape[0] - base["space_order"]
    r = base["space_order"] // 2
    h = np.array([dtype(v) for v in base["spacing"]], dtype=dtype)
    origin = np.array([r, r, r], dtype=dtype)        # no damping layers in C5
    x = 80.0 * shot + 40.0

    def to_grid(coords):
        return np.asarray(coords, dtype=dtype) / h + origin
    src = [(20.0, x, (n - 1) * 5.0)]
    rec = [(20.0, x, 10.0 * i) for i in range(n)]
    shape = base["velocity"].shape
    p = dict(base)
    p["src_intervals"], p["src_values"], p["src_offsets"] = _tables(
        shape, to_grid(src), 4, dtype)
    p["rec_intervals"], p["rec_values"], p["rec_offsets"] = _tables(
        shape, to_grid(rec), 4, dtype)
    # np.zeros maps untouched zero pages; zeros_like would write 1.7 GB per shot
    p["u"] = np.zeros(base["u"].shape, dtype=base["u"].dtype)
    p["receivers"] = np.zeros(base["receivers"].shape, dtype=base["receivers"].dtype)
    p["shot"] = shot
    return p


def slab_3d(rank=0, world=1, planes_per_gpu=128, n=1040, space_order=16,
            density=True, timesteps=60):
    """C4-shaped slab workload for weak scaling: the extended grid is
    (world*planes_per_gpu + 2r) x n x n (variable density, order 16, 40-point
    cubic damping layers on every side but the top); this builds ONLY the
    arrays of ``rank``'s z-slab (owned planes plus r ghost / halo planes on
    either side), straight from closed-form model functions, so that no rank
    ever holds the global model.  Returns the local problem dict with
    slab_up / slab_down set (simwave_b200/slab.py)."""
    from simwave_b200 import slab
    dtype = np.float32
    r = space_order // 2
    nz = world * planes_per_gpu + 2 * r
    lo, hi = slab.split_planes(nz, r, world)[rank]
    a, b = lo - r, hi + r
    h = (10.0, 10.0, 10.0)
    nbl = ((0, 40), (40, 40), (40, 40))
    vmin, vmax = 1500.0, 4500.0

    # physical-domain index of every extended plane/row/column (edge padding
    # = clipping the index, like SpaceModel.extended_velocity_model)
    def phys(count, before, after):
        idx = np.arange(count, dtype=np.float64) - (before + r)
        return np.clip(idx, 0, count - 2 * r - before - after - 1)
    zfull, xfull, yfull = phys(nz, *nbl[0]), phys(n, *nbl[1]), phys(n, *nbl[2])

    # damping mask: what np.pad(linear_ramp, end_values=nbl) builds axis by
    # axis (SpaceModel.damping_mask), in closed form
    def layer_depth(count, before, after):
        idx = np.arange(count, dtype=np.float64)
        d = np.zeros(count)
        inner0, inner1 = r + before, count - r - after
        d = np.where(idx < inner0, inner0 - idx, d)
        d = np.where(idx >= inner1, idx - inner1 + 1, d)
        d[:r] = -1
        d[count - r:] = -1                       # halo: no damping
        return d
    dzfull, dxfull, dyfull = (layer_depth(nz, *nbl[0]), layer_depth(n, *nbl[1]),
                              layer_depth(n, *nbl[2]))
    nx_l, ny_l = float(nbl[1][0]), float(nbl[2][0])
    zscale = max(1.0, float(nz - 2 * r - nbl[0][1] - 1))

    # The fields are closed-form functions of the indices, evaluated in float64
    # a block of planes at a time -- on the GPU when there is one (a 1040^3
    # slab is 1.1e9 points per field), with NumPy otherwise; same expressions.
    try:
        import torch
        xp = torch if torch.cuda.is_available() else None
    except ImportError:
        xp = None

    def block(z0, z1):
        if xp is not None:
            dev = "cuda"
            t = lambda v: torch.as_tensor(v, dtype=torch.float64, device=dev)  # noqa: E731
            sin, cos, clip, where, maximum = (torch.sin, torch.cos, torch.clamp,
                                              torch.where, torch.clamp_min)
        else:
            t = lambda v: np.asarray(v, dtype=np.float64)                      # noqa: E731
            sin, cos, clip, where = np.sin, np.cos, np.clip, np.where
            maximum = np.maximum
        zg = t(zfull[z0:z1])[:, None, None]
        xg = t(xfull)[None, :, None]
        yg = t(yfull)[None, None, :]
        bumps = (sin(zg / 37.0 + 0.3) * cos(xg / 53.0) +
                 sin(yg / 41.0 + 1.1) * cos(zg / 61.0 + xg / 97.0))
        depth = zg / zscale
        vel = vmin + (vmax - vmin) * clip(0.15 + 0.6 * depth + 0.12 * bumps, 0, 1)
        rho = None
        if density:
            rho = 1000.0 + 1500.0 * clip(0.2 + 0.5 * depth + 0.15 * cos(
                xg / 45.0 + yg / 58.0 + zg / 71.0), 0, 1)
        dz = t(dzfull[z0:z1])[:, None, None]
        dx = t(dxfull)[None, :, None]
        dy = t(dyfull)[None, None, :]
        m = maximum(dz, 0.0) + 0 * dx + 0 * dy
        m = m + (nx_l - m) * maximum(dx, 0.0) / nx_l
        m = m + (ny_l - m) * maximum(dy, 0.0) / ny_l
        m = where((dz < 0) | (dx < 0) | (dy < 0), 0.0 * m, m)
        damp = 0.001 * m ** 3
        out = []
        for f in (vel, rho, damp):
            if f is None:
                out.append(None)
            elif xp is not None:
                out.append(f.to(torch.float32).cpu().numpy())
            else:
                out.append(f.astype(dtype))
        return out

    vel = np.empty((b - a, n, n), dtype=dtype)
    rho = np.empty((b - a, n, n), dtype=dtype) if density else None
    damp = np.empty((b - a, n, n), dtype=dtype)
    step = 64
    for z0 in range(a, b, step):
        z1 = min(b, z0 + step)
        v_, r_, d_ = block(z0, z1)
        vel[z0 - a:z1 - a] = v_
        damp[z0 - a:z1 - a] = d_
        if density:
            rho[z0 - a:z1 - a] = r_

    hf = [dtype(x) for x in h]
    dt = dtype(fd.calculate_dt(3, space_order, hf, np.array([vmax], dtype=dtype)))
    shape = (nz, n, n)
    origin = np.array([nbl[0][0] + r, nbl[1][0] + r, nbl[2][0] + r], dtype=dtype)
    size = [(m_ - 2 * r - bb - aa - 1) * s for m_, (bb, aa), s in zip(shape, nbl, h)]

    def to_grid(coords):
        return np.asarray(coords, dtype=dtype) / np.array(hf, dtype=dtype) + origin
    src = [(20.0, size[1] / 2, size[2] / 2)]
    rec = [(20.0, size[1] / 2, size[2] * i / 1023.0) for i in range(1024)]
    own_lo = lo if rank > 0 else 0
    own_hi = hi if rank < world - 1 else nz
    tabs = {}
    for kind, coords in (("src", src), ("rec", rec)):
        iv, val, off = _tables(shape, to_grid(coords), 4, dtype)
        tabs[kind] = slab._clip_tables(iv, val, off, 3, own_lo, own_hi, a)
    wavelet = _ricker(8.0, 0.0, 4.0, max(timesteps, int(4.0 / float(dt)) + 1),
                      dtype)[:timesteps].copy()
    bc = [BC[x] for x in ("null_neumann", "null_dirichlet", "null_dirichlet",
                          "null_dirichlet", "null_dirichlet", "null_dirichlet")]
    if rank > 0:
        bc[0] = 0
    if rank < world - 1:
        bc[1] = 0
    return {
        "name": "slab_3d",
        "u": np.zeros((3,) + vel.shape, dtype=dtype),
        "velocity": np.ascontiguousarray(vel),
        "density": None if rho is None else np.ascontiguousarray(rho),
        "damp": np.ascontiguousarray(damp), "wavelet": wavelet,
        "coeff2": dtype(fd.half_coefficients(2, space_order)),
        "coeff1": dtype(fd.half_coefficients(1, space_order)),
        "bc": np.asarray(bc, dtype=np.uint64),
        "src_intervals": tabs["src"][0], "src_values": tabs["src"][1],
        "src_offsets": tabs["src"][2],
        "rec_intervals": tabs["rec"][0], "rec_values": tabs["rec"][1],
        "rec_offsets": tabs["rec"][2],
        "receivers": np.zeros((timesteps, 1024), dtype=dtype),
        "spacing": h, "saving_stride": 0, "dt": dt, "end_timestep": timesteps,
        "space_order": space_order, "full_timesteps": timesteps,
        "slab_up": int(rank > 0), "slab_down": int(rank < world - 1),
        "global_shape": shape, "owned_planes": hi - lo,
    }


# ---------------------------------------------------------------------------
# The same configurations through the public API (SpaceModel, TimeModel,
# Source, Receiver, RickerWavelet, Solver): what a simwave user writes.  The
# objects produce the same arrays as the builders above (tests/test_workloads.py).
_API_SPECS = {
    # name: (velocity, bounding box, spacing, space_order, damping length, bc,
    #        tf, source coordinates, receiver coordinates, window radius, f0)
    "readme_2d": lambda: (
        readme_2d_velocity(), (0, 5120, 0, 5120), (10., 10.), 4, 0,
        ("null_neumann", "null_dirichlet", "none", "null_dirichlet"), 1.0,
        [(2560., 2560.)], [(2560., 10. * i) for i in range(512)], 1, 10.0),
    "marmousi_2d": lambda: (
        marmousi_2d_velocity(), (0, 3500, 0, 17000), (10., 10.), 8,
        (0, 700, 700, 700),
        ("null_neumann", "null_dirichlet", "null_dirichlet", "null_dirichlet"),
        2.0, [(20., 8500.)], [(20., 10. * i) for i in range(1700)], 1, 10.0),
    "overthrust_3d": lambda: (
        _layered_3d((207, 801, 801), 2179.0, 6000.0, 2),
        (0, 4120, 0, 16000, 0, 16000), (20., 20., 20.), 8, 0,
        ("null_neumann", "null_dirichlet", "null_dirichlet", "null_dirichlet",
         "null_dirichlet", "null_dirichlet"), 4.0,
        [(20., 8000., 8000.)], [(20., 8000., 20. * i) for i in range(800)], 1, 8.0),
}


def api_solver(name, timesteps=None, compiler=None):
    """``simwave_b200.Solver`` of workload ``name`` built the way the
    reference's benchmark scripts build it (benchmark/overthrust_3D.py:77-126,
    benchmark/marmousi_2D.py:71-126, README.md:54-137).  ``timesteps`` shortens
    the run (tf is cut accordingly)."""
    import simwave_b200 as api
    (vel, box, h, order, damping, bc, tf, src, rec, radius, f0) = _API_SPECS[name]()
    space = api.SpaceModel(bounding_box=box, grid_spacing=h, velocity_model=vel,
                           space_order=order, dtype=np.float32)
    space.config_boundary(damping_length=damping, boundary_condition=bc,
                          damping_polynomial_degree=3, damping_alpha=0.001)
    time = api.TimeModel(space_model=space, tf=tf)
    if timesteps is not None and timesteps < time.timesteps:
        time = api.TimeModel(space_model=space,
                             tf=float(time.dt) * (timesteps - 1) * (1 - 1e-6))
    source = api.Source(space, coordinates=src, window_radius=radius)
    receiver = api.Receiver(space_model=space, coordinates=rec, window_radius=radius)
    wavelet = api.RickerWavelet(f0, time)
    return api.Solver(space_model=space, time_model=time, sources=source,
                      receivers=receiver, wavelet=wavelet, compiler=compiler)


WORKLOADS = {
    "readme_2d": readme_2d,
    "marmousi_2d": marmousi_2d,
    "overthrust_3d": overthrust_3d,
    "variable_density_3d": variable_density_3d,
    "shot_3d": shot_3d,
    "slab_3d": slab_3d,
}
