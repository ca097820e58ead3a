"""
ABI-level synthetic problems: the arrays `forward` receives, built directly
(no SpaceModel), seeded, with every feature switched on -- heterogeneous
velocity and density, damping layers, any boundary-condition mix, several
sources with overlapping windows, per-source wavelets, snapshots.

The same dict feeds the CPU checkers (oracle/oracle.py) and the CUDA library
(simwave_b200's C-ABI), so parity tests compare like with like.
"""
import numpy as np

from simwave_b200.kernel.frontend import fd, kws


def damping_profile(shape, nbl, halo, alpha, degree, dtype):
    """alpha * d**degree in the layers, 0 in the physical domain and halo
    (same construction as SpaceModel.damping_mask)."""
    inner = tuple(n - 2 * halo - b - a for n, (b, a) in zip(shape, nbl))
    mask = np.pad(np.zeros(inner, dtype=dtype), nbl, mode="linear_ramp",
                  end_values=nbl)
    mask = (mask ** degree) * alpha
    return np.pad(mask, [(halo, halo)] * len(shape)).astype(dtype)


def tables(shape, positions, radius, dtype):
    """Interval / weight / offset tables for grid positions (in extended
    grid points) with Kaiser-windowed-sinc weights."""
    intervals, values, offsets = [], [], [0]
    for pos in positions:
        p, v = kws.get_source_points(shape, [dtype(x) for x in pos], radius)
        intervals.append(p)
        values.append(v)
        offsets.append(offsets[-1] + v.size)
    return (np.concatenate(intervals).astype(np.uint64),
            np.concatenate(values).astype(dtype),
            np.asarray(offsets, dtype=np.uint64))


def make_problem(shape, space_order=8, density=False, dtype=np.float32,
                 timesteps=40, saving_stride=0, bc=None, nbl=None,
                 num_sources=2, num_receivers=12, src_radius=4, rec_radius=4,
                 multi_wavelet=False, seed=0, spacing=None, vmin=1500.0,
                 vmax=4500.0, alpha=0.002, degree=2, smooth_density=False,
                 src_positions=None, rec_positions=None, f0=None):
    """
    Returns a dict with the reference ABI's arrays (names follow
    oracle.abi_args).  ``shape`` is the extended shape (halo + layers
    included).  ``nbl`` is ((before, after), ..) per axis.
    """
    dtype = np.dtype(dtype).type
    rng = np.random.default_rng(seed)
    ndim = len(shape)
    r = space_order // 2
    if spacing is None:
        spacing = (10.0, 12.5, 8.0)[:ndim]
    if bc is None:
        bc = (2, 1, 0, 1, 2, 1)[:2 * ndim]
    if nbl is None:
        nbl = ((0, 0),) * ndim

    velocity = (vmin + (vmax - vmin) * rng.random(shape)).astype(dtype)
    rho = None
    if density:
        if smooth_density:
            axes = np.meshgrid(*[np.linspace(0, 1, n) for n in shape],
                               indexing="ij")
            rho = 1000.0 + 600.0 * sum(np.sin(3.0 * (a + 0.3 * i))
                                       for i, a in enumerate(axes))
            rho = rho.astype(dtype)
        else:
            rho = (1000.0 + 1500.0 * rng.random(shape)).astype(dtype)
    damp = damping_profile(shape, nbl, r, alpha, degree, dtype)

    c2 = dtype(fd.half_coefficients(2, space_order))
    c1 = dtype(fd.half_coefficients(1, space_order))
    dt = dtype(0.9 * fd.calculate_dt(ndim, space_order,
                                     [dtype(h) for h in spacing], velocity))

    def random_positions(count):
        lo = np.array([r + 0.5] * ndim)
        hi = np.array([n - r - 1.5 for n in shape])
        return lo + (hi - lo) * rng.random((count, ndim))

    if src_positions is None:
        src_positions = random_positions(num_sources)
    if rec_positions is None:
        rec_positions = random_positions(num_receivers)
    src_iv, src_val, src_off = tables(shape, src_positions, src_radius, dtype)
    rec_iv, rec_val, rec_off = tables(shape, rec_positions, rec_radius, dtype)
    num_sources = len(src_off) - 1
    num_receivers = len(rec_off) - 1

    # Ricker, peak well inside the run
    if f0 is None:
        f0 = 4.0 / (timesteps * float(dt))
    t = np.arange(timesteps) * float(dt)
    arg = np.pi * f0 * (t - 1.0 / f0)
    ricker = ((1 - 2 * arg ** 2) * np.exp(-arg ** 2)).astype(dtype)
    if multi_wavelet:
        scale = np.array([(1 + 0.5 * s) * (-1) ** s
                          for s in range(num_sources)])
        wavelet = np.ascontiguousarray(ricker[:, None] * scale[None, :],
                                       dtype=dtype)
        # exercise the "skip when exactly zero" branch per source
        wavelet[::7, 0] = 0
    else:
        wavelet = ricker
        wavelet[::9] = 0

    if saving_stride == 0:
        slots = 3
    else:
        assert saving_stride == 1 or timesteps % saving_stride == 1
        slots = len(range(0, timesteps, saving_stride)) + 2

    return {
        "u": np.zeros((slots,) + tuple(shape), dtype=dtype),
        "velocity": velocity, "density": rho, "damp": damp,
        "wavelet": wavelet, "coeff2": c2, "coeff1": c1,
        "bc": np.asarray(bc, dtype=np.uint64),
        "src_intervals": src_iv, "src_values": src_val, "src_offsets": src_off,
        "rec_intervals": rec_iv, "rec_values": rec_val, "rec_offsets": rec_off,
        "receivers": np.zeros((timesteps, num_receivers), dtype=dtype),
        "spacing": spacing, "saving_stride": saving_stride, "dt": dt,
        "end_timestep": timesteps, "space_order": space_order,
    }


def clone(p):
    """Deep copy (so two implementations can run on identical inputs)."""
    return {k: (v.copy() if isinstance(v, np.ndarray) else v)
            for k, v in p.items()}
