"""
Case definitions shared by the golden-vector generator (which runs them
through the *reference* package, imported from /root/reference) and by the
tests (which run them through simwave_b200).  Every builder takes the API
module as its first argument -- the two packages expose the same names, which
is the point of the drop-in.

The ``solution`` and ``u_saving`` / ``parallel`` configurations are the ones
of the reference's own tests (tests/test_solution.py:10-139,
tests/test_u_saving.py:7-146, tests/test_parallel_solution.py:9-104).
"""
import numpy as np

# --------------------------------------------------------------------------
# front-end cases: (name, dict) -> tables, masks, coefficients
# --------------------------------------------------------------------------
FRONTEND_CASES = {
    # reference tests/test_source.py:111-164 geometries
    "src2d_w1": dict(dimension=2, shape=(50, 50), bbox=(0, 500, 0, 500),
                     spacing=(10, 10), space_order=2, damping_length=50,
                     dtype="float32", window_radius=1,
                     coords=[(250, 250)]),
    "src3d_w4": dict(dimension=3, shape=(50, 50, 50),
                     bbox=(0, 500, 0, 500, 0, 500), spacing=(10, 10, 10),
                     space_order=4, damping_length=0, dtype="float32",
                     window_radius=4, coords=[(255, 250, 100)]),
    # off-grid positions, windows clipped by the grid edge, all radii
    "line2d_w4": dict(dimension=2, shape=(64, 96), bbox=(0, 630, 0, 950),
                      spacing=(10, 10), space_order=8,
                      damping_length=(0, 70, 30, 30), dtype="float32",
                      window_radius=4,
                      coords=[(3.0, 7.5 * i + 1.25) for i in range(120)]),
    "edge3d_w8": dict(dimension=3, shape=(20, 24, 28),
                      bbox=(0, 190, 0, 230, 0, 270), spacing=(10, 10, 10),
                      space_order=4, damping_length=(0, 20, 10, 10, 0, 30),
                      dtype="float32", window_radius=8,
                      coords=[(0, 0, 0), (190, 230, 270), (95.5, 3.25, 266.0),
                              (10, 115, 135), (187.5, 229.9, 0.1)]),
    "f64_2d_w3": dict(dimension=2, shape=(40, 40), bbox=(-40, 440, -40, 440),
                      spacing=(12.0, 12.0), space_order=6, damping_length=24,
                      dtype="float64", window_radius=3,
                      coords=[(200, 200), (260.5, 13.7), (-40, 440)]),
    "f64_3d_w10": dict(dimension=3, shape=(16, 16, 16),
                       bbox=(0, 150, 0, 150, 0, 150), spacing=(10, 5, 7.5),
                       space_order=20, damping_length=0, dtype="float64",
                       window_radius=10,
                       coords=[(75, 75, 75), (1.0, 149.0, 33.3)]),
}
for _w in range(1, 11):
    FRONTEND_CASES["radius2d_w%d" % _w] = dict(
        dimension=2, shape=(32, 32), bbox=(0, 310, 0, 310), spacing=(10, 10),
        space_order=2 * _w, damping_length=20, dtype="float32",
        window_radius=_w, coords=[(155.0, 152.5), (7.0, 301.0)])


def build_space_model(api, case, velocity=None, density=None):
    dtype = np.dtype(case["dtype"]).type
    shape = tuple(case["shape"])
    if velocity is None:
        velocity = np.full(shape, 1500.0, dtype=dtype)
    model = api.SpaceModel(
        bounding_box=case["bbox"], grid_spacing=case["spacing"],
        velocity_model=velocity, density_model=density,
        space_order=case["space_order"], dtype=dtype)
    model.config_boundary(
        damping_length=case["damping_length"],
        boundary_condition=case.get("boundary_condition", "none"),
        damping_polynomial_degree=case.get("degree", 3),
        damping_alpha=case.get("alpha", 0.001))
    return model


def frontend_outputs(api, case):
    """Everything the front end feeds to the kernel for one case."""
    model = build_space_model(api, case)
    src = api.Source(model, coordinates=case["coords"],
                     window_radius=case["window_radius"])
    points, values, offsets = src.interpolated_points_and_values
    time_model = api.TimeModel(space_model=model, tf=0.25)
    ricker = api.RickerWavelet(12.0, time_model)
    return {
        "shape": np.array(model.shape),
        "extended_shape": np.array(model.extended_shape),
        "nbl": np.array(model.nbl),
        "grid_positions": src.grid_positions,
        "adjusted_grid_positions": src.adjusted_grid_positions,
        "points": points, "values": values, "offsets": offsets,
        "damping_mask": model.damping_mask,
        "coeff2": model.fd_coefficients(2),
        "coeff1": model.fd_coefficients(1),
        "dt": np.array(time_model.dt),
        "timesteps": np.array(time_model.timesteps),
        "ricker": ricker.values,
    }


# --------------------------------------------------------------------------
# end-to-end cases (Solver.forward through the public API)
# --------------------------------------------------------------------------
def solution_solver(api, dimension, space_order, compiler, density=False,
                    density_value=1.0, saving_stride=0):
    """Reference tests/test_solution.py:26-131 (and the golden generator
    tests/reference_solution/generator.py:6-98)."""
    if dimension == 2:
        shape = (500,) * 2
        bbox = (0, 5000, 0, 5000)
        spacing = (10, 10)
        damping_length = 100
        bc = ("null_neumann", "null_dirichlet", "none", "null_dirichlet")
        position = [(2500, 2500)]
        tf, f0 = 1.0, 10.0
    else:
        shape = (100,) * 3
        bbox = (0, 1000, 0, 1000, 0, 1000)
        spacing = (10, 10, 10)
        damping_length = 50
        bc = ("null_neumann", "null_dirichlet", "none", "null_dirichlet",
              "null_neumann", "null_dirichlet")
        position = [(500, 495, 505)]
        tf, f0 = 0.4, 15.0

    vel = np.zeros(shape=shape, dtype=np.float32)
    vel[:] = 1500.0
    vel[shape[0] // 2:] = 2000.0
    den = None
    if density:
        den = np.zeros(shape=shape, dtype=np.float32)
        den[:] = density_value

    space_model = api.SpaceModel(
        bounding_box=bbox, grid_spacing=spacing, velocity_model=vel,
        density_model=den, space_order=space_order, dtype=np.float32)
    space_model.config_boundary(
        damping_length=damping_length, boundary_condition=bc,
        damping_polynomial_degree=3, damping_alpha=0.001)
    time_model = api.TimeModel(space_model=space_model, tf=tf,
                               saving_stride=saving_stride)
    source = api.Source(space_model=space_model, coordinates=position,
                        window_radius=1)
    receiver = api.Receiver(space_model=space_model, coordinates=position,
                            window_radius=1)
    ricker = api.RickerWavelet(f0, time_model)
    return api.Solver(space_model=space_model, time_model=time_model,
                      sources=source, receivers=receiver, wavelet=ricker,
                      compiler=compiler)


def u_saving_solver(api, dimension, density, saving_stride, compiler):
    """Reference tests/test_u_saving.py:7-146."""
    if dimension == 2:
        vel = np.zeros(shape=(512, 512), dtype=np.float32)
        vel[:] = 1500.0
        vel[250:] = 3000.0
        bbox, spacing = (0, 5120, 0, 5120), (10, 10)
        bc = ("null_neumann", "null_dirichlet", "none", "null_dirichlet")
        src = [(2560, 2560)]
        rec = [(2560, i) for i in range(0, 5120, 10)]
        tf, f0 = 1.0, 10.0
    else:
        vel = np.zeros(shape=(100, 100, 100), dtype=np.float32)
        vel[:] = 1500.0
        bbox, spacing = (0, 1000, 0, 1000, 0, 1000), (10, 10, 10)
        bc = ("null_neumann", "null_dirichlet", "null_dirichlet",
              "null_dirichlet", "null_dirichlet", "null_dirichlet")
        src = [(500, 500, 500)]
        rec = [(500, 500, i) for i in range(0, 1000, 10)]
        tf, f0 = 0.4, 15.0
    den = None
    if density:
        den = np.zeros(shape=vel.shape, dtype=np.float32)
        den[:] = 5

    space_model = api.SpaceModel(
        bounding_box=bbox, grid_spacing=spacing, velocity_model=vel,
        density_model=den, space_order=4, dtype=np.float32)
    space_model.config_boundary(
        damping_length=0, boundary_condition=bc,
        damping_polynomial_degree=3, damping_alpha=0.001)
    time_model = api.TimeModel(space_model=space_model, tf=tf,
                               saving_stride=saving_stride)
    source = api.Source(space_model, coordinates=src, window_radius=4)
    receiver = api.Receiver(space_model=space_model, coordinates=rec,
                            window_radius=4)
    ricker = api.RickerWavelet(f0, time_model)
    return api.Solver(space_model=space_model, time_model=time_model,
                      sources=source, receivers=receiver, wavelet=ricker,
                      compiler=compiler)


def parallel_solver(api, dimension, density, dtype, compiler):
    """Reference tests/test_parallel_solution.py:9-104: multi-source,
    window_radius 8, damping layer, NN/ND on every axis."""
    if dimension == 2:
        shape = (128, 128)
        bbox = (0, 1280, 0, 1280)
        spacing = (10., 10.)
        bc = ("null_neumann", "null_dirichlet") * 2
        src = [(10, i) for i in range(128, 1280, 128)]
        rec = [(10, i) for i in range(0, 1280, 10)]
    else:
        shape = (128, 128, 128)
        bbox = (0, 1280, 0, 1280, 0, 1280)
        spacing = (10., 10., 10.)
        bc = ("null_neumann", "null_dirichlet") * 3
        src = [(10, 640, i) for i in range(128, 1280, 128)]
        rec = [(10, 650, i) for i in range(0, 1280, 10)]
    den = None
    if density:
        den = np.zeros(shape=shape, dtype=dtype)
        den[:] = 5
    vel = np.zeros(shape=shape, dtype=dtype)
    vel[:] = 1500.0

    space_model = api.SpaceModel(
        bounding_box=bbox, grid_spacing=spacing, velocity_model=vel,
        density_model=den, space_order=4, dtype=dtype)
    space_model.config_boundary(damping_length=128, boundary_condition=bc)
    time_model = api.TimeModel(space_model=space_model, tf=0.4,
                               saving_stride=0)
    source = api.Source(space_model, coordinates=src, window_radius=8)
    receiver = api.Receiver(space_model=space_model, coordinates=rec,
                            window_radius=8)
    ricker = api.RickerWavelet(10.0, time_model)
    return api.Solver(space_model=space_model, time_model=time_model,
                      sources=source, receivers=receiver, wavelet=ricker,
                      compiler=compiler)


# small heterogeneous end-to-end cases: random velocity AND density, damping,
# mixed boundary conditions, several sources with their own wavelets
SMALL_CASES = {
    "small2d_const_f32": dict(dimension=2, density=False, dtype="float32",
                              space_order=8, stride=0),
    "small2d_var_f32": dict(dimension=2, density=True, dtype="float32",
                            space_order=4, stride=3),
    "small3d_const_f32": dict(dimension=3, density=False, dtype="float32",
                              space_order=8, stride=0),
    "small3d_var_f32": dict(dimension=3, density=True, dtype="float32",
                            space_order=6, stride=2),
    "small2d_var_f64": dict(dimension=2, density=True, dtype="float64",
                            space_order=12, stride=0),
    "small3d_var_f64": dict(dimension=3, density=True, dtype="float64",
                            space_order=4, stride=1),
}


def small_solver(api, name, compiler):
    cfg = SMALL_CASES[name]
    dtype = np.dtype(cfg["dtype"]).type
    rng = np.random.default_rng(sum(map(ord, name)))
    if cfg["dimension"] == 2:
        shape = (44, 52)
        bbox = (0, 430, 0, 510)
        spacing = (10., 10.)
        damping = (0, 40, 30, 50)
        bc = ("null_neumann", "null_dirichlet", "none", "null_dirichlet")
        src = [(215.0, 255.0), (12.5, 100.0), (300.0, 480.0)]
        rec = [(20.0, 10.0 * i + 2.5) for i in range(50)]
    else:
        # cubic on purpose: the reference's variable-density kernel steps the
        # x first derivatives by nx instead of ny
        shape = (22, 22, 22)
        bbox = (0, 210, 0, 210, 0, 210)
        spacing = (10., 10., 10.)
        damping = (0, 30, 20, 20, 30, 0)
        bc = ("null_neumann", "null_dirichlet", "none", "null_dirichlet",
              "null_neumann", "none")
        src = [(105.0, 100.0, 110.0), (5.0, 30.0, 200.0)]
        rec = [(15.0, 105.0, 10.0 * i + 5.0) for i in range(20)]

    vel = (1500.0 + 2500.0 * rng.random(shape)).astype(dtype)
    den = None
    if cfg["density"]:
        den = (1000.0 + 1500.0 * rng.random(shape)).astype(dtype)

    space_model = api.SpaceModel(
        bounding_box=bbox, grid_spacing=spacing, velocity_model=vel,
        density_model=den, space_order=cfg["space_order"], dtype=dtype)
    space_model.config_boundary(
        damping_length=damping, boundary_condition=bc,
        damping_polynomial_degree=2, damping_alpha=0.002)
    time_model = api.TimeModel(space_model=space_model, tf=0.12,
                               saving_stride=cfg["stride"])
    source = api.Source(space_model, coordinates=src, window_radius=3)
    receiver = api.Receiver(space_model=space_model, coordinates=rec,
                            window_radius=2)
    base = api.RickerWavelet(18.0, time_model).values
    multi = np.stack([base * (1.0 + 0.5 * s) * (-1) ** s
                      for s in range(len(src))], axis=1)
    wavelet = api.MultiWavelet(multi, time_model)
    return api.Solver(space_model=space_model, time_model=time_model,
                      sources=source, receivers=receiver, wavelet=wavelet,
                      compiler=compiler)
