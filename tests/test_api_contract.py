"""
The API the drop-in keeps: known answers the reference's own front-end tests
pin for ``Compiler``, ``SpaceModel`` and ``TimeModel`` (tests/test_compiler.py,
tests/test_space_model.py, tests/test_time_model.py of the reference), restated
as tables and checked against simwave_b200's mirror of those classes.  A user
script written for simwave must see the same shapes, layer widths, step counts,
flags and errors here.
"""
import numpy as np
import pytest

from simwave_b200 import Compiler, SpaceModel, TimeModel


def ones_model(dimension, points=50, **kwargs):
    extent = kwargs.pop("extent", 1000)
    return SpaceModel(bounding_box=(0, extent) * dimension,
                      grid_spacing=(10,) * dimension,
                      velocity_model=1500 * np.ones((points,) * dimension),
                      **kwargs)


# ---- Compiler (reference tests/test_compiler.py:7-48) -----------------------
CFLAGS = {
    # (cc, language, cflags given): cflags kept
    ("gcc", "c", None): "-O3 -fPIC -Wall -std=c99 -shared",
    ("gcc", "c", "-O3 -fPIC"): "-O3 -fPIC -shared",
    ("icc", "c", "-O3 -shared"): "-O3 -shared",
    ("icc", "c", "-O3"): "-O3 -shared",
    ("clang", "c", None): "-O3 -fPIC -Wall -std=c99 -shared",
    ("gcc", "cpu_openmp", "-O3 -fPIC"): "-O3 -fPIC -shared -fopenmp",
    ("pgcc", "cpu_openmp", "-shared"): "-shared -mp",
    ("icc", "cpu_openmp", "-O3"): "-O3 -shared -openmp",
    ("clang", "cpu_openmp", "-O3"): "-O3 -shared -fopenmp",
    ("foo", "cpu_openmp", "-O3"): "-O3 -shared",
}
OPENMP_FLAG = {"gcc": "-fopenmp", "icc": "-openmp", "pgcc": "-mp",
               "clang": "-fopenmp", "foo": None}


@pytest.mark.parametrize("key", sorted(CFLAGS, key=str))
def test_compiler_flags(key):
    cc, language, cflags = key
    assert Compiler(cc=cc, language=language, cflags=cflags).cflags == CFLAGS[key]


def test_compiler_attributes_and_validation():
    for cc, flag in OPENMP_FLAG.items():
        compiler = Compiler(cc=cc)
        assert compiler.cc == cc and compiler.get_openmp_flag() == flag
    for language in ("c", "cpu_openmp", "gpu_openmp", "gpu_openacc", "cuda"):
        assert Compiler(language=language).language == language
    with pytest.raises(ValueError):
        Compiler(language="cpu_mpi")
    with pytest.raises(TypeError):
        Compiler(cc=3)


# ---- SpaceModel (reference tests/test_space_model.py) ------------------------
@pytest.mark.parametrize("dimension", [2, 3])
def test_space_model_defaults(dimension):
    vel = 1500 * np.ones((101,) * dimension)
    den = 15 * np.ones((101,) * dimension)
    model = SpaceModel(bounding_box=(0, 1000) * dimension,
                       grid_spacing=(10,) * dimension, velocity_model=vel,
                       density_model=den, space_order=4, dtype=np.float32)
    assert model.bounding_box == (0, 1000) * dimension
    assert model.grid_spacing == (10,) * dimension
    assert np.array_equal(model.velocity_model, vel)
    assert np.array_equal(model.density_model, den)
    assert (model.space_order, model.dimension, model.dtype) == (4, dimension, np.float32)
    assert model.damping_length == (0.0,) * 2 * dimension
    assert model.boundary_condition == ("none",) * 2 * dimension
    assert model.damping_polynomial_degree == 3 and model.damping_alpha == 0.001
    assert model.grid.shape == model.shape and model.grid.dtype == model.dtype


SHAPES = [   # bounding box, spacing -> grid shape (:38-57)
    ((0, 1000, 0, 1000), (10, 10), (101, 101)),
    ((100., 1000, 0, 1250.5), (10, 10), (91, 126)),
    ((-100, 1000, 500, 1000), (20, 10), (56, 51)),
    ((0, 1000, 0, 1000, 0, 1000.0), (5, 10, 20), (201, 101, 51)),
    ((10, 100, -10, 100, 0.0, 100.0), (5, 5, 5), (19, 23, 21)),
]


@pytest.mark.parametrize("bbox,spacing,shape", SHAPES)
def test_space_model_shape(bbox, spacing, shape):
    model = SpaceModel(bounding_box=bbox, grid_spacing=spacing,
                       velocity_model=1500 * np.ones((50,) * len(shape)))
    assert model.shape == shape


EXTENDED = [   # damping length, space order -> extended shape (:59-91)
    (2, 100, 2, (123, 123)), (2, 100, 4, (125, 125)), (2, 150, 16, (147, 147)),
    (2, (150, 100, 150, 100), 8, (134, 134)), (2, (150, 100, 0, 100), 8, (134, 119)),
    (3, 100, 2, (123,) * 3), (3, 100, 4, (125,) * 3), (3, 150, 16, (147,) * 3),
    (3, (150, 100, 150, 100, 150, 100), 8, (134,) * 3),
    (3, (150, 100, 0, 100, 0, 150), 8, (134, 119, 124)),
]


@pytest.mark.parametrize("dimension,damping,order,extended", EXTENDED)
def test_space_model_extended_arrays(dimension, damping, order, extended):
    model = SpaceModel(bounding_box=(0, 1000) * dimension,
                       grid_spacing=(10,) * dimension,
                       velocity_model=1500 * np.ones((50,) * dimension),
                       density_model=15 * np.ones((50,) * dimension),
                       space_order=order)
    model.config_boundary(damping_length=damping)
    assert model.extended_shape == extended
    for array in (model.damping_mask, model.extended_grid,
                  model.extended_velocity_model, model.extended_density_model):
        assert array.shape == extended


@pytest.mark.parametrize("dimension,damping,nbl", [
    (2, None, (0,) * 4), (3, None, (0,) * 6), (2, 120, (12,) * 4),
    (3, 90, (9,) * 6), (2, (50, 60, 75, 80), (5, 6, 7, 8)),
    (3, (0, 10, 8, 50, 20, 30), (0, 1, 0, 5, 2, 3))])
def test_space_model_layer_widths(dimension, damping, nbl):
    model = ones_model(dimension)
    if damping is not None:
        model.config_boundary(damping_length=damping)
    assert model.nbl == nbl


@pytest.mark.parametrize("dimension,order", [(2, 2), (2, 4), (2, 20), (3, 2),
                                             (3, 8), (3, 10)])
def test_space_model_halo(dimension, order):
    model = ones_model(dimension, extent=500, space_order=order)
    assert model.halo_size == (order // 2,) * 2 * dimension


@pytest.mark.parametrize("dimension,damping,bc,degree,alpha", [
    (2, 500, "none", 2, 0.1),
    (2, (50, 50, 40, 40), "null_neumann", 4, 0.001),
    (3, 100, "null_dirichlet", 4, 0.001),
    (2, 75, ("none", "null_neumann", "none", "null_dirichlet"), 4, 0.001),
    (3, (5, 5, 5, 5, 6, 6), "null_dirichlet", 1, 0.001)])
def test_space_model_config_boundary(dimension, damping, bc, degree, alpha):
    model = ones_model(dimension, extent=500)
    model.config_boundary(damping_length=damping, boundary_condition=bc,
                          damping_polynomial_degree=degree, damping_alpha=alpha)
    faces = 2 * dimension
    assert model.damping_length == (
        (damping,) * faces if isinstance(damping, (int, float)) else damping)
    assert model.boundary_condition == ((bc,) * faces if isinstance(bc, str) else bc)
    assert (model.damping_alpha, model.damping_polynomial_degree) == (alpha, degree)


def test_space_model_rejects_bad_orders():
    for order in (3, 0, 22):
        with pytest.raises(ValueError):
            ones_model(2, space_order=order)


# ---- TimeModel (reference tests/test_time_model.py) --------------------------
TIMESTEPS = [   # dt, tf, t0, saving stride -> timesteps (:33-44)
    (0.001, 1.0, 0.0, 0, 1001), (0.001, 1.0, 0.5, 0, 501), (0.002, 2.0, 0.0, 0, 1001),
    (0.001, 1.0, 0.0, 1, 1001), (0.001, 1.0, 0.5, 1, 501), (0.002, 2.0, 0.0, 1, 1001),
    (0.001, 1.0, 0.0, 2, 1001), (0.001, 1.0, 0.5, 6, 505), (0.002, 2.0, 0.0, 3, 1003),
]


@pytest.mark.parametrize("dt,tf,t0,stride,timesteps", TIMESTEPS)
def test_time_model_timesteps(dt, tf, t0, stride, timesteps):
    space = SpaceModel(bounding_box=(0, 100, 0, 100), grid_spacing=(10, 10),
                       velocity_model=1500 * np.ones((10, 10)))
    time = TimeModel(space_model=space, tf=tf, t0=t0, saving_stride=stride)
    time.dt = dt
    assert time.timesteps == timesteps
    assert time.space_model is space
    assert (time.tf, time.t0, time.saving_stride) == (tf, t0, stride)
    assert time.dt == space.dtype(dt)


# ---- Source / Receiver (reference tests/test_source.py:5-103) ----------------
from simwave_b200 import Source, Receiver, RickerWavelet, MultiWavelet  # noqa: E402


@pytest.mark.parametrize("dimension,coords", [
    (2, [(0, 25)]), (2, [(0, 25), [250, 250]]), (3, [(0.5, 50.8, 500)])])
def test_source_attributes(dimension, coords):
    space = ones_model(dimension, extent=500)
    for cls in (Source, Receiver):
        src = cls(space, coordinates=coords, window_radius=8)
        assert src.space_model is space and src.window_radius == 8
        assert src.count == len(coords)
        assert np.array_equal(src.coordinates, space.dtype(coords))


GRID_POSITIONS = [   # bounding box, spacing, coordinates -> grid position (:38-63)
    ((0, 5120, 0, 5120), (10, 10), (0, 512), (0, 51.2)),
    ((0, 5120, 0, 5120), (10, 5), (120, 500), (12, 100)),
    ((0, 500, 20, 500, 0, 200), (20, 20, 20), (10, 20, 100), (0.5, 0.0, 5.0)),
]


@pytest.mark.parametrize("bbox,spacing,coords,expected", GRID_POSITIONS)
def test_source_grid_positions(bbox, spacing, coords, expected):
    space = SpaceModel(bounding_box=bbox, grid_spacing=spacing,
                       velocity_model=np.full((50,) * len(spacing), 1500.0,
                                              dtype=np.float32))
    src = Source(space, coordinates=coords, window_radius=4)
    assert np.array_equal(src.grid_positions,
                          np.asarray([expected], dtype=space.dtype))


ADJUSTED = [   # damping length, space order, coordinates -> position in the extended grid (:65-103)
    (2, 500, 2, (0, 0), (51, 51)), (2, 0, 2, (0, 0), (1, 1)),
    (2, 0, 4, (250, 100), (27, 12)), (2, 50, 16, (255, 255), (38.5, 38.5)),
    (3, (50, 0, 50, 0, 0, 0), 16, (255, 255, 0), (38.5, 38.5, 8)),
    (3, 2, 4, (250, 100, 100), (27, 12, 12)),
]


@pytest.mark.parametrize("dimension,damping,order,coords,expected", ADJUSTED)
def test_source_positions_in_the_extended_grid(dimension, damping, order,
                                               coords, expected):
    space = ones_model(dimension, extent=500, space_order=order)
    space.config_boundary(damping_length=damping)
    src = Source(space, coordinates=coords)
    assert np.array_equal(src.adjusted_grid_positions, space.dtype([expected]))


def test_source_errors_and_wavelets():
    space = ones_model(2, extent=500)
    with pytest.raises(ValueError):
        Source(space, coordinates="0,0")
    with pytest.raises(Exception, match="out of bounds"):
        Source(space, coordinates=[(0, 501)]).grid_positions
    time = TimeModel(space_model=space, tf=0.2)
    ricker = RickerWavelet(10.0, time)
    assert ricker.values.shape == (time.timesteps,) and ricker.num_sources == 1
    assert ricker.timesteps == time.timesteps
    multi = MultiWavelet(np.ones((time.timesteps, 3)), time)
    assert multi.num_sources == 3 and multi.values.flags["C_CONTIGUOUS"]
    with pytest.raises(ValueError):
        MultiWavelet(np.ones((time.timesteps + 1, 3)), time)


def test_every_environment_knob_is_documented():
    """Every SIMWAVE_* variable the sources read is described in
    INTEGRATION.md or in include/simwave_cuda.h."""
    import glob
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sources = glob.glob(os.path.join(root, "simwave_b200", "csrc", "*.c*")) + \
        glob.glob(os.path.join(root, "simwave_b200", "csrc", "*.h")) + \
        glob.glob(os.path.join(root, "simwave_b200", "**", "*.py"), recursive=True)
    names = set()
    for path in sources:
        with open(path) as f:
            names.update(re.findall(r"SIMWAVE_(?:CUDA|B200)_[A-Z0-9_]+", f.read()))
    names = {n for n in names if not n.endswith("_")}     # macro prefixes
    docs = ""
    for path in ("INTEGRATION.md", os.path.join("include", "simwave_cuda.h")):
        with open(os.path.join(root, path)) as f:
            docs += f.read()
    missing = sorted(n for n in names if n not in docs)
    assert not missing, missing
