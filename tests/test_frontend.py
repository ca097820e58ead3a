"""
Front-end parity: every array simwave_b200's front end hands to the kernel is
bit-identical to what the reference front end produced in this environment
(tests/golden/frontend.npz, made by make_golden.py from /root/reference).
North-star requirement: "source/receiver grid indices and interpolation
weights must be bit-exact".
"""
import numpy as np
import pytest

import simwave_b200 as api
import cases

FIELDS = ["shape", "extended_shape", "nbl", "grid_positions",
          "adjusted_grid_positions", "points", "values", "offsets",
          "damping_mask", "coeff2", "coeff1", "dt", "timesteps", "ricker"]


@pytest.mark.parametrize("name", sorted(cases.FRONTEND_CASES))
def test_frontend_tables_bit_exact(golden, name):
    ref = golden("frontend")
    out = cases.frontend_outputs(api, cases.FRONTEND_CASES[name])
    for field in FIELDS:
        expected = ref["{}/{}".format(name, field)]
        got = np.asarray(out[field])
        assert got.dtype == expected.dtype, (field, got.dtype, expected.dtype)
        assert got.shape == expected.shape, field
        assert np.array_equal(got, expected), field


@pytest.mark.parametrize("space_order", range(2, 21, 2))
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_fd_coefficients(golden, space_order, dtype):
    """Weights vs the reference front end (whose `findiff` dependency is
    replaced by an independent sympy solve in oracle/ref_stubs)."""
    ref = golden("frontend")
    vel = np.full((8, 8), 1500.0, dtype=dtype)
    sm = api.SpaceModel((0, 70, 0, 70), (10, 10), vel,
                        space_order=space_order, dtype=dtype)
    tag = "fd/so{}_{}".format(space_order, np.dtype(dtype).name)
    for deriv, key in ((2, "_c2"), (1, "_c1")):
        got = sm.fd_coefficients(deriv)
        assert got.dtype == dtype
        assert np.array_equal(got, ref[tag + key])


def test_fd_known_values():
    """Reference tests/test_space_model.py:133-155 (orders 2, 4, 8)."""
    vel = np.full((8, 8), 1500.0, dtype=np.float32)
    expect = {
        2: [-2.0, 1.0],
        4: [-2.5, 1.33333333e+00, -8.33333333e-02],
        8: [-2.84722222e+00, 1.60000000e+00, -2.00000000e-01,
            2.53968254e-02, -1.78571429e-03],
    }
    for so, coeffs in expect.items():
        sm = api.SpaceModel((0, 70, 0, 70), (10, 10), vel, space_order=so)
        assert np.allclose(sm.fd_coefficients(2), np.float32(coeffs))
    # appendix B.1 of SURVEY.md: first-derivative weights, order 8
    sm = api.SpaceModel((0, 70, 0, 70), (10, 10), vel, space_order=8)
    assert np.allclose(sm.fd_coefficients(1),
                       np.float32([0, 0.8, -0.2, 0.03809524, -0.0035714286]))


def test_kws_full_axis_matches_window():
    """The windowed evaluation used for the tables agrees bit-for-bit with
    the full-axis evaluation the reference performs (kws.py:44-135)."""
    from simwave_b200.kernel.frontend import kws
    rng = np.random.default_rng(7)
    for w in range(1, 11):
        for n in (5, 31, 400):
            for pos in np.float32(rng.uniform(0, n - 1, size=6)):
                full = kws.kaiser_windowing_sinc(n, pos, w)
                b0, e0, v0 = kws.get_kws_valid_points(full)
                b1, e1, v1 = kws.axis_window(n, pos, w)
                assert (b0, e0) == (b1, e1)
                assert np.array_equal(v0, v1)


def test_blockwise_interpolation_is_identical(monkeypatch):
    """SpaceModel.interpolate walks the target grid in blocks of planes
    (bounded float64 temporaries); any block size gives the same bits as one
    evaluation over the whole meshgrid (reference model.py:210-261)."""
    from scipy.interpolate import RegularGridInterpolator
    from simwave_b200 import SpaceModel
    rng = np.random.default_rng(0)
    for shape, bbox, spacing in (((23, 31), (0, 880, 0, 1200), (7.0, 9.0)),
                                 ((9, 11, 13), (0, 400, 0, 500, 0, 300), (11., 13., 9.))):
        vel = (1500 + 3000 * rng.random(shape)).astype(np.float32)
        monkeypatch.setattr(SpaceModel, "_interp_block_bytes", 1 << 12)
        small = SpaceModel(bounding_box=bbox, grid_spacing=spacing,
                           velocity_model=vel, space_order=4)
        monkeypatch.setattr(SpaceModel, "_interp_block_bytes", 1 << 30)
        whole = SpaceModel(bounding_box=bbox, grid_spacing=spacing,
                           velocity_model=vel, space_order=4)
        assert small.velocity_model.shape == whole.velocity_model.shape
        assert np.array_equal(small.velocity_model, whole.velocity_model)
        # and the one-shot evaluation the reference does
        n = len(shape)
        bounds = whole._axis_bounds()
        axes = [np.linspace(lo, hi, shape[i]) for i, (lo, hi) in enumerate(bounds)]
        tgt = [np.linspace(lo, hi, whole.shape[i]) for i, (lo, hi) in enumerate(bounds)]
        ref = RegularGridInterpolator(tuple(axes), vel)(
            tuple(np.meshgrid(*tgt, indexing="ij"))).astype(np.float32)
        assert np.array_equal(whole.velocity_model, ref)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("half_width", [1, 4, 10])
def test_batched_tables_equal_one_by_one(half_width, dtype):
    """The tables of a whole acquisition are built in one vectorised pass
    (kws.get_source_points_batch); they must equal the position-by-position
    construction bit for bit, windows clipped at the grid ends included."""
    from simwave_b200.kernel.frontend import kws
    rng = np.random.default_rng(half_width)
    for shape in ((57, 131), (23, 31, 29)):
        top = np.array(shape) - 1
        pos = (rng.random((64, len(shape))) * top).astype(dtype)
        pos[:4] = 0
        pos[4:8] = top
        pos[8:12] = np.round(pos[8:12])
        iv, weights, offsets = kws.get_source_points_batch(shape, pos, half_width)
        one = [kws.get_source_points(shape, [dtype(x) for x in p], half_width)
               for p in pos]
        assert np.array_equal(iv, np.concatenate([a for a, _ in one]))
        assert np.array_equal(weights, np.concatenate([b for _, b in one]))
        assert np.array_equal(np.diff(offsets.astype(np.int64)),
                              [b.size for _, b in one])
        assert iv.dtype == np.uint and weights.dtype == np.float32


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_batched_tables_equal_one_by_one_tables(dtype):
    """The batched table builder promotes `index - position` like the
    one-by-one path for either model precision (NumPy >= 2 scalar rules, the
    floor this package states in README.md)."""
    from simwave_b200.kernel.frontend import kws
    assert int(np.__version__.split(".")[0]) >= 2
    rng = np.random.default_rng(7)
    shape = (60, 75, 90)
    locations = (rng.random((40, 3)) * (np.array(shape) - 1)).astype(dtype)
    iv, values, offsets = kws.get_source_points_batch(shape, locations, 4)
    for i, location in enumerate(locations):
        p, v = kws.get_source_points(shape, [dtype(x) for x in location], 4)
        assert np.array_equal(iv[6 * i:6 * i + 6], p)
        assert np.array_equal(values[int(offsets[i]):int(offsets[i + 1])], v)


def test_extended_arrays_are_kept_until_the_boundary_configuration_changes():
    vel = np.linspace(1500, 3000, 30 * 40, dtype=np.float32).reshape(30, 40)
    sm = api.SpaceModel((0, 290, 0, 390), (10, 10), vel, density_model=vel / 2,
                        space_order=4)
    sm.config_boundary(damping_length=(0, 30, 20, 20),
                       boundary_condition="null_dirichlet")
    v1, d1, m1, t1 = (sm.extended_velocity_model, sm.extended_density_model,
                      sm.damping_mask, sm.model_token)
    assert sm.extended_velocity_model is v1 and sm.damping_mask is m1
    assert sm.extended_density_model is d1 and sm.model_token == t1 != 0
    for a in (v1, d1, m1, sm.velocity_model, sm.density_model):
        assert not a.flags.writeable
        with pytest.raises(ValueError):
            a[0, 0] = 1
    # same values as a fresh build
    fresh = api.SpaceModel((0, 290, 0, 390), (10, 10), vel, density_model=vel / 2,
                           space_order=4)
    fresh.config_boundary(damping_length=(0, 30, 20, 20),
                          boundary_condition="null_dirichlet")
    assert np.array_equal(fresh.extended_velocity_model, v1)
    assert np.array_equal(fresh.damping_mask, m1)
    assert fresh.model_token != t1
    # a new configuration rebuilds them under a new token
    sm.config_boundary(damping_length=40, boundary_condition="none")
    assert sm.extended_velocity_model is not v1
    assert sm.extended_velocity_model.shape == sm.extended_shape
    assert sm.model_token != t1


def test_middleware_withdraws_its_hints_after_the_call():
    """Solver.forward's data-path hints travel through
    simwave_cuda_set_hint around the call only."""
    from simwave_b200.kernel.backend.middleware import Middleware

    class Setter:            # stands in for the ctypes function object
        def __init__(self):
            self.calls = []

        def __call__(self, key, value):
            self.calls.append((key, value))
            return 0

    class FakeLib:
        simwave_cuda_set_hint = Setter()

    lib = FakeLib()
    done = Middleware._apply_hints(lib, {'wavefield_in_zero': 1, 'model_resident': 42})
    assert done == ['wavefield_in_zero', 'model_resident']
    Middleware._apply_hints(lib, dict.fromkeys(done, 0))
    assert lib.simwave_cuda_set_hint.calls == [(1, 1), (3, 42), (1, 0), (3, 0)]
    assert Middleware._apply_hints(object(), {'wavefield_out': 1}) == []


@pytest.mark.parametrize("seed", range(12))
def test_damping_mask_equals_the_reference_procedure(seed):
    """SpaceModel.damping_mask gathers from the mask of a one-cell domain (the
    ramp values only depend on the depth inside the layers); it must equal,
    bit for bit, what the reference computes over the whole grid
    (model.py:378-406: np.pad linear_ramp, power, alpha, halo padding)."""
    rng = np.random.default_rng(seed)
    dim = 2 + seed % 2
    shape = tuple(int(x) for x in rng.integers(3, 28, size=dim))
    dtype = (np.float32, np.float64)[seed % 2 if seed > 1 else seed]
    h = tuple(float(x) for x in rng.uniform(5, 20, size=dim))
    box = []
    for n, hh in zip(shape, h):
        box += [0.0, (n - 1) * hh]
    sm = api.SpaceModel(bounding_box=tuple(box), grid_spacing=h,
                        velocity_model=rng.uniform(1500, 3000, size=shape).astype(dtype),
                        space_order=int(rng.choice([2, 4, 8, 16])), dtype=dtype)
    lengths = tuple(float(rng.integers(0, 12)) * hh for hh in h for _ in range(2))
    sm.config_boundary(damping_length=lengths,
                       boundary_condition=("null_neumann",) * (2 * dim),
                       damping_polynomial_degree=int(rng.integers(1, 5)),
                       damping_alpha=float(rng.uniform(1e-4, 1e-2)))
    want = np.pad(array=np.zeros(sm.shape, dtype=sm.dtype), pad_width=sm.nbl_pad_width,
                  mode="linear_ramp", end_values=sm.nbl_pad_width)
    want = (want ** sm.damping_polynomial_degree) * sm.damping_alpha
    want = np.pad(array=want, pad_width=sm.halo_pad_width)
    got = sm.damping_mask
    assert got.dtype == want.dtype and got.shape == want.shape
    assert np.array_equal(got, want)
