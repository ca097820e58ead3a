"""
GPU parity tests proper: the CUDA backend, called through the drop-in C ABI
(`forward` of the shim libraries), against the CPU oracle on identical seeded
inputs, and against the golden vectors of the unmodified reference.

Tolerances (float32, from BASELINE.md / SURVEY.md section 8d):
  * default ("fast") math mode: relative L2 <= 1e-5 for runs of <= 500 steps
    (the north star's stated float32 tolerance); float64 <= 1e-12;
  * strict math mode (SIMWAVE_CUDA_MATH=strict): the arithmetic is rounded
    exactly where the reference rounds it, so wavefield and receivers are
    required to be BIT-IDENTICAL to the sequential C oracle.
Most tests below run in strict mode (autouse fixture) because bit identity is
the sharpest check of the loop structure, boundary logic, source ordering and
snapshot plumbing; the *_default_math tests repeat the sweep in the default
mode against the tolerance.
"""
import ctypes
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "oracle"))

import oracle  # noqa: E402
import problems  # noqa: E402
import cases  # noqa: E402
import simwave_b200 as api  # noqa: E402
from cuda_abi import cuda_forward, core  # noqa: E402
from conftest import rel_l2  # noqa: E402

pytestmark = pytest.mark.gpu

REL_L2_TOL = 1e-5   # float32, <= 500 steps (north star / BASELINE.md section 3)
REL_L2_TOL_F64 = 1e-12


@pytest.fixture(autouse=True)
def strict_math(monkeypatch):
    monkeypatch.setenv("SIMWAVE_CUDA_MATH", "strict")


def assert_close(a, b, dtype=np.float32):
    tol = REL_L2_TOL if np.dtype(dtype) == np.float32 else REL_L2_TOL_F64
    assert np.abs(a["u"]).max() > 0
    eu, er = rel_l2(b["u"], a["u"]), rel_l2(b["receivers"], a["receivers"])
    assert eu <= tol and er <= tol, (eu, er)


def run_pair(p, env=None, monkeypatch=None):
    a, b = problems.clone(p), problems.clone(p)
    oracle.forward(a)
    if env:
        for k, v in env.items():
            monkeypatch.setenv(k, v)
    cuda_forward(b)
    return a, b


def assert_identical(a, b):
    assert np.abs(a["u"]).max() > 0
    assert np.array_equal(a["receivers"], b["receivers"]), \
        "receivers differ: rel-L2 %.3e" % rel_l2(b["receivers"], a["receivers"])
    assert np.array_equal(a["u"], b["u"]), \
        "wavefield differs: rel-L2 %.3e" % rel_l2(b["u"], a["u"])


# shape, order, density, dtype, stride, steps, bc
ABI_CASES = [
    ((40, 52), 2, False, np.float32, 0, 30, (2, 1, 0, 1)),
    ((40, 52), 8, False, np.float32, 1, 30, (1, 2, 2, 0)),
    ((44, 36), 4, True, np.float32, 2, 31, (2, 2, 2, 2)),
    ((64, 70), 20, True, np.float64, 3, 31, (0, 1, 2, 1)),
    ((40, 52), 6, False, np.float64, 4, 29, (1, 1, 1, 1)),
    ((45, 50), 10, True, np.float32, 0, 40, (0, 0, 0, 0)),
    ((20, 22, 24), 2, False, np.float32, 0, 20, (2, 1, 0, 1, 2, 1)),
    ((30, 32, 34), 8, False, np.float32, 5, 21, (1, 2, 1, 2, 1, 2)),
    ((24, 22, 22), 4, True, np.float32, 0, 20, (2, 2, 2, 2, 2, 2)),
    ((22, 26, 20), 4, True, np.float32, 2, 21, (2, 1, 0, 1, 2, 1)),   # nx != ny
    ((22, 20, 27), 6, True, np.float64, 1, 15, (0, 1, 2, 0, 1, 2)),   # nx != ny
    ((36, 36, 36), 12, True, np.float64, 0, 12, (2, 1, 2, 1, 2, 1)),
    ((40, 38, 42), 16, False, np.float32, 0, 16, (2, 1, 0, 1, 0, 1)),
    ((64, 64, 64), 20, False, np.float32, 0, 10, (2, 1, 1, 1, 1, 1)),
    ((33, 35, 37), 6, False, np.float64, 0, 18, (1, 0, 2, 2, 0, 1)),
]


@pytest.mark.parametrize("shape,order,density,dtype,stride,steps,bc", ABI_CASES)
def test_strict_mode_bit_identical_to_oracle(shape, order, density, dtype,
                                             stride, steps, bc):
    ndim = len(shape)
    r = order // 2
    nbl = tuple((0, 3) if a == 0 else (2, 4) for a in range(ndim))
    p = problems.make_problem(
        shape=shape, space_order=order, density=density, dtype=dtype,
        timesteps=steps, saving_stride=stride, nbl=nbl, bc=bc,
        num_sources=3, num_receivers=9, src_radius=min(4, r + 1),
        rec_radius=2, multi_wavelet=(stride % 2 == 0), seed=order + stride)
    a, b = run_pair(p)
    assert_identical(a, b)


@pytest.mark.parametrize("shape,order,density,dtype,stride,steps,bc", ABI_CASES)
def test_default_math_within_tolerance(shape, order, density, dtype, stride,
                                       steps, bc, monkeypatch):
    ndim = len(shape)
    r = order // 2
    nbl = tuple((0, 3) if a == 0 else (2, 4) for a in range(ndim))
    p = problems.make_problem(
        shape=shape, space_order=order, density=density, dtype=dtype,
        timesteps=steps, saving_stride=stride, nbl=nbl, bc=bc,
        num_sources=3, num_receivers=9, src_radius=min(4, r + 1),
        rec_radius=2, multi_wavelet=(stride % 2 == 0), seed=order + stride)
    monkeypatch.delenv("SIMWAVE_CUDA_MATH")
    a, b = run_pair(p)
    assert_close(a, b, dtype)


@pytest.mark.parametrize("ndim", [2, 3])
def test_sources_in_the_halo_and_on_boundary_planes(ndim):
    """Windows that reach into the halo and sit on Dirichlet / Neumann
    planes: the fused boundary logic must treat the increment exactly as the
    reference's add-then-boundary sequence does."""
    shape = (40, 44) if ndim == 2 else (30, 34, 32)
    r = 4
    for bc in [(2, 1) * ndim, (1, 2) * ndim, (0, 0) * ndim, (2, 2) * ndim]:
        hi = [n - 1 for n in shape]
        src = [[r + 0.3] * ndim, [h - r - 0.6 for h in hi],
               [r + 1.5] + [shape[i] / 2 for i in range(1, ndim)],
               [shape[0] / 2] * (ndim - 1) + [hi[-1] - r - 1.2]]
        p = problems.make_problem(
            shape=shape, space_order=8, timesteps=25, bc=bc, seed=5,
            src_positions=np.array(src), src_radius=4, num_receivers=10,
            rec_positions=np.array(src + [[r + 0.1] * ndim]), rec_radius=4,
            multi_wavelet=True)
        a, b = run_pair(p)
        assert_identical(a, b)


def test_overlapping_sources_keep_sequential_order():
    """reference tests/test_parallel_solution.py geometry: nine sources whose
    radius-8 windows overlap; sequential C adds them in index order."""
    src = np.array([[12.0, 40.0, 13.0 + 6.4 * i] for i in range(9)])
    p = problems.make_problem(shape=(48, 80, 84), space_order=4, timesteps=20,
                              src_positions=src, src_radius=8,
                              num_receivers=20, rec_radius=8, seed=2)
    a, b = run_pair(p)
    assert_identical(a, b)


@pytest.mark.parametrize("ndim", [2, 3])
def test_fused_and_separate_boundaries_agree(ndim, monkeypatch):
    shape = (50, 46) if ndim == 2 else (30, 28, 32)
    p = problems.make_problem(shape=shape, space_order=6, timesteps=20,
                              bc=(2, 1, 1, 2, 2, 2)[:2 * ndim], seed=9,
                              num_sources=3, src_radius=3)
    fused = problems.clone(p)
    cuda_forward(fused)
    monkeypatch.setenv("SIMWAVE_CUDA_BC", "separate")
    separate = problems.clone(p)
    cuda_forward(separate)
    assert np.array_equal(fused["u"], separate["u"])
    assert np.array_equal(fused["receivers"], separate["receivers"])


def test_tiny_grid_takes_separate_boundary_path():
    """extent < 3r+2: mirror sources are no longer interior cells."""
    p = problems.make_problem(shape=(13, 14), space_order=8, timesteps=12,
                              bc=(2, 2, 2, 1), num_sources=1, src_radius=1,
                              num_receivers=3, rec_radius=1, seed=4)
    a, b = run_pair(p)
    assert_identical(a, b)


@pytest.mark.parametrize("shape,order,density,steps", [
    ((60, 64), 8, False, 60), ((40, 44, 48), 8, False, 60),
    ((36, 36, 36), 4, True, 60), ((120, 130), 8, False, 500),
    ((70, 150, 140), 8, False, 300), ((48, 50, 52), 16, True, 200)])
def test_fast_math_within_tolerance(shape, order, density, steps, monkeypatch):
    nbl = ((0, 6),) + ((5, 5),) * (len(shape) - 1)
    p = problems.make_problem(shape=shape, space_order=order, density=density,
                              timesteps=steps, seed=3, smooth_density=True,
                              nbl=nbl)
    a, b = run_pair(p, {"SIMWAVE_CUDA_MATH": "fast"}, monkeypatch)
    assert_close(a, b)


def as_float64(p):
    """The same problem with every float32 array widened (identical inputs)."""
    q = problems.clone(p)
    for k, v in list(q.items()):
        if isinstance(v, np.ndarray) and v.dtype == np.float32:
            q[k] = v.astype(np.float64)
        elif isinstance(v, np.float32):
            q[k] = np.float64(v)
    return q


@pytest.mark.parametrize("shape,order,steps", [((200, 210), 8, 1500),
                                               ((56, 60, 64), 8, 900)])
def test_fast_math_long_run_stays_at_the_float32_noise_floor(shape, order, steps,
                                                             monkeypatch):
    """Long loops (per-point random velocity: the worst case for rounding
    noise).  Two float32 evaluations of the same recursion drift apart like the
    sum of their rounding noises, so the distance to the reference's float32
    kernel alone says little; measured against the reference's float64 run on
    the same inputs, FAST mode must be as accurate as the reference's own
    float32 kernel (within 25 %), and stay under the stated 1e-4 of the
    float32 kernel."""
    nbl = ((0, 6),) + ((5, 5),) * (len(shape) - 1)
    p = problems.make_problem(shape=shape, space_order=order, timesteps=steps,
                              seed=3, nbl=nbl)
    ref32, ref64 = problems.clone(p), as_float64(p)
    oracle.forward(ref32)
    oracle.forward(ref64)
    monkeypatch.setenv("SIMWAVE_CUDA_MATH", "fast")
    b = problems.clone(p)
    cuda_forward(b)
    for key in ("u", "receivers"):
        floor = rel_l2(ref32[key], ref64[key])
        to_truth = rel_l2(b[key], ref64[key])
        to_ref32 = rel_l2(b[key], ref32[key])
        assert to_truth <= 1.25 * floor, (key, to_truth, floor)
        assert to_ref32 <= 1e-4, (key, to_ref32)


def test_nonzero_initial_fields_are_honoured():
    """`u` is in/out: slots 0..2 may carry an initial condition."""
    p = problems.make_problem(shape=(36, 40), space_order=4, timesteps=15,
                              seed=8)
    rng = np.random.default_rng(1)
    p["u"][0] = rng.standard_normal(p["u"][0].shape).astype(np.float32) * 1e-3
    p["u"][1] = p["u"][0] * 0.5
    p["u"][2, :2, :] = 0.25      # halo cells of the third slot
    a, b = run_pair(p)
    assert_identical(a, b)


def test_timestep_window_and_error_reporting():
    p = problems.make_problem(shape=(30, 30), space_order=4, timesteps=12)
    a, b = problems.clone(p), problems.clone(p)
    for q in (a, b):
        q["begin_timestep"], q["end_timestep"] = 3, 9
    oracle.forward(a)
    cuda_forward(b)
    assert_identical(a, b)
    bad = problems.clone(p)
    bad["bc"][0] = 7
    with pytest.raises(RuntimeError, match="boundary condition"):
        cuda_forward(bad)
    assert core().simwave_cuda_last_launch_count() > 0


# ---------------------------------------------------------------------------
# through the public API, against the golden vectors of the reference
# ---------------------------------------------------------------------------
CUDA = dict(language="cuda")


@pytest.mark.parametrize("name", sorted(cases.SMALL_CASES))
def test_small_cases_match_reference_golden(golden, name):
    ref = golden("forward_small")
    solver = cases.small_solver(api, name, api.Compiler(**CUDA))
    u, recv = solver.forward()
    assert np.array_equal(recv, ref[name + "/recv"])
    assert np.array_equal(u[ref[name + "/u_idx"]], ref[name + "/u"])


@pytest.mark.parametrize("dimension,space_order,density", [
    (2, 2, False), (2, 8, False), (3, 2, False), (3, 8, False),
    (2, 2, True), (3, 2, True)])
def test_solution_default_math(golden, dimension, space_order, density,
                               monkeypatch):
    """Reference tests/test_solution.py with language='cuda' in the default
    math mode: the reference's own bar for its GPU path, np.allclose(atol=1e-4)
    (tests/test_gpu_solution.py:139), plus the north star's relative-L2
    tolerance, which is the tighter of the two here."""
    monkeypatch.delenv("SIMWAVE_CUDA_MATH")
    ref = golden("solution_%dd_so%d" % (dimension, space_order))
    solver = cases.solution_solver(api, dimension, space_order,
                                   api.Compiler(**CUDA), density=density)
    u, recv = solver.forward()
    if dimension == 2:
        assert np.allclose(u, ref["u_reference_npy"], atol=1e-4)
        assert np.allclose(u, ref["u"], atol=1e-4)
        assert rel_l2(u, ref["u"]) <= REL_L2_TOL
        if not density:
            assert rel_l2(recv, ref["recv"]) <= REL_L2_TOL
    else:
        f = u[0]
        c = [n // 2 for n in f.shape]
        planes = (f[c[0]], f[:, c[1]], f[:, :, c[2]])
        for got, key in zip(planes, ("plane_z", "plane_x", "plane_y")):
            assert np.allclose(got, ref[key], atol=1e-4)
            assert rel_l2(got, ref[key]) <= REL_L2_TOL


@pytest.mark.parametrize("dimension,space_order,density", [
    (2, 2, False), (2, 8, False), (3, 2, False), (3, 8, False),
    (2, 2, True), (3, 2, True)])
def test_solution(golden, dimension, space_order, density):
    """Reference tests/test_solution.py with language='cuda': the reference's
    own .npy within its GPU bar atol=1e-4 (tests/test_gpu_solution.py:139)
    and its CPU bar atol=1e-5; the regenerated field bit for bit."""
    ref = golden("solution_%dd_so%d" % (dimension, space_order))
    solver = cases.solution_solver(api, dimension, space_order,
                                   api.Compiler(**CUDA), density=density)
    u, recv = solver.forward()
    if dimension == 2:
        assert np.allclose(u, ref["u_reference_npy"], atol=1e-5)
        if not density:
            assert np.array_equal(u, ref["u"])
            assert np.array_equal(recv, ref["recv"])
        else:
            assert np.allclose(u, ref["u"], atol=1e-5)
    else:
        f = u[0]
        c = [n // 2 for n in f.shape]
        planes = (f[c[0]], f[:, c[1]], f[:, :, c[2]])
        for got, key in zip(planes, ("plane_z", "plane_x", "plane_y")):
            if density:
                assert np.allclose(got, ref[key], atol=1e-5)
            else:
                assert np.array_equal(got, ref[key])


@pytest.mark.parametrize("dimension,density", [(2, False), (3, False),
                                               (2, True), (3, True)])
def test_parallel_solution_f64(golden, dimension, density):
    """Reference tests/test_parallel_solution.py (float64, atol=1e-8)."""
    ref = golden("parallel_f64")
    tag = "{}d_{}".format(dimension, "var" if density else "const")
    solver = cases.parallel_solver(api, dimension, density, np.float64,
                                   api.Compiler(**CUDA))
    u, recv = solver.forward()
    assert np.allclose(recv, ref[tag + "/recv"], atol=1e-8)
    if dimension == 2:
        assert np.allclose(u, ref[tag + "/u"], atol=1e-8)
    else:
        f = u[0]
        c = [n // 2 for n in f.shape]
        assert np.allclose(f[c[0]], ref[tag + "/plane_z"], atol=1e-8)
        assert np.allclose(f[:, :, c[2]], ref[tag + "/plane_y"], atol=1e-8)


@pytest.mark.parametrize("dimension,density,stride", [
    (2, False, 1), (2, False, 2), (2, False, 5), (2, True, 2),
    (3, False, 1), (3, False, 2), (3, False, 3), (3, False, 4), (3, True, 5)])
def test_u_saving(dimension, density, stride):
    """Reference tests/test_u_saving.py: the last saved snapshot is
    bit-identical whatever the saving stride."""
    comp = api.Compiler(**CUDA)
    base, _ = cases.u_saving_solver(api, dimension, density, 0, comp).forward()
    last, _ = cases.u_saving_solver(api, dimension, density, stride,
                                    comp).forward()
    assert np.array_equal(base[-1], last[-1])


# ---------------------------------------------------------------------------
# tiled 3D kernel: every radius and tile configuration against the plain kernel
# ---------------------------------------------------------------------------
def _tiled_vs_plain(order, tile, monkeypatch, shape=(37, 75, 150), steps=6,
                    bc=(2, 1, 2, 1, 2, 1), math="strict", density=False,
                    dtype=np.float32):
    r = order // 2
    p = problems.make_problem(
        shape=shape, space_order=order, timesteps=steps, bc=bc, seed=order,
        density=density, dtype=dtype,
        nbl=((0, 3), (2, 2), (3, 2)), num_sources=2, src_radius=min(4, r + 1),
        num_receivers=6, rec_radius=2)
    monkeypatch.setenv("SIMWAVE_CUDA_MATH", math)
    monkeypatch.setenv("SIMWAVE_CUDA_KERNEL", "simple")
    plain = problems.clone(p)
    cuda_forward(plain)
    monkeypatch.setenv("SIMWAVE_CUDA_KERNEL", "auto")
    monkeypatch.setenv("SIMWAVE_CUDA_TILE", tile)
    tiled = problems.clone(p)
    cuda_forward(tiled)
    assert np.abs(plain["u"]).max() > 0
    return plain, tiled


@pytest.mark.parametrize("math", ["strict", "fast"])
@pytest.mark.parametrize("order", [2, 4, 6, 8, 10, 12, 14, 16, 18, 20])
def test_tiled_kernel_every_radius(order, math, monkeypatch):
    for cfg in (0, 1, 2, 3):
        plain, tiled = _tiled_vs_plain(order, "%d:11" % cfg, monkeypatch,
                                       math=math)
        assert np.array_equal(plain["u"], tiled["u"]), (order, cfg)
        assert np.array_equal(plain["receivers"], tiled["receivers"])


@pytest.mark.parametrize("math", ["strict", "fast"])
@pytest.mark.parametrize("order", [2, 4, 8, 10, 12, 16, 20])
def test_tiled_variable_density_every_radius(order, math, monkeypatch):
    """Variable density through the tiled kernel (streamed density
    derivatives) against the plain kernel, nx == ny."""
    for cfg in (0, 1, 2, 3):
        plain, tiled = _tiled_vs_plain(order, "%d:9" % cfg, monkeypatch,
                                       shape=(40, 150, 150), math=math,
                                       density=True)
        assert np.array_equal(plain["u"], tiled["u"]), (order, cfg)
        assert np.array_equal(plain["receivers"], tiled["receivers"])


@pytest.mark.parametrize("math", ["strict", "fast"])
@pytest.mark.parametrize("order", [2, 4, 6, 8, 10, 12, 14, 16, 18, 20])
def test_tiled_float64_kernel_every_radius(order, math, monkeypatch):
    """float64 3D (what simwave's own overthrust benchmark script runs): the
    tiled kernel of sw_step_tiled3d64.cuh against the plain kernel, bit for
    bit, on a grid whose extents are no multiples of the tile, with damping
    layers, mixed boundary conditions and several z-chunk lengths."""
    for zchunk in (0, 7, 1000):
        plain, tiled = _tiled_vs_plain(order, "0:%d" % zchunk, monkeypatch,
                                       shape=(45, 61, 107), math=math,
                                       dtype=np.float64)
        assert np.array_equal(plain["u"], tiled["u"]), (order, zchunk)
        assert np.array_equal(plain["receivers"], tiled["receivers"])


@pytest.mark.parametrize("math", ["strict", "fast"])
@pytest.mark.parametrize("order", [2, 4, 8, 12, 16, 20])
def test_tiled_float64_variable_density_every_radius(order, math, monkeypatch):
    """float64 3D with a density model (streamed density derivatives, as in
    the float32 kernel) against the plain kernel, bit for bit, nx == ny."""
    for zchunk in (0, 9):
        plain, tiled = _tiled_vs_plain(order, "0:%d" % zchunk, monkeypatch,
                                       shape=(45, 88, 88), math=math,
                                       density=True, dtype=np.float64)
        assert np.array_equal(plain["u"], tiled["u"]), (order, zchunk)
        assert np.array_equal(plain["receivers"], tiled["receivers"])


def test_tiled_float64_variable_density_matches_oracle(monkeypatch):
    p = problems.make_problem(shape=(46, 72, 72), space_order=8, density=True,
                              timesteps=25, seed=14, smooth_density=True,
                              dtype=np.float64, nbl=((0, 5), (4, 4), (4, 4)))
    a, b = run_pair(p, {"SIMWAVE_CUDA_MATH": "strict"}, monkeypatch)
    assert_identical(a, b)


@pytest.mark.parametrize("bc", [(2, 2, 2, 2, 2, 2), (1, 1, 1, 1, 1, 1),
                                (0, 2, 1, 0, 2, 1), (0, 0, 0, 0, 0, 0)])
def test_tiled_float64_kernel_boundaries_and_oracle(bc, monkeypatch):
    """float64 3D through the tiled kernel against the CPU oracle (strict: bit
    for bit), every boundary mix, fused and separate boundary passes."""
    p = problems.make_problem(shape=(44, 70, 99), space_order=8, timesteps=25,
                              bc=bc, seed=21, dtype=np.float64,
                              nbl=((0, 5), (4, 4), (4, 3)))
    a, b = run_pair(p, {"SIMWAVE_CUDA_MATH": "strict"}, monkeypatch)
    assert_identical(a, b)
    monkeypatch.setenv("SIMWAVE_CUDA_BC", "separate")
    c = problems.clone(p)
    cuda_forward(c)
    assert np.array_equal(a["u"], c["u"])


def test_tiled_variable_density_matches_oracle():
    p = problems.make_problem(shape=(50, 90, 90), space_order=8, density=True,
                              timesteps=30, seed=12, smooth_density=True,
                              nbl=((0, 5), (4, 4), (4, 4)))
    a, b = run_pair(p)
    assert_identical(a, b)


@pytest.mark.parametrize("cfg", range(8))
@pytest.mark.parametrize("bc", [(2, 2, 2, 2, 2, 2), (1, 1, 1, 1, 1, 1),
                                (0, 2, 1, 0, 2, 1)])
def test_tiled_kernel_every_configuration(cfg, bc, monkeypatch):
    for zchunk in (0, 5, 1000):
        plain, tiled = _tiled_vs_plain(8, "%d:%d" % (cfg, zchunk), monkeypatch,
                                       shape=(41, 70, 203), bc=bc)
        assert np.array_equal(plain["u"], tiled["u"]), (cfg, zchunk)


def test_tiled_kernel_fast_math_matches_plain_fast_math(monkeypatch):
    for cfg in range(8):
        plain, tiled = _tiled_vs_plain(8, "%d:0" % cfg, monkeypatch,
                                       math="fast", steps=12)
        assert np.array_equal(tiled["u"], plain["u"]), cfg


def test_tiled_kernel_in_place_between_snapshots(monkeypatch):
    """saving_stride > 1: the reference updates in place (next == prev)."""
    p = problems.make_problem(shape=(30, 50, 90), space_order=8, timesteps=21,
                              saving_stride=4, seed=6)
    a, b = run_pair(p)
    assert_identical(a, b)


# ---------------------------------------------------------------------------
# persistent 2D time-loop kernel against the per-step launches
# ---------------------------------------------------------------------------
LOOP_PER_STEP, LOOP_GRID, LOOP_RESIDENT = 0, 1, 2


def loop_kind():
    lib = core()
    lib.simwave_cuda_last_loop_kind.restype = ctypes.c_int
    return lib.simwave_cuda_last_loop_kind()


@pytest.fixture(params=["resident", "grid"])
def loop2d(request, monkeypatch):
    """Both persistent 2D loops: the grid-barrier one (default) and the
    tile-resident one (opt-in, falls back to the other where it does not fit)."""
    if request.param == "resident":
        monkeypatch.setenv("SIMWAVE_CUDA_LOOP2D", "resident")
    return request.param


@pytest.mark.parametrize("math", ["strict", "fast"])
@pytest.mark.parametrize("shape,order,density,dtype,bc", [
    ((90, 300), 8, False, np.float32, (2, 1, 1, 1)),
    ((64, 70), 4, True, np.float32, (2, 2, 2, 2)),
    ((70, 66), 20, True, np.float64, (0, 1, 2, 1)),
    ((13, 14), 8, False, np.float32, (2, 2, 2, 1)),      # separate boundary passes
    ((300, 700), 2, False, np.float64, (1, 0, 0, 2))])
def test_persistent_2d_loop_matches_per_step_launches(shape, order, density,
                                                      dtype, bc, math, loop2d,
                                                      monkeypatch):
    """One cooperative launch for the whole time loop (sw_loop2d.cuh,
    sw_loop2d_resident.cuh) must reproduce the three-kernels-per-step path bit
    for bit: wavefield slots and traces, many overlapping sources and
    per-source wavelets included."""
    small = min(shape) < 20
    p = problems.make_problem(
        shape=shape, space_order=order, density=density, dtype=dtype,
        timesteps=45, bc=bc, seed=order, multi_wavelet=True,
        nbl=((0, 0), (0, 0)) if small else ((0, 5), (4, 6)),
        num_sources=1 if small else 5, src_radius=1 if small else 4,
        num_receivers=3 if small else 40, rec_radius=1 if small else 3)
    monkeypatch.setenv("SIMWAVE_CUDA_MATH", math)
    monkeypatch.setenv("SIMWAVE_CUDA_LOOP", "launch")
    per_step = problems.clone(p)
    cuda_forward(per_step)
    launches_per_step = core().simwave_cuda_last_launch_count()
    assert loop_kind() == LOOP_PER_STEP
    monkeypatch.delenv("SIMWAVE_CUDA_LOOP")
    persistent = problems.clone(p)
    cuda_forward(persistent)
    launches_persistent = core().simwave_cuda_last_launch_count()
    # many overlapping sources are added by a separate phase, which only the
    # grid-barrier loop has; grids below 3r+2 take separate boundary passes
    assert loop_kind() in (LOOP_GRID, LOOP_RESIDENT)
    if loop2d == "grid":
        assert loop_kind() == LOOP_GRID
    assert np.abs(per_step["u"]).max() > 0
    assert np.array_equal(per_step["u"], persistent["u"])
    assert np.array_equal(per_step["receivers"], persistent["receivers"])
    assert launches_persistent < launches_per_step / 10


def test_persistent_2d_loop_timestep_windows(loop2d):
    """Two consecutive windows of the loop through the plan API equal one run."""
    from simwave_b200 import slab
    p = problems.make_problem(shape=(80, 90), space_order=6, timesteps=31, seed=2)
    whole = problems.clone(p)
    cuda_forward(whole)
    q = problems.clone(p)
    plan = slab.Plan(q)
    plan.run(1, 10)
    plan.run(11, 31)
    plan.download()
    plan.destroy()
    assert np.array_equal(whole["u"], q["u"])
    assert np.array_equal(whole["receivers"], q["receivers"])


# ---------------------------------------------------------------------------
# host data path: pageable and page-locked callers see the same results
# ---------------------------------------------------------------------------
def test_pinned_and_pageable_host_buffers_agree():
    import torch
    p = problems.make_problem(shape=(40, 150, 160), space_order=8, timesteps=10,
                              seed=6, nbl=((0, 4), (3, 3), (3, 3)))
    rng = np.random.default_rng(5)
    p["u"][1, 8:20] = 1e-3 * rng.standard_normal(p["u"][1, 8:20].shape)
    pageable = problems.clone(p)
    cuda_forward(pageable)
    pinned = problems.clone(p)
    keep = []
    for key in ("u", "velocity", "damp", "receivers"):
        t = torch.from_numpy(pinned[key]).pin_memory()
        keep.append(t)
        pinned[key] = t.numpy()
    cuda_forward(pinned)
    core().simwave_cuda_release_cache()
    assert np.array_equal(pageable["u"], pinned["u"])
    assert np.array_equal(pageable["receivers"], pinned["receivers"])


def test_allocation_cache_keeps_the_last_working_set_only():
    """A survey over changing shapes must not pile up dead device blocks: after
    a forward() the cache holds what that call used, nothing an earlier,
    larger problem left behind."""
    lib = core()
    lib.simwave_cuda_cached_bytes.restype = ctypes.c_ulonglong
    lib.simwave_cuda_release_cache()
    big = problems.make_problem(shape=(60, 200, 210), space_order=8, timesteps=3, seed=1)
    small = problems.make_problem(shape=(30, 64, 70), space_order=8, timesteps=3, seed=1)
    cuda_forward(big)
    after_big = lib.simwave_cuda_cached_bytes()
    assert after_big >= 5 * big["velocity"].nbytes        # fields of the big problem are kept
    cuda_forward(small)
    after_small = lib.simwave_cuda_cached_bytes()
    assert 0 < after_small < big["velocity"].nbytes        # ... and gone after a smaller call
    cuda_forward(small)
    assert lib.simwave_cuda_cached_bytes() == after_small  # steady state of a survey
    lib.simwave_cuda_release_cache()
    assert lib.simwave_cuda_cached_bytes() == 0


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("density", [False, True])
@pytest.mark.parametrize("order", [2, 4, 6, 8, 10, 12, 14, 16, 20])
def test_persistent_2d_loop_every_radius(order, density, dtype, loop2d, monkeypatch):
    """Every strip shape of the persistent 2D kernels (rows per thread depend on
    radius, precision and density) against the per-step launches, bit for bit,
    on a grid whose extents are not multiples of the tile."""
    p = problems.make_problem(
        shape=(61 + 2 * order, 167 + order), space_order=order, density=density,
        dtype=dtype, timesteps=9, bc=(2, 1, 1, 2), seed=order + 1,
        nbl=((0, 4), (3, 5)), num_sources=2, src_radius=2, num_receivers=5,
        rec_radius=2)
    monkeypatch.setenv("SIMWAVE_CUDA_LOOP", "launch")
    per_step = problems.clone(p)
    cuda_forward(per_step)
    monkeypatch.delenv("SIMWAVE_CUDA_LOOP")
    persistent = problems.clone(p)
    cuda_forward(persistent)
    assert loop_kind() == (LOOP_GRID if loop2d == "grid" else LOOP_RESIDENT)
    assert np.abs(per_step["u"]).max() > 0
    assert np.array_equal(per_step["u"], persistent["u"])
    assert np.array_equal(per_step["receivers"], persistent["receivers"])


@pytest.mark.parametrize("math", ["strict", "fast"])
@pytest.mark.parametrize("shape,order,bc,nbl", [
    ((421, 1841), 8, (2, 1, 1, 1), ((0, 70), (70, 70))),     # Marmousi-sized: ~145 tiles
    ((517, 517), 4, (2, 1, 0, 1), ((0, 0), (0, 0))),         # README-sized
    ((333, 1203), 2, (2, 2, 2, 2), ((0, 9), (7, 8)))])
def test_resident_2d_loop_many_tiles(shape, order, bc, nbl, math, monkeypatch):
    """The tile-resident loop on grids that fill the whole device with tiles:
    halo strips from every neighbour, receivers along a line that crosses many
    tiles (windows straddling tile corners), a source next to a tile corner,
    Neumann / Dirichlet faces on the edge tiles -- bit for bit against the
    per-step launches."""
    p = problems.make_problem(
        shape=shape, space_order=order, timesteps=60, bc=bc, seed=order + 3,
        nbl=nbl, num_sources=2, src_radius=4, num_receivers=300, rec_radius=4)
    monkeypatch.setenv("SIMWAVE_CUDA_MATH", math)
    monkeypatch.setenv("SIMWAVE_CUDA_LOOP", "launch")
    per_step = problems.clone(p)
    cuda_forward(per_step)
    monkeypatch.delenv("SIMWAVE_CUDA_LOOP")
    monkeypatch.setenv("SIMWAVE_CUDA_LOOP2D", "resident")
    resident = problems.clone(p)
    cuda_forward(resident)
    assert loop_kind() == LOOP_RESIDENT
    assert np.abs(per_step["u"]).max() > 0
    assert np.abs(per_step["receivers"]).max() > 0
    assert np.array_equal(per_step["u"], resident["u"])
    assert np.array_equal(per_step["receivers"], resident["receivers"])


# ---------------------------------------------------------------------------
# data-path hints (include/simwave_cuda.h, simwave_cuda_set_hint) and the fused
# small kernels: exact, so everything is compared bit for bit
# ---------------------------------------------------------------------------
HINT_ZERO_IN, HINT_OUT, HINT_MODEL = 1, 2, 3


@pytest.fixture
def hints():
    lib = core()
    lib.simwave_cuda_set_hint.argtypes = [ctypes.c_int, ctypes.c_longlong]

    def set_hint(key, value):
        assert lib.simwave_cuda_set_hint(key, value) == 0
    yield set_hint
    for key in (HINT_ZERO_IN, HINT_OUT, HINT_MODEL):
        lib.simwave_cuda_set_hint(key, 0)


@pytest.mark.parametrize("shape,order", [((44, 40, 48), 8), ((60, 72), 4)])
def test_hints_change_the_data_path_not_the_results(shape, order, hints, monkeypatch):
    monkeypatch.setenv("SIMWAVE_CUDA_MATH", "fast")
    T = 25
    p = problems.make_problem(shape=shape, space_order=order, timesteps=T, seed=21,
                              nbl=((0, 3),) * len(shape))
    base = problems.clone(p)
    cuda_forward(base)
    # zero wavefield on entry + only the slot the Solver returns
    hints(HINT_ZERO_IN, 1)
    hints(HINT_OUT, 1)
    q = problems.clone(p)
    q["u"][...] = np.nan          # never read
    cuda_forward(q)
    keep = T % 3
    assert np.array_equal(q["u"][keep], base["u"][keep])
    assert all(np.isnan(q["u"][s]).all() for s in range(3) if s != keep)
    assert np.array_equal(q["receivers"], base["receivers"])
    # receivers only
    hints(HINT_OUT, 2)
    q = problems.clone(p)
    q["u"][...] = np.nan
    cuda_forward(q)
    assert np.isnan(q["u"]).all()
    assert np.array_equal(q["receivers"], base["receivers"])
    # unknown hints are refused
    assert core().simwave_cuda_set_hint(9, 1) == -1
    assert core().simwave_cuda_set_hint(HINT_OUT, 5) == -1


def test_resident_model_serves_the_next_shots(hints, monkeypatch):
    """A survey under one model token: the second shot (other source and
    receiver tables) reuses the device model; results equal fresh runs.  A
    changed model under a NEW token is picked up."""
    monkeypatch.setenv("SIMWAVE_CUDA_MATH", "fast")
    shape = (40, 44, 48)
    shots = []
    for k, seed in enumerate((3, 4)):
        p = problems.make_problem(shape=shape, space_order=8, timesteps=20, seed=31,
                                  density=True, nbl=((0, 4), (3, 3), (4, 2)))
        q = problems.make_problem(shape=shape, space_order=8, timesteps=20, seed=seed,
                                  density=True, nbl=((0, 4), (3, 3), (4, 2)))
        for key in ("src_intervals", "src_values", "src_offsets", "rec_intervals",
                    "rec_values", "rec_offsets", "wavelet"):
            p[key] = q[key]
        shots.append(p)
    assert np.array_equal(shots[0]["velocity"], shots[1]["velocity"])
    fresh = [problems.clone(s) for s in shots]
    for f in fresh:
        cuda_forward(f)
    hints(HINT_MODEL, 77)
    for s, f in zip(shots, fresh):
        got = problems.clone(s)
        cuda_forward(got)
        assert np.array_equal(got["u"], f["u"])
        assert np.array_equal(got["receivers"], f["receivers"])
    # a stale model would be wrong here: other velocity under the same token is
    # the caller's broken promise, under a new token it must be re-read
    other = problems.clone(shots[0])
    other["velocity"] *= np.float32(1.02)
    want = problems.clone(other)
    hints(HINT_MODEL, 0)
    cuda_forward(want)
    hints(HINT_MODEL, 78)
    got = problems.clone(other)
    cuda_forward(got)
    assert np.array_equal(got["u"], want["u"])
    assert not np.array_equal(got["u"], fresh[0]["u"])


@pytest.mark.parametrize("math", ["strict", "fast"])
def test_sources_fused_into_the_step_kernel_match_the_source_kernel(math, monkeypatch):
    """Tiled 3D kernel: interior source windows are added by the thread that
    owns the cell; same bits as the stand-alone source kernel, next to a
    Neumann face, with overlapping windows and per-source wavelets."""
    monkeypatch.setenv("SIMWAVE_CUDA_MATH", math)
    r = 4
    shape = (40, 48, 80)
    src = [(r + 3.3, 20.2, 30.7), (r + 4.1, 21.0, 31.4), (30.5, 40.2, 70.9)]
    p = problems.make_problem(shape=shape, space_order=2 * r, timesteps=14, seed=5,
                              bc=(2, 1, 2, 1, 1, 2), nbl=((0, 3), (3, 3), (3, 3)),
                              num_sources=3, src_positions=src, multi_wavelet=True,
                              src_radius=3)
    fused = problems.clone(p)
    cuda_forward(fused)
    launches_fused = core().simwave_cuda_last_launch_count()
    monkeypatch.setenv("SIMWAVE_CUDA_SOURCES", "kernel")
    separate = problems.clone(p)
    cuda_forward(separate)
    assert core().simwave_cuda_last_launch_count() == launches_fused + 14
    assert np.abs(fused["u"]).max() > 0
    assert np.array_equal(fused["u"], separate["u"])
    assert np.array_equal(fused["receivers"], separate["receivers"])
    if math == "strict":
        ref = problems.clone(p)
        oracle.forward(ref)
        assert_identical(ref, fused)


@pytest.mark.parametrize("shape", [(36, 40, 44), (50, 64)])
def test_receivers_on_the_side_stream_match_inline_launches(shape, monkeypatch):
    monkeypatch.setenv("SIMWAVE_CUDA_MATH", "fast")
    monkeypatch.setenv("SIMWAVE_CUDA_LOOP", "launch")
    p = problems.make_problem(shape=shape, space_order=4, timesteps=30, seed=12,
                              num_receivers=40)
    side = problems.clone(p)
    cuda_forward(side)
    monkeypatch.setenv("SIMWAVE_CUDA_RECEIVERS", "inline")
    inline = problems.clone(p)
    cuda_forward(inline)
    assert np.abs(side["receivers"]).max() > 0
    assert np.array_equal(side["receivers"], inline["receivers"])
    assert np.array_equal(side["u"], inline["u"])
