"""
Pins the CPU checker (oracle/wave_oracle.c, "the port") before anything is
compared against it:

* bit-for-bit against the reference's own wave.c compiled into oracle/_ref/
  (when those prebuilt files are present) on seeded problems that switch on
  every feature, and
* bit-for-bit against the golden vectors that make_golden.py produced by
  running the unmodified reference package end to end -- here the port is
  driven through simwave_b200's own front end via the custom-kernel hook
  (Compiler(cfile=...)), so this also pins the whole host side.

No GPU involved.
"""
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

PORT_SOURCE = os.path.join(REPO, "oracle", "wave_oracle.c")

needs_ref = pytest.mark.skipif(
    not oracle.available("ref"),
    reason="oracle/_ref not built (needs /root/reference at build time)")


def port_compiler(dimension, density, openmp=False):
    """simwave_b200 Compiler that builds the port through the cfile hook."""
    flags = "-O3 -fPIC -Wall -std=c99 -shared -DNDIM={} -DVARDEN={}".format(
        dimension, int(bool(density)))
    if openmp:
        return api.Compiler(cc="/usr/bin/gcc", language="cpu_openmp",
                            cflags=flags + " -fopenmp -DORACLE_OMP",
                            cfile=PORT_SOURCE)
    return api.Compiler(cc="/usr/bin/gcc", language="c", cflags=flags,
                        cfile=PORT_SOURCE)


VARIANTS = [
    # shape, order, density, dtype, stride, timesteps
    ((40, 52), 2, False, np.float32, 0, 30),
    ((40, 52), 8, False, np.float32, 1, 30),
    ((44, 36), 4, True, np.float32, 2, 31),
    ((44, 36), 20, True, np.float64, 3, 31),
    ((40, 52), 6, False, np.float64, 4, 29),
    ((20, 22, 24), 2, False, np.float32, 0, 20),
    ((20, 22, 24), 8, False, np.float32, 5, 21),
    ((24, 22, 22), 4, True, np.float32, 0, 20),
    ((22, 26, 20), 4, True, np.float32, 2, 21),     # nx != ny: the ir*nx quirk
    ((22, 20, 27), 6, True, np.float64, 1, 15),     # nx != ny, float64
    ((28, 28, 28), 12, True, np.float64, 0, 12),
]


@needs_ref
@pytest.mark.parametrize("shape,order,density,dtype,stride,steps", VARIANTS)
def test_port_bit_identical_to_reference(shape, order, density, dtype, stride,
                                         steps):
    ndim = len(shape)
    r = order // 2
    nbl = tuple((0, 3) if a == 0 else (2, 4) for a in range(ndim))
    p = problems.make_problem(
        shape=tuple(n + 0 for n in shape), space_order=order, density=density,
        dtype=dtype, timesteps=steps, saving_stride=stride, nbl=nbl,
        num_sources=3, num_receivers=9, src_radius=min(4, r + 1),
        rec_radius=2, multi_wavelet=(stride % 2 == 0), seed=order + stride)
    a, b = problems.clone(p), problems.clone(p)
    oracle.forward(a, kind="ref")
    oracle.forward(b, kind="port")
    assert np.abs(a["u"]).max() > 0
    assert np.array_equal(a["u"], b["u"])
    assert np.array_equal(a["receivers"], b["receivers"])


@needs_ref
def test_port_openmp_matches_reference_openmp():
    p = problems.make_problem(shape=(30, 32, 34), space_order=8, timesteps=12,
                              num_sources=2, seed=3)
    a, b = problems.clone(p), problems.clone(p)
    oracle.forward(a, kind="ref", variant="omp")
    oracle.forward(b, kind="port", variant="omp")
    assert np.array_equal(a["u"], b["u"])
    assert np.array_equal(a["receivers"], b["receivers"])


@pytest.mark.parametrize("name", sorted(cases.SMALL_CASES))
def test_port_through_front_end_matches_golden(golden, workdir, name):
    """Front end + port == unmodified reference, bit for bit."""
    cfg = cases.SMALL_CASES[name]
    ref = golden("forward_small")
    solver = cases.small_solver(
        api, name, port_compiler(cfg["dimension"], cfg["density"]))
    u, recv = solver.forward()
    assert solver.time_model.timesteps == int(ref[name + "/timesteps"])
    assert np.array_equal(recv, ref[name + "/recv"])
    assert np.array_equal(u[ref[name + "/u_idx"]], ref[name + "/u"])
    l2 = np.sqrt(np.sum(u.astype(np.float64).reshape(u.shape[0], -1) ** 2,
                        axis=1))
    assert np.array_equal(l2, ref[name + "/u_l2"])


@pytest.mark.parametrize("space_order", [2, 8])
def test_port_solution_2d_golden(golden, workdir, space_order):
    """Reference tests/test_solution.py, 2D: the regenerated field bit for
    bit, the reference's checked-in .npy within its own atol=1e-5."""
    ref = golden("solution_2d_so%d" % space_order)
    solver = cases.solution_solver(api, 2, space_order, port_compiler(2, 0))
    u, recv = solver.forward()
    assert np.array_equal(u, ref["u"])
    assert np.array_equal(recv, ref["recv"])
    assert np.allclose(u, ref["u_reference_npy"], atol=1e-5)


@pytest.mark.slow
@pytest.mark.parametrize("space_order", [2, 8])
def test_port_solution_3d_golden(golden, workdir, space_order):
    """Reference tests/test_solution.py, 3D (the reference's 3D .npy files
    are absent from its tree; the golden is a regeneration)."""
    ref = golden("solution_3d_so%d" % space_order)
    solver = cases.solution_solver(api, 3, space_order, port_compiler(3, 0))
    u, recv = solver.forward()
    f = u[0]
    c = [n // 2 for n in f.shape]
    assert np.array_equal(f[c[0]], ref["plane_z"])
    assert np.array_equal(f[:, c[1]], ref["plane_x"])
    assert np.array_equal(f[:, :, c[2]], ref["plane_y"])
    assert np.array_equal(recv, ref["recv"])
    assert float(np.sqrt(np.sum(f.astype(np.float64) ** 2))) == float(ref["l2"])


def test_density_one_equals_constant_density(workdir):
    """Reference tests/test_solution.py:15-16,22-23: a unit density field
    must reproduce the constant-density result (allclose, atol=1e-5)."""
    base = cases.small_solver(api, "small2d_const_f32", port_compiler(2, 0))
    u0, r0 = base.forward()
    sm = base.space_model
    den = np.ones(sm.shape, dtype=np.float32)
    sm2 = api.SpaceModel(sm.bounding_box, sm.grid_spacing, sm.velocity_model,
                         density_model=den, space_order=sm.space_order)
    sm2.config_boundary(
        damping_length=sm.damping_length,
        boundary_condition=sm.boundary_condition,
        damping_polynomial_degree=sm.damping_polynomial_degree,
        damping_alpha=sm.damping_alpha)
    tm = api.TimeModel(sm2, tf=0.12)
    solver = api.Solver(
        sm2, tm, api.Source(sm2, base.sources.coordinates, 3),
        api.Receiver(sm2, base.receivers.coordinates, 2),
        api.MultiWavelet(base.wavelet.values, tm), port_compiler(2, 1))
    u1, r1 = solver.forward()
    assert np.allclose(u0, u1, atol=1e-5)
    assert np.allclose(r0, r1, atol=1e-5)
