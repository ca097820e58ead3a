"""
The adjoint operator (include/simwave_cuda.h section 1b; SURVEY.md section 8
f4).  The reference has no adjoint to compare with, so the contract is the
defining property <F w, d> = <w, F^T d> with F the REFERENCE's forward kernel:

  * CPU: the checker's restatement (oracle.adjoint: table exchange around the
    compiled reference forward) passes the dot-product test in float64 to
    1e-10 for every boundary-condition mix, 2D / 3D, several orders, damping
    layers, windows on Neumann and Dirichlet face planes;
  * GPU: the CUDA `adjoint` is bit-identical to that checker in strict mode,
    passes the dot-product test against the CUDA `forward` itself, and agrees
    within the float32 tolerance in the default math mode; Solver.adjoint goes
    through the same entry point.
"""
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "oracle"))

import oracle  # noqa: E402
import problems  # noqa: E402
from conftest import rel_l2  # noqa: E402

CASES = [
    # shape, order, bc
    ((40, 46), 2, (1, 1, 0, 1)),
    ((40, 46), 8, (2, 1, 2, 2)),
    ((44, 40), 4, (2, 2, 2, 2)),
    ((24, 26, 28), 2, (2, 1, 0, 1, 2, 2)),
    ((24, 26, 28), 8, (2, 2, 2, 2, 2, 2)),
    ((30, 26, 28), 4, (2, 2, 1, 2, 0, 2)),
]


def interior_problem(shape, order, bc, dtype, steps=30, seed=3, radius=2):
    """Sources / receivers whose windows lie among the interior points, some of
    them on the face planes themselves."""
    ndim, r = len(shape), order // 2
    rng = np.random.default_rng(seed)

    def positions(count):
        lo = np.array([r + float(radius)] * ndim)
        hi = np.array([n - r - 1.0 - radius for n in shape])
        return lo + (hi - lo) * rng.random((count, ndim))
    src, rec = positions(3), positions(5)
    src[0, 0] = r + radius                       # window starts on the z-before plane
    rec[0, -1] = shape[-1] - r - 1.0 - radius    # ... ends on the last-axis-after plane
    rec[1, 0] = r + radius + 0.3
    return problems.make_problem(
        shape=shape, space_order=order, dtype=dtype, timesteps=steps, bc=bc,
        nbl=((3, 3),) * ndim, num_sources=3, num_receivers=5, multi_wavelet=True,
        seed=seed, src_radius=radius, rec_radius=radius, src_positions=src,
        rec_positions=rec)


def dot_test(p, forward, adjoint, rng):
    """(<F w, d>, <w, F^T d>) for random w, d."""
    w = rng.standard_normal(p["wavelet"].shape).astype(p["wavelet"].dtype)
    d = rng.standard_normal(p["receivers"].shape).astype(p["receivers"].dtype)
    f = problems.clone(p)
    f["wavelet"] = w.copy()
    forward(f)
    a = problems.clone(p)
    a["receivers"] = d.copy()
    a["wavelet"] = np.zeros_like(w)
    adjoint(a)
    return (np.vdot(f["receivers"].astype(np.float64), d.astype(np.float64)),
            np.vdot(w.astype(np.float64), a["wavelet"].astype(np.float64)))


@pytest.mark.parametrize("shape,order,bc", CASES)
def test_checker_adjoint_is_the_transpose_of_the_reference_forward(shape, order, bc):
    p = interior_problem(shape, order, bc, np.float64)
    lhs, rhs = dot_test(p, oracle.forward, oracle.adjoint, np.random.default_rng(0))
    assert abs(lhs) > 0
    assert abs(lhs - rhs) <= 1e-10 * abs(lhs), (lhs, rhs)


def test_checker_adjoint_of_one_shared_wavelet_and_of_a_timestep_window():
    p = interior_problem((40, 46), 4, (2, 1, 0, 1), np.float64)
    p["wavelet"] = np.ascontiguousarray(p["wavelet"][:, 0])     # one wavelet for all sources
    p["begin_timestep"], p["end_timestep"] = 1, 24
    rng = np.random.default_rng(1)
    w = np.zeros_like(p["wavelet"])
    w[:24] = rng.standard_normal(24)
    d = np.zeros_like(p["receivers"])
    d[:24] = rng.standard_normal((24, 5))
    f = problems.clone(p)
    f["wavelet"] = w.copy()
    oracle.forward(f)
    a = problems.clone(p)
    a["receivers"] = d.copy()
    a["wavelet"] = np.full_like(w, 7.0)
    oracle.adjoint(a)
    assert np.all(a["wavelet"][24:] == 7.0)          # rows outside the range untouched
    lhs, rhs = np.vdot(f["receivers"][:24], d[:24]), np.vdot(w[:24], a["wavelet"][:24])
    assert abs(lhs - rhs) <= 1e-10 * abs(lhs)


# ---------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("shape,order,bc", CASES)
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_cuda_adjoint_strict_is_bit_identical_to_the_checker(shape, order, bc, dtype,
                                                             monkeypatch):
    from cuda_abi import cuda_adjoint
    monkeypatch.setenv("SIMWAVE_CUDA_MATH", "strict")
    p = interior_problem(shape, order, bc, dtype)
    rng = np.random.default_rng(5)
    p["receivers"] = rng.standard_normal(p["receivers"].shape).astype(dtype)
    a, b = problems.clone(p), problems.clone(p)
    a["wavelet"][...] = 0
    b["wavelet"][...] = 0
    oracle.adjoint(a)
    cuda_adjoint(b)
    assert np.abs(a["wavelet"]).max() > 0
    assert np.array_equal(a["wavelet"], b["wavelet"])
    assert np.array_equal(a["u"], b["u"])


@pytest.mark.gpu
@pytest.mark.parametrize("shape,order,bc", CASES)
def test_cuda_adjoint_is_the_transpose_of_cuda_forward(shape, order, bc, monkeypatch):
    from cuda_abi import cuda_adjoint, cuda_forward
    monkeypatch.setenv("SIMWAVE_CUDA_MATH", "fast")
    p = interior_problem(shape, order, bc, np.float64)
    lhs, rhs = dot_test(p, cuda_forward, cuda_adjoint, np.random.default_rng(2))
    assert abs(lhs - rhs) <= 1e-10 * abs(lhs), (lhs, rhs)
    # float32, default math: the same property at float32 accuracy, and the
    # adjoint source within the stated tolerance of the float64 checker
    q = interior_problem(shape, order, bc, np.float32)
    lhs, rhs = dot_test(q, cuda_forward, cuda_adjoint, np.random.default_rng(2))
    assert abs(lhs - rhs) <= 2e-4 * abs(lhs), (lhs, rhs)
    d = np.random.default_rng(4).standard_normal(q["receivers"].shape)
    g32, g64 = problems.clone(q), problems.clone(p)
    g32["receivers"] = d.astype(np.float32)
    g64["receivers"] = d.astype(np.float32).astype(np.float64)
    for k in ("velocity", "damp", "src_values", "rec_values", "coeff2"):
        g64[k] = q[k].astype(np.float64)
    g64["dt"] = np.float64(q["dt"])
    cuda_adjoint(g32)
    oracle.adjoint(g64)
    assert rel_l2(g32["wavelet"], g64["wavelet"]) <= 1e-5


@pytest.mark.gpu
def test_adjoint_refuses_what_it_does_not_cover():
    from cuda_abi import cuda_adjoint
    p = problems.make_problem(shape=(30, 30), space_order=4, timesteps=10, density=True)
    with pytest.raises(RuntimeError, match="constant density"):
        cuda_adjoint(p)


@pytest.mark.gpu
def test_solver_adjoint_through_the_public_api(monkeypatch):
    import simwave_b200 as api
    monkeypatch.setenv("SIMWAVE_CUDA_MATH", "fast")
    vel = np.full((61, 71), 1500.0, dtype=np.float64)
    vel[30:] = 2200.0
    space = api.SpaceModel((0, 600, 0, 700), (10., 10.), vel, space_order=4,
                           dtype=np.float64)
    space.config_boundary(damping_length=(0, 50, 50, 50),
                          boundary_condition=("null_neumann", "null_dirichlet",
                                              "none", "null_dirichlet"))
    time = api.TimeModel(space_model=space, tf=0.12)
    src = api.Source(space, coordinates=[(100., 350.)], window_radius=2)
    rec = api.Receiver(space, coordinates=[(60., 50. + 20 * i) for i in range(30)],
                       window_radius=2)
    solver = api.Solver(space, time, src, rec, api.RickerWavelet(15.0, time))
    u, recv = solver.forward()
    rng = np.random.default_rng(0)
    d = rng.standard_normal(recv.shape)
    lam, g = solver.adjoint(d)
    assert g.shape == (time.timesteps,) and lam.shape == u.shape
    w = solver.wavelet.values
    lhs, rhs = np.vdot(recv, d), np.vdot(w, g)
    assert abs(lhs - rhs) <= 1e-10 * abs(lhs), (lhs, rhs)
    with pytest.raises(ValueError):
        solver.adjoint(d[:-1])
