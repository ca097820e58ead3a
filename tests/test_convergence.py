"""
Quality of the numerical solution against the analytic 2D Green's function:
the reference's tests/test_convergence.py (accuracy, convergence in time,
convergence in space; same models, tolerances and fitted rates, :77-139) run
through simwave_b200's public API with ``Compiler(language='cuda')``.

float64, homogeneous 1.5 km/s medium, one source and one receiver, the
receiver trace compared with the Hankel-function solution.
"""
import numpy as np
import numpy.linalg as la
import pytest
from scipy.special import hankel2

import simwave_b200 as simwave

pytestmark = pytest.mark.gpu


def _solver(bbox, spacing, vel, order, dt, t0, tf, src, rec, f0):
    space_model = simwave.SpaceModel(
        bounding_box=bbox, grid_spacing=spacing,
        velocity_model=vel * np.ones((100, 100), dtype=np.float64),
        space_order=order)
    time_model = simwave.TimeModel(space_model=space_model, t0=t0, tf=tf)
    time_model.dt = dt
    source = simwave.Source(space_model, coordinates=src)
    receiver = simwave.Receiver(space_model, coordinates=rec)
    ricker = simwave.RickerWavelet(f0, time_model)
    solver = simwave.Solver(space_model, time_model, source, receiver, ricker,
                            compiler=simwave.Compiler(language="cuda"))
    return space_model, time_model, solver


def analytical_solution(space_model, freq, src, recs, dt):
    # reference tests/test_convergence.py:28-51
    time_model = simwave.TimeModel(space_model=space_model, t0=0, tf=3000)
    time_model.dt = dt
    ricker = simwave.RickerWavelet(freq, time_model)
    nf = int(time_model.timesteps / 2 + 1)
    df = 1 / time_model.tf
    frequencies = df * np.arange(nf)
    q = np.fft.fft(ricker.values)[:nf]
    xg = np.array([rec[0] for rec in recs])
    zg = np.array([rec[1] for rec in recs])
    r = np.sqrt((xg - src[0]) ** 2 + (zg - src[1]) ** 2)
    k = 2 * np.pi * frequencies / np.unique(space_model.velocity_model)
    u = np.zeros((nf), dtype=complex)
    u[1:-1] = hankel2(0, k[1:-1][None, :] * r[:, None])
    ui = np.fft.ifft(- 1j * np.pi * u * q, time_model.timesteps)
    return 1 / (2 * np.pi) * np.real(ui)


def accuracy(spacing, bbox, order, dt, t0, tf, c, f0, src, rec):
    space_model, time_model, solver = _solver(bbox, spacing, c, order, dt, t0,
                                              tf, src, rec, f0)
    u_num = solver.forward()[-1].flatten() / spacing[0] ** 2
    u_exact = analytical_solution(
        space_model, f0, src[0], rec, dt).flatten()[:time_model.timesteps]
    return la.norm(u_num - u_exact) / np.sqrt(u_num.size)


COMMON = dict(bbox=(-40, 440, -40, 440), t0=0, tf=150.075, c=1.5, f0=0.09,
              src=[(200, 200)], rec=[(260, 260)])


def test_accuracy():
    assert accuracy(spacing=(0.5, 0.5), order=8, dt=0.1, **COMMON) < 1e-4


def test_convergence_in_time():
    steps = [0.1, 0.075, 0.04, 0.025]
    accs = [accuracy(spacing=(0.5, 0.5), order=8, dt=dt, **COMMON)
            for dt in steps]
    rate = np.poly1d(np.polyfit(np.log(steps), np.log(accs), 1))[1]
    assert rate > 1.7


def test_convergence_in_space():
    spacings = [2.0, 2.5, 4.0]
    for order, min_rate in zip([2, 4, 6, 8, 10], [1.7, 4, 6, 7.7, 8.7]):
        errs = [accuracy(spacing=(h, h), order=order, dt=0.025, **COMMON)
                for h in spacings]
        rate = np.poly1d(np.polyfit(np.log(spacings), np.log(errs), 1))[1]
        assert rate > min_rate, (order, rate)
