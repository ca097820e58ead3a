#!/usr/bin/env python
"""
SURVEY.md section 8 f3: a 1024^3 variable-density model through the PUBLIC API
(SpaceModel -> TimeModel -> Source / Receiver -> Solver.forward), the path on
which the reference front end runs out of memory (model.py:210-261 builds
float64 coordinate meshes of the whole grid).

    python tools/api_1024.py [--n 1024] [--timesteps 100] [--gpus 1]

Prints one JSON line: grid, time steps, wall time of every stage, peak host
RSS, loop throughput.  With --gpus N the drop-in forward() itself spreads the
z-slabs over N devices of this process (SIMWAVE_CUDA_NGPUS).
"""
import argparse
import contextlib
import json
import os
import resource
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))


def rss_gb():
    return resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 2 ** 20


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1024)
    ap.add_argument("--timesteps", type=int, default=100)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--space-order", type=int, default=16)
    args = ap.parse_args()
    if args.gpus > 1:
        os.environ["SIMWAVE_CUDA_NGPUS"] = str(args.gpus)
    import simwave_b200 as api
    from cuda_abi import last_timing

    n, h = args.n, 10.0
    # physical model: the grid minus the 40-point layers C4 puts on every side
    # but the top (extended grid = n + 16 per axis at order 16)
    phys = (n - 40, n - 80, n - 80)
    stages = {}
    t0 = time.perf_counter()
    z = np.arange(phys[0], dtype=np.float32)[:, None, None] / phys[0]
    x = np.arange(phys[1], dtype=np.float32)[None, :, None]
    y = np.arange(phys[2], dtype=np.float32)[None, None, :]
    vel = (1500.0 + 2400.0 * z + 300.0 * np.sin(x / 53.0) * np.cos(y / 41.0)).astype(np.float32)
    rho = (1000.0 + 900.0 * z + 400.0 * np.cos(x / 45.0 + y / 58.0)).astype(np.float32)
    stages["synthetic_model_s"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    box = (0, (phys[0] - 1) * h, 0, (phys[1] - 1) * h, 0, (phys[2] - 1) * h)
    space = api.SpaceModel(bounding_box=box, grid_spacing=(h, h, h), velocity_model=vel,
                           density_model=rho, space_order=args.space_order,
                           dtype=np.float32)
    del vel, rho
    space.config_boundary(
        damping_length=(0, 40 * h, 40 * h, 40 * h, 40 * h, 40 * h),
        boundary_condition=("null_neumann", "null_dirichlet", "null_dirichlet",
                            "null_dirichlet", "null_dirichlet", "null_dirichlet"),
        damping_polynomial_degree=3, damping_alpha=0.001)
    stages["space_model_s"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    time_model = api.TimeModel(space_model=space, tf=1.0)
    tf = float(time_model.dt) * (args.timesteps - 1) * (1 - 1e-6)
    time_model = api.TimeModel(space_model=space, tf=tf)
    size = [(m - 1) * h for m in phys]
    source = api.Source(space, coordinates=[(20.0, size[1] / 2, size[2] / 2)], window_radius=4)
    receiver = api.Receiver(space, coordinates=[(20.0, size[1] / 2, size[2] * i / 1023.0)
                                                for i in range(1024)], window_radius=4)
    wavelet = api.RickerWavelet(8.0, time_model)
    solver = api.Solver(space, time_model, source, receiver, wavelet)
    stages["acquisition_s"] = time.perf_counter() - t0

    t0 = time.perf_counter()
    ext = space.extended_velocity_model.shape
    space.extended_density_model, space.damping_mask
    stages["extended_arrays_s"] = time.perf_counter() - t0
    rss_before = rss_gb()

    runs = []
    for _ in range(2):          # the second call finds the model on the device
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(sys.stderr):
            u, rec = solver.forward()
        runs.append({"wall_s": time.perf_counter() - t0, **last_timing()})
    r = args.space_order // 2
    pts = float(np.prod([m - 2 * r for m in ext]))
    T = time_model.timesteps
    print(json.dumps({
        "what": "1024^3-class variable-density model through SpaceModel -> Solver.forward()",
        "extended_grid": list(ext), "space_order": args.space_order, "timesteps": T,
        "gpus": args.gpus, "stages": stages, "forward_calls": runs,
        "loop_gpts_per_s": pts * T / runs[-1]["loop"] / 1e9,
        "e2e_gpts_per_s": pts * T / runs[-1]["wall_s"] / 1e9,
        "peak_host_rss_gb": rss_gb(), "host_rss_before_forward_gb": rss_before,
        "max_abs_wavefield": float(np.abs(u).max()),
        "max_abs_trace": float(np.abs(rec).max())}))


if __name__ == "__main__":
    main()
