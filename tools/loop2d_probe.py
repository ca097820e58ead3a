#!/usr/bin/env python
"""Where the time of a 2D step goes (development tool): device time per step of
the persistent loop for the Marmousi-shaped workload with receivers / sources
removed, and for a tiny grid (pure barrier cost)."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import workloads  # noqa: E402
import problems  # noqa: E402
from simwave_b200 import slab  # noqa: E402


def timed(p, label):
    T = p["end_timestep"]
    plan = slab.Plan(p)
    best = None
    for i in range(3):
        plan.reset()
        t = plan.run(1, T)
        best = t if best is None else min(best, t)
    plan.destroy()
    print("%-44s %7.2f us/step" % (label, 1e6 * best / T), flush=True)


def without_receivers(p):
    q = dict(p)
    q["rec_intervals"] = p["rec_intervals"][:4]
    q["rec_values"] = p["rec_values"][:int(p["rec_offsets"][1])]
    q["rec_offsets"] = p["rec_offsets"][:2]
    q["receivers"] = np.zeros((p["receivers"].shape[0], 1), dtype=p["receivers"].dtype)
    return q


def main():
    p = workloads.marmousi_2d(timesteps=600)
    c1 = workloads.readme_2d(timesteps=300)
    for mode in ("resident", "grid", "launch"):
        os.environ["SIMWAVE_CUDA_LOOP2D"] = mode
        if mode == "launch":
            os.environ["SIMWAVE_CUDA_LOOP"] = "launch"
        timed(c1, "readme 2D, 512 receivers [%s]" % mode)
        timed(workloads.marmousi_2d(timesteps=600, dtype=np.float64),
              "marmousi float64, 1700 receivers [%s]" % mode)
        timed(p, "marmousi, 1700 receivers [%s]" % mode)
        timed(without_receivers(p), "marmousi, 1 receiver [%s]" % mode)
        q = without_receivers(p)
        q["wavelet"] = np.zeros_like(q["wavelet"])
        timed(q, "marmousi, 1 receiver, silent source [%s]" % mode)
        tiny = problems.make_problem(shape=(40, 40), space_order=8, timesteps=600,
                                     num_sources=1, num_receivers=1, seed=1)
        timed(tiny, "40x40 grid, 1 receiver [%s]" % mode)


if __name__ == "__main__":
    main()
