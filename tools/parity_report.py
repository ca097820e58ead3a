#!/usr/bin/env python
"""
Parity report (development / DESIGN.md evidence): runs seeded problems on the
GPU through the `forward` C-ABI and on the CPU oracle, prints relative-L2 and
max-abs differences of the wavefield and the receiver traces.

    python tools/parity_report.py [--long]
"""
import argparse
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests"), os.path.join(REPO, "oracle"),
          os.path.join(REPO, "tests", "golden")):
    sys.path.insert(0, p)
import oracle  # noqa: E402
import problems  # noqa: E402
from cuda_abi import cuda_forward  # noqa: E402
from conftest import rel_l2  # noqa: E402

CASES = [
    # name, shape, order, density, steps, dtype
    ("2d so8 T=500", (120, 130), 8, False, 500, np.float32),
    ("2d so8 T=1500", (200, 210), 8, False, 1500, np.float32),
    ("2d so4 vd T=500", (120, 130), 4, True, 500, np.float32),
    ("3d so8 T=300", (70, 150, 140), 8, False, 300, np.float32),
    ("3d so8 T=600", (70, 150, 140), 8, False, 600, np.float32),
    ("3d so16 vd T=200", (48, 50, 52), 16, True, 200, np.float32),
    ("3d so4 T=300", (60, 90, 100), 4, False, 300, np.float32),
    ("3d so8 f64 T=200", (50, 60, 70), 8, False, 200, np.float64),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--smooth", action="store_true",
                    help="smooth velocity instead of per-point random")
    args = ap.parse_args()
    print("math mode:", os.environ.get("SIMWAVE_CUDA_MATH", "fast (default)"))
    for name, shape, order, density, steps, dtype in CASES:
        nbl = ((0, 6),) + ((5, 5),) * (len(shape) - 1)
        p = problems.make_problem(shape=shape, space_order=order,
                                  density=density, timesteps=steps, seed=3,
                                  smooth_density=True, nbl=nbl, dtype=dtype)
        a, b = problems.clone(p), problems.clone(p)
        t0 = time.time()
        oracle.forward(a)
        t1 = time.time()
        cuda_forward(b)
        mu = np.abs(a["u"]).max()
        mr = np.abs(a["receivers"]).max()
        print("%-18s u: rel-L2 %.2e max-abs/max|u| %.2e   rec: rel-L2 %.2e "
              "max-abs/max %.2e   (oracle %.1fs)" % (
                  name, rel_l2(b["u"], a["u"]),
                  np.abs(b["u"] - a["u"]).max() / mu,
                  rel_l2(b["receivers"], a["receivers"]),
                  np.abs(b["receivers"] - a["receivers"]).max() / mr,
                  t1 - t0), flush=True)


if __name__ == "__main__":
    main()
