#!/usr/bin/env python
"""
Tile-configuration sweep for the tiled 3D kernel (development tool).

    python tools/sweep.py [--workload overthrust_3d] [--timesteps 40]
                          [--cfgs 0,1,2] [--zchunks 0,64] [--math strict,fast]

Builds the workload once, then times the device-resident time loop (plan API,
CUDA events) for every combination; prints one line per combination.
"""
import argparse
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
import workloads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="overthrust_3d")
    ap.add_argument("--timesteps", type=int, default=40)
    ap.add_argument("--cfgs", default="0,1,2,3,4,5,6")
    ap.add_argument("--zchunks", default="0")
    ap.add_argument("--math", default="strict,fast")
    ap.add_argument("--space-order", type=int, default=None)
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--kw", action="append", default=[],
                    help="extra integer keyword of the workload builder, key=value")
    ap.add_argument("--no-simple", action="store_true", help="skip the plain kernel")
    ap.add_argument("--dtype", default=None, choices=["f32", "f64"],
                    help="precision of the workload (builders that take a dtype)")
    ap.add_argument("--prefetch", default="",
                    help="comma list of SIMWAVE_CUDA_PREFETCH distances to cross with the tiles")
    args = ap.parse_args()

    kwargs = {"timesteps": args.timesteps}
    if args.space_order:
        kwargs["space_order"] = args.space_order
    if args.dtype:
        import numpy as np
        kwargs["dtype"] = np.float32 if args.dtype == "f32" else np.float64
    for kv in args.kw:
        k, v = kv.split("=")
        kwargs[k] = int(v)
    p = workloads.WORKLOADS[args.workload](**kwargs)
    pts = workloads.interior_points(p)
    bpp = workloads.bytes_per_point(p)
    T = p["end_timestep"]
    peak, _ = bench.measured_peak()

    from simwave_b200 import slab

    combos = [] if args.no_simple else [("simple", "-", m, None) for m in args.math.split(",")]
    for m in args.math.split(","):
        for c in args.cfgs.split(","):
            for z in args.zchunks.split(","):
                for pf in (args.prefetch.split(",") if args.prefetch else [None]):
                    combos.append(("tiled", "%s:%s" % (c, z), m, pf))

    for kind, tile, math, extra in combos:
        os.environ["SIMWAVE_CUDA_MATH"] = math
        tile_label = tile
        pf = extra
        if pf is not None:
            os.environ["SIMWAVE_CUDA_PREFETCH"] = pf
            tile_label += "/pf%s" % pf

        os.environ["SIMWAVE_CUDA_KERNEL"] = "simple" if kind == "simple" else "auto"
        if kind == "tiled":
            os.environ["SIMWAVE_CUDA_TILE"] = tile
        try:
            plan = slab.Plan(p)
        except RuntimeError as e:
            print("%-7s %-14s %-6s  FAILED: %s" % (kind, tile_label, math, e))
            continue
        best = None
        for i in range(args.repeat + 1):
            plan.reset()
            t = plan.run(1, T)
            if i > 0:
                best = t if best is None else min(best, t)
        plan.destroy()
        if best:
            g = pts * T / best / 1e9
            print("%-7s %-14s %-6s  %8.3f ms/step  %8.2f Gpts/s  %5.1f%% of %d GB/s"
                  % (kind, tile_label, math, 1e3 * best / T, g, 100 * g * bpp / peak, peak),
                  flush=True)


if __name__ == "__main__":
    main()
