#!/usr/bin/env python
"""
Tile-configuration sweep for the tiled 3D kernel (development tool).

    python tools/sweep.py [--workload overthrust_3d] [--timesteps 40]
                          [--cfgs 0,1,2] [--zchunks 0,64] [--math strict,fast]

Builds the workload once, then times the device-resident time loop (plan API,
CUDA events) for every combination; prints one line per combination.
"""
import argparse
import ctypes
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
import workloads  # noqa: E402
from cuda_abi import core  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="overthrust_3d")
    ap.add_argument("--timesteps", type=int, default=40)
    ap.add_argument("--cfgs", default="0,1,2,3,4,5,6")
    ap.add_argument("--zchunks", default="0")
    ap.add_argument("--math", default="strict,fast")
    ap.add_argument("--space-order", type=int, default=None)
    ap.add_argument("--repeat", type=int, default=2)
    args = ap.parse_args()

    kwargs = {"timesteps": args.timesteps}
    if args.space_order:
        kwargs["space_order"] = args.space_order
    p = workloads.WORKLOADS[args.workload](**kwargs)
    pts = workloads.interior_points(p)
    bpp = workloads.bytes_per_point(p)
    T = p["end_timestep"]
    peak, _ = bench.measured_peak()

    lib = core()
    lib.simwave_plan_create.restype = ctypes.c_void_p
    lib.simwave_plan_run.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                     ctypes.POINTER(ctypes.c_double)]
    lib.simwave_plan_reset.argtypes = [ctypes.c_void_p]
    lib.simwave_plan_destroy.argtypes = [ctypes.c_void_p]

    combos = [("simple", "-", m) for m in args.math.split(",")]
    for m in args.math.split(","):
        for c in args.cfgs.split(","):
            for z in args.zchunks.split(","):
                combos.append(("tiled", "%s:%s" % (c, z), m))

    for kind, tile, math in combos:
        os.environ["SIMWAVE_CUDA_MATH"] = math
        os.environ["SIMWAVE_CUDA_KERNEL"] = "simple" if kind == "simple" else "auto"
        if kind == "tiled":
            os.environ["SIMWAVE_CUDA_TILE"] = tile
        keep = []
        pb = bench.make_problem_struct(p, keep)
        plan = lib.simwave_plan_create(ctypes.byref(pb))
        if not plan:
            print("%-7s %-8s %-6s  FAILED: %s" % (kind, tile, math,
                                                 lib.simwave_cuda_last_error().decode()))
            continue
        loop = ctypes.c_double()
        best = None
        for i in range(args.repeat + 1):
            lib.simwave_plan_reset(plan)
            if lib.simwave_plan_run(plan, 1, T, ctypes.byref(loop)) != 0:
                print("run failed:", lib.simwave_cuda_last_error().decode())
                break
            if i > 0:
                best = loop.value if best is None else min(best, loop.value)
        lib.simwave_plan_destroy(plan)
        if best:
            g = pts * T / best / 1e9
            print("%-7s %-8s %-6s  %8.3f ms/step  %8.2f Gpts/s  %5.1f%% of %d GB/s"
                  % (kind, tile, math, 1e3 * best / T, g, 100 * g * bpp / peak, peak),
                  flush=True)


if __name__ == "__main__":
    main()
