#!/usr/bin/env python
"""
bench.py -- throughput of the acoustic forward-modelling time loop.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload overthrust_3d|...] [--timesteps T]

One "step" is one pass of the hot path over one synthetic shot: the whole time
loop (`forward`) of the workload.  The default workload is the configuration
the metric is quoted on, BASELINE.json configs[2]: the Overthrust-shaped 3D
model, space order 8, constant density, float32, 207x801x801 grid
(215x809x809 with halo), all 2651 time steps of tf = 4 s.

Printed JSON line (rank 0):
  value     Gpts/s with the problem resident in HBM, timed with CUDA events on
            the stream the kernels are launched on (max over ranks)
  e2e       the same metric through the drop-in `forward` C-ABI with host
            buffers: H2D of the model, time loop, D2H of the wavefield slots
            and receiver traces, wall clock around the call
  roofline  algorithmic bytes (20 B per grid-point update)
            / device time against the measured HBM peak
  cpu_baseline  the reference's cpu_openmp kernel (oracle/_ref) on a bounded
            number of time steps of the same arrays, on this host's cores

`--impl reference` times the reference's own CPU implementation alone.
Under torchrun every rank simulates its own shot of the workload (shot
parallelism: no data-path collective, weak scaling).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
for _p in (REPO, os.path.join(REPO, "tests"), os.path.join(REPO, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# thread placement for the CPU baseline must be set before libgomp loads
os.environ.setdefault("OMP_PROC_BIND", "true")
os.environ.setdefault("OMP_PLACES", "cores")

import numpy as np  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="overthrust_3d")
    ap.add_argument("--timesteps", type=int, default=None,
                    help="time steps per forward (default: the workload's own)")
    ap.add_argument("--cpu-timesteps", type=int, default=None,
                    help="time steps of the CPU baseline sample")
    ap.add_argument("--no-cpu", action="store_true",
                    help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-slab", action="store_true",
                    help="skip the slab-decomposition leg")
    ap.add_argument("--slab-only", action="store_true",
                    help="run only the slab-decomposition leg and print it")
    ap.add_argument("--slab-planes", type=int, default=512,
                    help="owned z-planes per GPU in the slab leg")
    ap.add_argument("--slab-n", type=int, default=1040)
    ap.add_argument("--slab-timesteps", type=int, default=60)
    ap.add_argument("--shots", type=int, default=0,
                    help="C5 survey leg (with --workload shot_3d): this many "
                         "shots over one shared model, dealt round-robin to the "
                         "ranks, each through the drop-in forward()")
    return ap.parse_args()


def measured_peak():
    """HBM copy bandwidth of this pool's B200s in GB/s: the driver-written
    MEASURED_PEAKS.json when present ("measured"), else the profiling recipe's
    fallback of 6.65 TB/s ("fallback")."""
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            peaks = json.load(f)
        value = float(peaks["hbm_gbs"])
        if value > 0:
            return value, "measured"
    except (OSError, ValueError, KeyError, TypeError):
        pass
    return 6650.0, "fallback"


def profiled_traffic(workload):
    """Bytes per launch of the dominant kernel from the committed ncu capture
    (profiles/traffic.json), or None."""
    path = os.path.join(REPO, "profiles", "traffic.json")
    try:
        with open(path) as f:
            e = json.load(f).get(workload)
        return None if e is None else int(e["dram_bytes_read"] + e["dram_bytes_write"])
    except (OSError, ValueError, KeyError):
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                 "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for name, flag in zip(names, f[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        # median over the busier half of the samples (the timed region)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------
def cpu_reference_run(p, timesteps, variant="ompfast"):
    """Reference CPU kernel on the first `timesteps` steps of problem p."""
    import oracle
    import problems
    kind = oracle.best_kind()
    if kind == "ref" and not oracle.available("ref", p["velocity"].ndim,
                                              p.get("density") is not None,
                                              p["velocity"].dtype, variant):
        variant = "omp"
    if kind == "port":
        variant = "omp"
    q = dict(p)
    q["u"] = np.zeros_like(p["u"])
    q["receivers"] = np.zeros_like(p["receivers"])
    q["end_timestep"] = timesteps
    seconds = oracle.forward(q, kind=kind, variant=variant)
    del problems
    return seconds, kind, variant


def _host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# taken before libgomp binds the main thread to one core (OMP_PROC_BIND)
HOST_THREADS = _host_threads()


def host_threads():
    return HOST_THREADS


def run_reference(args, p, rank, world):
    """--impl reference: the reference's cpu_openmp path alone."""
    import workloads
    if rank != 0:
        return
    pts = workloads.interior_points(p)
    sample = args.cpu_timesteps or max(1, int(3e9 // pts))
    os.environ.setdefault("OMP_NUM_THREADS", str(host_threads()))
    times = []
    for i in range(args.warmup + args.steps):
        s, kind, variant = cpu_reference_run(p, sample)
        if i >= args.warmup:
            times.append(s)
    total = sum(times)
    value = pts * sample * len(times) / total / 1e9
    line = {
        "impl": "reference", "metric": "Gpts/s", "value": value,
        "unit": "Gpts/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, p, sample),
        "cpu_baseline": {
            "value": value, "unit": "Gpts/s", "cores": host_threads(),
            "kind": "reference" if kind == "ref" else "port",
            "sample": "%d of %d time steps of the workload, %s build"
                      % (sample, p["full_timesteps"], variant)},
        "e2e": {"value": value, "unit": "Gpts/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, p, timesteps):
    return {
        "workload": "%s: grid %s (extended, halo %d), space_order %d, %s "
                    "density, float32, %d time steps per step, 1 source, "
                    "%d receivers" % (
                        p["name"], "x".join(map(str, p["velocity"].shape)),
                        p["space_order"] // 2, p["space_order"],
                        "variable" if p.get("density") is not None else "constant",
                        timesteps, len(p["rec_offsets"]) - 1),
        "l2_policy": "inputs larger than L2 (each field %.0f MiB)"
                     % (p["velocity"].nbytes / 2 ** 20),
        "parallelism": "shot-parallel, one shot per GPU" if args.gpus > 1
                       else "single GPU",
    }


# ---------------------------------------------------------------------------
def pinned_like(a):
    """Copy of ndarray `a` in page-locked host memory (torch allocator)."""
    import torch
    t = torch.empty(a.shape, dtype=torch.from_numpy(np.empty(0, a.dtype)).dtype,
                    pin_memory=True)
    out = t.numpy()
    out[...] = a
    out_keepalive.append(t)
    return out


out_keepalive = []


def run_slab_leg(args, rank, world, dist, barrier, max_over_ranks):
    """One large 3D model split into z-slabs, one per GPU, ghost planes kept
    current on the device over NVLink (simwave_b200/slab.py).  Weak scaling:
    --slab-planes owned planes per GPU.  Returns the "slab" object of the
    bench line (None on ranks other than 0)."""
    import workloads
    from simwave_b200 import slab
    q = workloads.slab_3d(rank=rank, world=world,
                          planes_per_gpu=args.slab_planes, n=args.slab_n,
                          timesteps=args.slab_timesteps)
    r = q["space_order"] // 2
    T = q["end_timestep"]
    plan = slab.Plan(q)
    if world > 1:
        def gather_bytes(b):
            out = [None] * world
            dist.all_gather_object(out, b)
            return out
        slab.connect_neighbours(plan, rank, world, gather_bytes)
    times = []
    for i in range(1 + max(1, args.steps)):
        plan.reset()
        barrier()
        t = plan.run(1, T)
        if i >= 1:
            times.append(t)
    total = max_over_ranks(sum(times))
    plan.destroy()
    nx, ny = q["velocity"].shape[1:]
    pts = q["owned_planes"] * (nx - 2 * r) * (ny - 2 * r)
    bpp = workloads.bytes_per_point(q)
    halo = (int(q["slab_up"]) + int(q["slab_down"])) * r * nx * ny * 4
    if rank != 0:
        return None
    peak, _ = measured_peak()
    value = world * pts * T * len(times) / total / 1e9
    return {
        "value": value, "unit": "Gpts/s", "scaling": "weak",
        "ms_per_timestep": 1e3 * total / len(times) / T,
        "roofline_frac_per_gpu": value / world * bpp / peak,
        "config": {
            "workload": "slab_3d (C4-shaped): global grid %s (extended), "
                        "variable density, space_order %d, %d owned planes per "
                        "GPU, %d time steps" % (
                            "x".join(map(str, q["global_shape"])),
                            q["space_order"], q["owned_planes"], T),
            "exchange": "per time step, r=%d planes per face pushed into the "
                        "neighbour's ghost planes through CUDA IPC peer "
                        "mappings (NVLink), device-side step flags" % r,
            "halo_bytes_per_step_per_gpu": halo},
    }


def run_survey_leg(args, p, rank, world, barrier, max_over_ranks):
    """C5: a multi-shot survey.  The model stays on the host once per rank
    (page-locked); every shot is one drop-in forward() call with its own
    tables, wavefield and traces; shot s runs on rank s mod world.  No
    data-path collective.  Returns the "survey" object of the bench line."""
    import workloads
    from cuda_abi import cuda_forward
    T = p["end_timestep"]
    pts = workloads.interior_points(p)
    host = dict(p)
    for key in ("velocity", "damp", "wavelet"):
        host[key] = pinned_like(host[key])
    host["u"] = pinned_like(p["u"])
    host["receivers"] = pinned_like(p["receivers"])
    mine = list(range(rank, args.shots, world))
    traces = {}

    def shoot(shot):
        q = workloads.reshoot(p, shot)
        for key in ("src_intervals", "src_values", "src_offsets",
                    "rec_intervals", "rec_values", "rec_offsets"):
            host[key] = q[key]
        host["u"][...] = 0
        host["receivers"][...] = 0
        cuda_forward(host)
        traces[shot] = float(np.abs(host["receivers"]).max())

    if mine:
        shoot(mine[0])                      # warm-up: allocation caches, clocks
    barrier()
    t0 = time.perf_counter()
    for shot in mine:
        shoot(shot)
    seconds = time.perf_counter() - t0
    barrier()
    total = max_over_ranks(seconds)
    if rank != 0:
        return None
    return {"value": args.shots * pts * T / total / 1e9, "unit": "Gpts/s",
            "shots": args.shots, "shots_per_second": args.shots / total,
            "seconds": total, "scaling": "strong",
            "config": {"workload": "survey of %d shots over %s" % (
                args.shots, workload_config(args, p, T)["workload"]),
                "parallelism": "shot s on rank s mod %d, one forward() per "
                               "shot, model arrays shared on the host" % world},
            "max_abs_trace_of_first_shot": traces.get(0)}


def run_ours(args, p, rank, world, local_rank):
    import torch
    import workloads
    from cuda_abi import core, cuda_forward, last_timing

    torch.cuda.set_device(local_rank)
    os.environ["SIMWAVE_CUDA_DEVICE"] = str(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from simwave_b200 import slab
    lib = core()
    if lib.simwave_cuda_device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device; the CUDA backend has no "
                         "CPU fallback")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.slab_only:
        res = run_slab_leg(args, rank, world, dist, barrier, max_over_ranks)
        if rank == 0:
            print(json.dumps({"slab": res, "n_gpus": world}))
        if dist is not None:
            dist.destroy_process_group()
        return

    T = p["end_timestep"]
    pts = workloads.interior_points(p)
    bpp = workloads.bytes_per_point(p)

    # ---- device-resident throughput (plan API) -----------------------------
    plan = slab.Plan(p)

    def one_step():
        plan.reset()
        return plan.run(1, T)

    for _ in range(args.warmup):
        one_step()
    barrier()
    with ClockSampler(local_rank) as clocks:
        wall0 = time.perf_counter()
        device_seconds = [one_step() for _ in range(args.steps)]
        barrier()
        wall = time.perf_counter() - wall0
    launches = plan.launches() * args.steps
    dev_total = max_over_ranks(sum(device_seconds))
    wall = max_over_ranks(wall)
    plan.destroy()

    value = world * pts * T * args.steps / dev_total / 1e9
    achieved = pts * T * args.steps * bpp / dev_total / 1e9   # per GPU
    peak, peak_kind = measured_peak()

    # ---- end to end through the drop-in forward() ---------------------------
    e2e = None
    if not args.no_e2e:
        host = dict(p)
        for key in ("velocity", "damp", "density", "wavelet"):
            if host.get(key) is not None:
                host[key] = pinned_like(host[key])
        host["u"] = pinned_like(p["u"])
        host["receivers"] = pinned_like(p["receivers"])
        h2d = sum(host[k].nbytes for k in
                  ("velocity", "damp", "wavelet", "src_intervals", "src_values",
                   "src_offsets", "rec_intervals", "rec_values", "rec_offsets"))
        if host.get("density") is not None:
            h2d += host["density"].nbytes
        if host["u"].shape[0] == 3:
            # a page-locked three-slot wavefield is uploaded outright (cheaper
            # than scanning it for zeros on the host)
            h2d += host["u"].nbytes
        d2h = host["u"].nbytes + host["receivers"].nbytes
        e2e_warm = max(1, min(args.warmup, 1))
        times = []
        for i in range(e2e_warm + args.steps):
            host["u"][...] = 0
            host["receivers"][...] = 0
            barrier()
            t0 = time.perf_counter()
            cuda_forward(host)
            torch.cuda.synchronize()
            dt_wall = time.perf_counter() - t0
            if i >= e2e_warm:
                times.append(dt_wall)
        tm = last_timing()
        e2e_total = max_over_ranks(sum(times))
        e2e = {"value": world * pts * T * len(times) / e2e_total / 1e9,
               "unit": "Gpts/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h),
               "seconds_per_step": e2e_total / len(times),
               "breakdown_last_call": tm,
               "host_memory": "pinned (torch pin_memory)"}

    # ---- CPU baseline beside it (rank 0, N == 1 only) ------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sample = args.cpu_timesteps or max(1, int(3e9 // pts))
        os.environ.setdefault("OMP_NUM_THREADS", str(host_threads()))
        try:
            cpu_reference_run(p, 1)      # warm the pages / threads
            s, kind, variant = cpu_reference_run(p, sample)
            cpu = {"value": pts * sample / s / 1e9, "unit": "Gpts/s",
                   "cores": host_threads(),
                   "kind": "reference" if kind == "ref" else "port",
                   "sample": "%d of %d time steps of the same arrays, "
                             "%s build, OMP_PROC_BIND=true" % (
                                 sample, p["full_timesteps"], variant)}
        except Exception as e:   # the baseline must not sink the bench line
            cpu = {"value": None, "unit": "Gpts/s", "cores": host_threads(),
                   "kind": "port", "sample": "failed: %s" % e}

    # ---- slab decomposition leg (C4-shaped, weak scaling) --------------------
    slab_result = None
    if not args.no_slab:
        try:
            slab_result = run_slab_leg(args, rank, world, dist, barrier,
                                       max_over_ranks)
        except Exception as e:      # the extra leg must not sink the bench line
            if world > 1:
                raise               # ranks wait on each other: fail together
            slab_result = {"error": "%s: %s" % (type(e).__name__, e)}

    survey = None
    if args.shots > 0 and p["name"] == "shot_3d":
        survey = run_survey_leg(args, p, rank, world, barrier, max_over_ranks)

    if rank == 0:
        line = {
            "metric": "Gpts/s", "value": value, "unit": "Gpts/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if p["velocity"].dtype == np.float32 else "f64",
            "data": "synthetic",
            "config": workload_config(args, p, T),
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak,
                "traffic": profiled_traffic(p["name"])
                if os.environ.get("SIMWAVE_CUDA_MATH", "fast") == "fast" else None,
                "peak_source": peak_kind,
                "bytes_per_point": bpp,
                "note": "dominant kernel = stencil step; achieved = %d B x "
                        "interior points x time steps / CUDA-event time of "
                        "the loop (source and receiver kernels included)" % bpp},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "slab": slab_result,
            "survey": survey,
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
            "wall_ms_per_step": 1e3 * wall / args.steps,
            "math": os.environ.get("SIMWAVE_CUDA_MATH", "fast"),
        }
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    import workloads
    builder = workloads.WORKLOADS[args.workload]
    if args.impl == "reference" and rank != 0:
        return
    kwargs = {}
    if args.timesteps:
        kwargs["timesteps"] = args.timesteps
    if args.workload == "shot_3d":
        kwargs["shot"] = rank
    p = None if args.slab_only else builder(**kwargs)
    if args.impl == "reference":
        run_reference(args, p, rank, world)
    else:
        run_ours(args, p, rank, world, local_rank)


if __name__ == "__main__":
    main()
