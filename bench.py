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
  e2e       the same metric through the drop-in `forward` C-ABI with pinned
            host buffers: H2D of the model, time loop, D2H of the wavefield
            slots and receiver traces, wall clock around the call
  e2e_pageable / e2e_api  (N = 1) the same call with pageable NumPy arrays,
            and simwave_b200.Solver.forward() of the public API, front end
            included
  roofline  algorithmic bytes (20 B per grid-point update)
            / device time against the measured HBM peak
  cpu_baseline  the reference's cpu_openmp kernel (oracle/_ref) on a bounded
            number of time steps of the same arrays, on this host's cores
  slab / slab_strong  C4-shaped variable-density so-16 model split into z-slabs
            over the ranks: 1024 owned planes per GPU (weak), and the 1040^3
            grid itself split N ways (strong, N > 1)
  survey    C5: 64 shots x 512^3 through forward(), shot s on rank s mod N
  configs_2d  (N = 1) C1 and C2: device loop, forward() and the CPU kernel
  c3_f64    (N = 1) C3 in float64, the precision the reference's own benchmark
            script uses: device loop over 300 time steps, 40 B per point

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
# tests/ holds the ctypes callers of the C-ABI (cuda_abi.py, abi.py); oracle/
# is put on the path only by the CPU legs (cpu_reference_run)
for _p in (REPO, os.path.join(REPO, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

# Thread placement of the CPU arm must be in the environment before libgomp
# loads -- but ONLY in a process that runs nothing else: with OMP_PROC_BIND set,
# libgomp pins the main thread to one core when it initialises (at `import
# torch`), every thread created later inherits that one-core mask, and the
# staging / drain / prefault threads of the CUDA library then share a single
# core (measured: pageable uploads 5 GB/s instead of 27; under torchrun all
# ranks' main threads land on core 0).  The GPU arm therefore runs its
# cpu_baseline legs in a child process (cpu_reference_child).
if "--impl" in sys.argv and "reference" in sys.argv or "--cpu-child" in sys.argv:
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
    ap.add_argument("--slab-planes", type=int, default=1024,
                    help="owned z-planes per GPU in the weak-scaling slab leg "
                         "(1024 = the 1024^3-per-GPU configuration of BASELINE.md)")
    ap.add_argument("--no-slab-strong", action="store_true",
                    help="skip the strong-scaling leg (C4 proper: 1040^3 split "
                         "over the ranks), which runs when N > 1")
    ap.add_argument("--slab-n", type=int, default=1040)
    ap.add_argument("--slab-timesteps", type=int, default=60)
    ap.add_argument("--shots", type=int, default=64,
                    help="C5 survey leg: this many shots over one shared 512^3 "
                         "model, dealt round-robin to the ranks, each through "
                         "the drop-in forward()")
    ap.add_argument("--shot-timesteps", type=int, default=300)
    ap.add_argument("--no-survey", action="store_true")
    ap.add_argument("--no-2d", action="store_true",
                    help="skip the two 2D configurations (C1, C2; N = 1 only)")
    ap.add_argument("--cpu-child", action="store_true",
                    help="internal: time the reference's CPU kernel on "
                         "--cpu-timesteps steps of --workload and print one JSON line")
    ap.add_argument("--no-f64", action="store_true",
                    help="skip the float64 C3 leg (N = 1 only)")
    ap.add_argument("--f64-timesteps", type=int, default=300)
    ap.add_argument("--no-api", action="store_true",
                    help="skip the e2e_api leg (Solver.forward, N = 1 only)")
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
def set_cpu_threads():
    """All host cores for the CPU arm.  torchrun exports OMP_NUM_THREADS=1 and
    libgomp may already be initialised by the time the CPU leg runs, so the
    count is set in the environment AND through the OpenMP runtime."""
    n = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n)
    except (OSError, AttributeError):
        pass
    return n


def cpu_reference_run(p, timesteps, variant="ompfast"):
    """Reference CPU kernel on the first `timesteps` steps of problem p."""
    oracle_dir = os.path.join(REPO, "oracle")
    if oracle_dir not in sys.path:
        sys.path.insert(0, oracle_dir)
    import oracle
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
    return seconds, kind, variant


def cpu_reference_child(workload, timesteps, builder_timesteps=None):
    """The reference's cpu_openmp kernel on the first `timesteps` steps of a
    named workload, in a child process of its own (all host threads, bound to
    cores), after one warm-up run.  Returns (seconds, kind, variant)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-child", "--workload", workload,
           "--cpu-timesteps", str(int(timesteps))]
    if builder_timesteps:
        cmd += ["--timesteps", str(int(builder_timesteps))]
    env = dict(os.environ)
    env.update({"OMP_NUM_THREADS": str(host_threads()), "OMP_PROC_BIND": "true",
                "OMP_PLACES": "cores", "CUDA_VISIBLE_DEVICES": ""})
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=1200)
    for line in reversed(out.stdout.strip().splitlines()):
        if line.startswith("{"):
            r = json.loads(line)
            return r["seconds"], r["kind"], r["variant"]
    raise RuntimeError("CPU child failed: " + (out.stderr.strip().splitlines() or ["?"])[-1])


def run_cpu_child(args):
    import workloads
    kwargs = {"timesteps": args.timesteps} if args.timesteps else {}
    p = workloads.WORKLOADS[args.workload](**kwargs)
    n = set_cpu_threads()
    steps = args.cpu_timesteps or p["end_timestep"]
    cpu_reference_run(p, min(steps, 2))          # warm the pages / threads
    seconds, kind, variant = cpu_reference_run(p, steps)
    print(json.dumps({"seconds": seconds, "kind": kind, "variant": variant, "threads": n}))


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
    set_cpu_threads()
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


def l2_policy(p):
    """How the timing rules' L2 clause is met by this workload: a time step
    streams five fields (u_prev, u_cur, u_next, c0, q [+ density terms])."""
    field = p["velocity"].nbytes
    if 5 * field > 2 * 126e6:
        return ("inputs larger than L2 (each field %.0f MiB, a time step streams "
                ">= 5 of them; no flush needed)" % (field / 2 ** 20))
    return ("working set %.1f MiB fits the 126 MB L2 by the nature of the "
            "configuration: consecutive time steps of one loop reuse it, "
            "nothing is flushed, and the roofline fraction says nothing about "
            "DRAM" % (5 * field / 2 ** 20))


def workload_config(args, p, timesteps):
    return {
        "workload": "%s: grid %s (extended, halo %d), space_order %d, %s "
                    "density, float32, %d time steps per step, 1 source, "
                    "%d receivers" % (
                        p["name"], "x".join(map(str, p["velocity"].shape)),
                        p["space_order"] // 2, p["space_order"],
                        "variable" if p.get("density") is not None else "constant",
                        timesteps, len(p["rec_offsets"]) - 1),
        "l2_policy": l2_policy(p),
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


def run_slab_leg(args, rank, world, dist, barrier, max_over_ranks, planes=None,
                 scaling="weak"):
    """One large 3D model split into z-slabs, one per GPU, ghost planes kept
    current on the device over NVLink (simwave_b200/slab.py).  Weak scaling:
    --slab-planes owned planes per GPU; strong scaling: `planes` = 1024 / N (C4
    proper, the 1040^3 grid).  Returns the "slab" object of the bench line
    (None on ranks other than 0)."""
    import workloads
    from simwave_b200 import slab
    planes = planes or args.slab_planes
    q = workloads.slab_3d(rank=rank, world=world,
                          planes_per_gpu=planes, n=args.slab_n,
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
    local = int(os.environ.get("LOCAL_RANK", "0"))
    with ClockSampler(local) as clocks:
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
        "value": value, "unit": "Gpts/s", "scaling": scaling,
        "ms_per_timestep": 1e3 * total / len(times) / T,
        "roofline_frac_per_gpu": value / world * bpp / peak,
        "clocks_rank0": clocks.summary(),
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


def set_hints(lib, **hints):
    """simwave_cuda_set_hint by name (include/simwave_cuda.h)."""
    codes = {"wavefield_in_zero": 1, "wavefield_out": 2, "model_resident": 3}
    lib.simwave_cuda_set_hint.argtypes = [ctypes.c_int, ctypes.c_longlong]
    for name, value in hints.items():
        if lib.simwave_cuda_set_hint(codes[name], int(value)) != 0:
            raise RuntimeError("hint %s refused" % name)


def run_survey_leg(args, rank, world, barrier, max_over_ranks):
    """C5: a multi-shot survey (64 shots x 512^3, so 8, a 512-receiver line per
    shot, 300 steps).  The model stays on the host once per rank (page-locked);
    every shot is one drop-in forward() call with its own tables, wavefield and
    traces; shot s runs on rank s mod world.  No data-path collective.  The
    survey driver tells the library what it knows about its own arrays
    (simwave_cuda_set_hint): `u` is zero on entry, only the traces are wanted
    back, and the model is the same for every shot.  Returns the "survey"
    object of the bench line."""
    import workloads
    from cuda_abi import core, cuda_forward, last_timing
    p = workloads.shot_3d(shot=0, timesteps=args.shot_timesteps)
    T = p["end_timestep"]
    pts = workloads.interior_points(p)
    host = dict(p)
    for key in ("velocity", "damp", "wavelet"):
        host[key] = pinned_like(host[key])
    host["u"] = pinned_like(p["u"])
    host["receivers"] = pinned_like(p["receivers"])
    mine = list(range(rank, args.shots, world))
    traces = {}
    device_seconds = []
    lib = core()

    def shoot(shot):
        q = workloads.reshoot(p, shot)
        for key in ("src_intervals", "src_values", "src_offsets",
                    "rec_intervals", "rec_values", "rec_offsets"):
            host[key] = q[key]
        host["receivers"][...] = 0
        cuda_forward(host)
        device_seconds.append(last_timing()["loop"])
        traces[shot] = float(np.abs(host["receivers"]).max())

    set_hints(lib, wavefield_in_zero=1, wavefield_out=2, model_resident=1000 + rank)
    try:
        if mine:
            shoot(mine[0])                  # warm-up: model upload, caches, clocks
        barrier()
        device_seconds.clear()
        t0 = time.perf_counter()
        for shot in mine:
            shoot(shot)
        seconds = time.perf_counter() - t0
        barrier()
    finally:
        set_hints(lib, wavefield_in_zero=0, wavefield_out=0, model_resident=0)
        lib.simwave_cuda_release_cache()
    total = max_over_ranks(seconds)
    device_total = max_over_ranks(sum(device_seconds))
    if rank != 0:
        return None
    work = args.shots * pts * T / 1e9
    return {"value": work / total, "unit": "Gpts/s",
            "device_value": work / device_total,
            "e2e_over_device": device_total / total,
            "shots": args.shots, "shots_per_second": args.shots / total,
            "seconds": total, "scaling": "strong",
            "config": {"workload": "survey of %d shots over %s" % (
                args.shots, workload_config(args, p, T)["workload"]),
                "parallelism": "shot s on rank s mod %d, one forward() per "
                               "shot, model arrays shared on the host" % world,
                "hints": "wavefield_in_zero, wavefield_out=none (traces only), "
                         "model_resident (one model upload per rank)",
                "note": "value = through forward() with host buffers, wall "
                        "clock (max over ranks); device_value = the same shots' "
                        "CUDA-event loop time"},
            "max_abs_trace_of_first_shot": traces.get(0)}


def run_2d_config(name, args, barrier):
    """C1 / C2 on one GPU: device-resident loop (plan API, CUDA events), the
    drop-in forward() with pageable host arrays, and the reference's
    cpu_openmp kernel on the whole run, for the "configs_2d" object."""
    import workloads
    from cuda_abi import cuda_forward
    from simwave_b200 import slab
    p = workloads.WORKLOADS[name]()
    T = p["end_timestep"]
    pts = workloads.interior_points(p)
    plan = slab.Plan(p)
    times = []
    for i in range(3 + max(2, args.steps)):
        plan.reset()
        t = plan.run(1, T)
        if i >= 3:
            times.append(t)
    launches = plan.launches()
    plan.destroy()
    dev = sum(times) / len(times)
    e2e = []
    for i in range(3):
        q = dict(p)
        q["u"] = np.zeros_like(p["u"])
        q["receivers"] = np.zeros_like(p["receivers"])
        barrier()
        t0 = time.perf_counter()
        cuda_forward(q)
        e2e.append(time.perf_counter() - t0)
    out = {
        "value": pts * T / dev / 1e9, "unit": "Gpts/s",
        "us_per_timestep": 1e6 * dev / T, "timesteps": T,
        "e2e": {"value": pts * T / min(e2e[1:]) / 1e9, "unit": "Gpts/s",
                "host_memory": "pageable (NumPy)",
                "h2d_bytes_per_step": int(sum(
                    p[k].nbytes for k in ("velocity", "damp", "wavelet", "src_values",
                                          "rec_values", "src_intervals",
                                          "rec_intervals"))),
                "d2h_bytes_per_step": int(p["u"].nbytes + p["receivers"].nbytes)},
        "gpu_launches_per_step": int(launches),
        "config": workload_config(args, p, T),
        "roofline_note": "frac vs HBM would be %.2f of the measured peak, but "
                         "the fields live in L2 (see l2_policy): the loop is "
                         "bound by per-step latency (one cooperative launch, a "
                         "grid barrier per step), not DRAM" % (
                             pts * T / dev * workloads.bytes_per_point(p) / 1e9
                             / measured_peak()[0]),
    }
    if not args.no_cpu:
        try:
            sec, kind, variant = cpu_reference_child(name, T)
            out["cpu_baseline"] = {
                "value": pts * T / sec / 1e9, "unit": "Gpts/s", "cores": host_threads(),
                "kind": "reference" if kind == "ref" else "port",
                "sample": "all %d time steps, %s build, OMP_PROC_BIND=true" % (T, variant)}
        except Exception as e:
            out["cpu_baseline"] = {"value": None, "sample": "failed: %s" % e}
    return out


def run_f64_config(args):
    """C3 in float64 -- the precision simwave's own benchmark script builds its
    model in (benchmark/overthrust_3D.py:82) -- on one GPU: device-resident
    loop (plan API, CUDA events) over a bounded number of time steps, against
    the 40 B per grid-point update of the float64 fields."""
    import workloads
    from simwave_b200 import slab
    T = args.f64_timesteps
    p = workloads.overthrust_3d(timesteps=T, dtype=np.float64)
    pts = workloads.interior_points(p)
    bpp = workloads.bytes_per_point(p)
    plan = slab.Plan(p)
    times = []
    for i in range(3):
        plan.reset()
        t = plan.run(1, T)
        if i >= 1:
            times.append(t)
    launches = plan.launches()
    plan.destroy()
    from cuda_abi import core
    core().simwave_cuda_release_cache()
    dev = sum(times) / len(times)
    peak, _ = measured_peak()
    return {
        "value": pts * T / dev / 1e9, "unit": "Gpts/s", "dtype": "f64",
        "ms_per_timestep": 1e3 * dev / T, "timesteps": T,
        "bytes_per_point": bpp,
        "roofline_frac": pts * T / dev * bpp / 1e9 / peak,
        "gpu_launches": int(launches),
        "config": workload_config(args, p, T),
    }


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
    def time_forward(host, calls, warm):
        times = []
        for i in range(warm + calls):
            host["u"][...] = 0
            host["receivers"][...] = 0
            barrier()
            t0 = time.perf_counter()
            cuda_forward(host)
            torch.cuda.synchronize()
            dt_wall = time.perf_counter() - t0
            if i >= warm:
                times.append(dt_wall)
        return times

    def abi_bytes(host, pinned):
        h2d = sum(host[k].nbytes for k in
                  ("velocity", "damp", "wavelet", "src_intervals", "src_values",
                   "src_offsets", "rec_intervals", "rec_values", "rec_offsets"))
        if host.get("density") is not None:
            h2d += host["density"].nbytes
        if pinned and host["u"].shape[0] == 3:
            # a page-locked three-slot wavefield is uploaded outright (cheaper
            # than scanning it for zeros on the host);
            # a pageable one is scanned and, being zero, never uploaded
            h2d += host["u"].nbytes
        return int(h2d), int(host["u"].nbytes + host["receivers"].nbytes)

    e2e = e2e_pageable = e2e_api = None
    if not args.no_e2e:
        host = dict(p)
        for key in ("velocity", "damp", "density", "wavelet"):
            if host.get(key) is not None:
                host[key] = pinned_like(host[key])
        host["u"] = pinned_like(p["u"])
        host["receivers"] = pinned_like(p["receivers"])
        h2d, d2h = abi_bytes(host, True)
        times = time_forward(host, args.steps, 1)
        tm = last_timing()
        e2e_total = max_over_ranks(sum(times))
        e2e = {"value": world * pts * T * len(times) / e2e_total / 1e9,
               "unit": "Gpts/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h,
               "seconds_per_step": e2e_total / len(times),
               "breakdown_last_call": tm,
               "host_memory": "pinned (torch pin_memory)",
               "call": "drop-in forward() C-ABI, every slot of u copied back"}
        del host
        out_keepalive.clear()

    if not args.no_e2e and world == 1:
        # the same call with the pageable NumPy arrays simwave's Solver passes
        host = dict(p)
        host["u"] = np.zeros_like(p["u"])
        host["receivers"] = np.zeros_like(p["receivers"])
        h2d, d2h = abi_bytes(host, False)
        times = time_forward(host, max(1, args.steps), 1)
        e2e_pageable = {"value": pts * T * len(times) / sum(times) / 1e9,
                        "unit": "Gpts/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h,
                        "seconds_per_step": sum(times) / len(times),
                        "breakdown_last_call": last_timing(),
                        "host_memory": "pageable (NumPy)",
                        "call": "drop-in forward() C-ABI, every slot of u copied back"}
        del host

    if not args.no_e2e and not args.no_api and world == 1 and \
            p["name"] in workloads._API_SPECS:
        # the call a simwave user makes: Solver.forward() of the public API,
        # front end included in the timed region (tables, kernel arguments,
        # halo stripping); the model objects are built once, as a user would
        import contextlib
        try:
            t0 = time.perf_counter()
            solver = workloads.api_solver(
                p["name"], timesteps=T if args.timesteps else None)
            build_s = time.perf_counter() - t0
            steps_api = solver.time_model.timesteps
            times = []
            for i in range(1 + max(1, args.steps)):
                barrier()
                t0 = time.perf_counter()
                with contextlib.redirect_stdout(sys.stderr):
                    u_api, rec_api = solver.forward()
                torch.cuda.synchronize()
                if i >= 1:
                    times.append(time.perf_counter() - t0)
            space = solver.space_model
            e2e_api = {
                "value": pts * steps_api * len(times) / sum(times) / 1e9,
                "unit": "Gpts/s", "seconds_per_step": sum(times) / len(times),
                "timesteps": steps_api,
                # model resident after the warm-up call: wavelet and tables go up,
                # one extended wavefield slot and the shot record come back
                "h2d_bytes_per_step": int(
                    solver.wavelet.values.nbytes + sum(
                        a.nbytes for acq in (solver.sources, solver.receivers)
                        for a in acq.interpolated_points_and_values)),
                "d2h_bytes_per_step": int(
                    np.prod(space.extended_shape) * u_api.itemsize + rec_api.nbytes),
                "breakdown_last_call": last_timing(),
                "model_build_seconds_not_timed": build_s,
                "host_memory": "pageable (NumPy)",
                "call": "simwave_b200.Solver.forward() (SpaceModel, TimeModel, "
                        "Source, Receiver, RickerWavelet as in the reference's "
                        "benchmark script); model resident on the device after "
                        "the warm-up call (SpaceModel.model_token), u known to "
                        "be zero on entry, only the returned wavefield slot and "
                        "the shot record copied back",
                "max_abs_wavefield": float(np.abs(u_api).max())}
            del solver, u_api, rec_api
        except Exception as e:       # an extra leg must not sink the bench line
            e2e_api = {"error": "%s: %s" % (type(e).__name__, e)}

    # ---- CPU baseline beside it (rank 0, N == 1 only) ------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sample = args.cpu_timesteps or max(1, int(3e9 // pts))
        try:
            s, kind, variant = cpu_reference_child(args.workload, sample, args.timesteps)
            cpu = {"value": pts * sample / s / 1e9, "unit": "Gpts/s",
                   "cores": host_threads(),
                   "kind": "reference" if kind == "ref" else "port",
                   "sample": "%d of %d time steps of the same workload (child process), "
                             "%s build, OMP_PROC_BIND=true" % (
                                 sample, p["full_timesteps"], variant)}
        except Exception as e:   # the baseline must not sink the bench line
            cpu = {"value": None, "unit": "Gpts/s", "cores": host_threads(),
                   "kind": "port", "sample": "failed: %s" % e}

    # ---- C5: multi-shot survey through forward() ------------------------------
    # (before the slab legs: their model generation on the GPU and their peer
    # mappings leave state behind that a survey process would not have)
    survey = None
    if args.shots > 0 and not args.no_survey:
        try:
            survey = run_survey_leg(args, rank, world, barrier, max_over_ranks)
        except Exception as e:
            if world > 1:
                raise
            survey = {"error": "%s: %s" % (type(e).__name__, e)}

    # ---- slab decomposition legs (C4-shaped) ---------------------------------
    slab_result = slab_strong = None
    if not args.no_slab:
        try:
            slab_result = run_slab_leg(args, rank, world, dist, barrier,
                                       max_over_ranks)
            if world > 1 and not args.no_slab_strong and 1024 % world == 0:
                slab_strong = run_slab_leg(args, rank, world, dist, barrier,
                                           max_over_ranks, planes=1024 // world,
                                           scaling="strong")
        except Exception as e:      # the extra leg must not sink the bench line
            if world > 1:
                raise               # ranks wait on each other: fail together
            slab_result = {"error": "%s: %s" % (type(e).__name__, e)}

    # ---- C1 / C2: the 2D configurations (N == 1) -------------------------------
    configs_2d = None
    if world == 1 and not args.no_2d:
        configs_2d = {}
        for name in ("readme_2d", "marmousi_2d"):
            try:
                configs_2d[name] = run_2d_config(name, args, barrier)
            except Exception as e:
                configs_2d[name] = {"error": "%s: %s" % (type(e).__name__, e)}

    # ---- C3 in float64, as the reference's benchmark script runs it (N == 1) ---
    c3_f64 = None
    if world == 1 and not args.no_f64 and p is not None and p.get("name") == "overthrust_3d":
        try:
            c3_f64 = run_f64_config(args)
        except Exception as e:
            c3_f64 = {"error": "%s: %s" % (type(e).__name__, e)}

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
            "e2e_pageable": e2e_pageable,
            "e2e_api": e2e_api,
            "slab": slab_result,
            "slab_strong": slab_strong,
            "survey": survey,
            "configs_2d": configs_2d,
            "c3_f64": c3_f64,
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
    if args.cpu_child:
        run_cpu_child(args)
        return
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
