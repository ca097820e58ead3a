"""
The measurement contract on the CPU side: `bench.py --impl reference` (the
reference's cpu_openmp kernel from oracle/_ref, or the port) prints ONE JSON
line with the agreed keys, and under torchrun only rank 0 works.  Also: the
product package never reaches into oracle/ (the oracle is test infrastructure).
"""
import json
import os
import re
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CONTRACT_KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup",
                 "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                 "dtype", "data", "config", "cpu_baseline", "e2e"}


def _run(cmd, **env):
    e = dict(os.environ, OMP_NUM_THREADS="4", **env)
    return subprocess.run(cmd, cwd=REPO, env=e, capture_output=True, text=True,
                          timeout=600)


def _json_lines(stdout):
    return [json.loads(line) for line in stdout.splitlines()
            if line.startswith("{")]


@pytest.mark.slow
def test_reference_arm_prints_the_contract_line():
    out = _run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1",
                "--warmup", "0", "--cpu-timesteps", "1",
                "--workload", "marmousi_2d"])
    assert out.returncode == 0, out.stderr[-2000:]
    lines = _json_lines(out.stdout)
    assert len(lines) == 1
    line = lines[0]
    assert CONTRACT_KEYS <= set(line)
    assert line["impl"] == "reference" and line["metric"] == "Gpts/s"
    assert line["value"] > 0 and line["higher_is_better"] is True
    assert line["vs_baseline"] is None and line["data"] == "synthetic"
    assert "marmousi_2d" in line["config"]["workload"]
    base = line["cpu_baseline"]
    assert base["kind"] in ("reference", "port") and base["cores"] >= 1
    assert base["value"] == line["value"] and "time steps" in base["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "Gpts/s",
                           "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.slow
def test_reference_arm_under_torchrun_runs_on_rank_zero_only():
    out = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                "--master-port", "29533", "bench.py", "--impl", "reference",
                "--gpus", "2", "--steps", "1", "--warmup", "0",
                "--cpu-timesteps", "1", "--workload", "readme_2d"])
    assert out.returncode == 0, out.stderr[-2000:]
    lines = _json_lines(out.stdout)
    assert len(lines) == 1 and lines[0]["n_gpus"] == 2


def test_product_never_touches_the_oracle():
    """Only tests/, smoke() and bench.py's CPU legs may use oracle/."""
    pattern = re.compile(r"\boracle\b")
    offenders = []
    for root, _, files in os.walk(os.path.join(REPO, "simwave_b200")):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h", ".c")) or name == "Makefile":
                path = os.path.join(root, name)
                with open(path, errors="replace") as f:
                    for number, text in enumerate(f, 1):
                        if pattern.search(text):
                            offenders.append("%s:%d" % (os.path.relpath(path, REPO), number))
    assert not offenders, offenders
    # and the shipped libraries link against nothing from oracle/
    lib = os.path.join(REPO, "simwave_b200", "lib", "libsimwave_b200.so")
    needed = subprocess.run(["readelf", "-d", lib], capture_output=True, text=True).stdout
    assert "oracle" not in needed and "ref_" not in needed


def test_cpu_arm_uses_every_host_core_even_under_torchrun(monkeypatch):
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm must not inherit it
    (round-1 SCALE records at N >= 2 ran the reference single-threaded)."""
    import ctypes
    sys.path.insert(0, REPO)
    import bench
    monkeypatch.setenv("OMP_NUM_THREADS", "1")
    gomp = ctypes.CDLL("libgomp.so.1")
    gomp.omp_set_num_threads(1)
    n = bench.set_cpu_threads()
    assert n == bench.host_threads() >= 1
    assert os.environ["OMP_NUM_THREADS"] == str(n)
    assert gomp.omp_get_max_threads() == n


def test_cuda_arm_imports_nothing_from_the_oracle():
    """tests/cuda_abi.py and tests/abi.py (the ctypes callers bench.py's GPU arm
    and smoke() use) are independent of oracle/."""
    for name in ("cuda_abi.py", "abi.py"):
        with open(os.path.join(REPO, "tests", name)) as f:
            text = f.read()
        assert not re.search(r"^\s*(import|from)\s+oracle\b", text, re.M), name


def test_cpu_baseline_runs_in_a_child_and_leaves_the_gpu_arm_unbound():
    """The GPU arm must not export OMP_PROC_BIND: libgomp would pin the main
    thread to one core at `import torch` and every helper thread of the CUDA
    library would inherit that mask.  Its cpu_baseline comes from a child
    process (`bench.py --cpu-child`), which does bind its OpenMP threads."""
    probe = ("import sys, os; sys.argv = ['bench.py']; sys.path.insert(0, %r); "
             "import bench; print(os.environ.get('OMP_PROC_BIND'), "
             "len(os.sched_getaffinity(0)) == bench.host_threads())" % REPO)
    env = {k: v for k, v in os.environ.items() if not k.startswith("OMP_")}
    out = subprocess.run([sys.executable, "-c", probe], cwd=REPO, env=env,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.split() == ["None", "True"]

    child = _run([sys.executable, "bench.py", "--cpu-child", "--workload", "readme_2d",
                  "--cpu-timesteps", "4"])
    assert child.returncode == 0, child.stderr[-2000:]
    (line,) = _json_lines(child.stdout)
    assert line["seconds"] > 0 and line["kind"] in ("ref", "port") and line["threads"] >= 1
