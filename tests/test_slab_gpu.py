"""
Slab decomposition on the device, under the driver's eyes: `world` processes
share cuda:0, each owns a z-slab of one seeded problem, ghost planes travel
through CUDA IPC peer mappings (fused peer stores or copy-engine pushes) with
device-side step flags.  The assembled wavefield must be BIT-IDENTICAL to the
same problem run as one plan through the drop-in forward(), strict and fast
math, and within the stated tolerance of the CPU oracle; traces are partial
sums added over ranks (exact when a window lies inside one slab, a different
association otherwise).

Cases put a source window across the cut, a Neumann top face, overlapping
sources with per-source wavelets, variable density and damping layers.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "oracle"))

import oracle  # noqa: E402
import problems  # noqa: E402
from cuda_abi import cuda_forward  # noqa: E402
from conftest import rel_l2  # noqa: E402
from simwave_b200 import slab  # noqa: E402

pytestmark = pytest.mark.gpu


def run_slabs(case, world, tmp_path):
    path = tmp_path / "case.json"
    path.write_text(json.dumps(case))
    procs = [subprocess.Popen([sys.executable, os.path.join(REPO, "tests", "slab_worker.py"),
                               str(rank), str(world), str(path), str(tmp_path)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for rank in range(world)]
    logs = []
    for proc in procs:
        try:
            out, _ = proc.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            for other in procs:
                other.kill()
            raise
        logs.append(out)
    assert all(proc.returncode == 0 for proc in procs), "\n".join(logs)[-4000:]
    parts, infos, traces = [], [], None
    for rank in range(world):
        with np.load(tmp_path / ("out_%d.npz" % rank)) as z:
            parts.append(z["u"])
            infos.append({"planes": tuple(z["planes"]), "owned": tuple(z["owned"])})
            traces = z["receivers"].astype(np.float64) if traces is None \
                else traces + z["receivers"]
    return parts, infos, traces


def problem_kwargs(shape, order, density, steps, bc, cut_planes):
    """A seeded problem with one source window across every cut plane."""
    r = order // 2
    src = [(c + 0.4, shape[1] * 0.45, shape[2] * 0.55) for c in cut_planes]
    src.append((r + 2.6, shape[1] * 0.5, shape[2] * 0.4))      # next to the top face
    return dict(shape=list(shape), space_order=order, density=density, timesteps=steps,
                bc=list(bc), nbl=[[0, 3], [2, 2], [3, 2]], num_sources=len(src),
                src_positions=src, num_receivers=10, src_radius=3, rec_radius=4,
                multi_wavelet=True, seed=7)


CASES = [
    # shape, order, density, steps, bc (Neumann top face), world
    ((84, 44, 72), 8, False, 24, (2, 1, 2, 1, 0, 2), 2),
    ((70, 40, 76), 4, True, 20, (2, 2, 1, 1, 2, 2), 3),
    ((100, 40, 70), 16, True, 12, (2, 1, 1, 1, 1, 1), 2),
]


@pytest.mark.parametrize("math,push", [("strict", "fused"), ("fast", "fused"),
                                       ("strict", "copy")])
@pytest.mark.parametrize("shape,order,density,steps,bc,world", CASES)
def test_slabs_on_one_device_equal_the_single_plan_run(shape, order, density, steps, bc,
                                                       world, math, push, tmp_path,
                                                       monkeypatch):
    r = order // 2
    cuts = [lo for lo, _ in slab.split_planes(shape[0], r, world)][1:]
    kwargs = problem_kwargs(shape, order, density, steps, bc, cuts)
    case = {"math": math, "push": push, "problem": kwargs,
            "passes": 2 if math == "fast" else 1}
    parts, infos, traces = run_slabs(case, world, tmp_path)
    u = slab.assemble_wavefield(parts, infos, shape[0])

    monkeypatch.setenv("SIMWAVE_CUDA_MATH", math)
    p = problems.make_problem(**kwargs)
    single = problems.clone(p)
    cuda_forward(single)
    assert np.abs(single["u"]).max() > 0
    assert np.array_equal(u, single["u"]), \
        "slab wavefield differs from the single-plan run: rel-L2 %.3e" % rel_l2(u, single["u"])
    assert rel_l2(traces, single["receivers"]) <= 2e-6

    cpu = problems.clone(p)
    oracle.forward(cpu)
    assert rel_l2(u, cpu["u"]) <= 1e-5
    assert rel_l2(traces, cpu["receivers"]) <= 1e-5


@pytest.mark.parametrize("math,push", [("strict", "fused"), ("fast", "fused"),
                                       ("fast", "copy")])
def test_float64_slabs_equal_the_single_plan_run(math, push, tmp_path, monkeypatch):
    """The float64 tiled kernel (sw_step_tiled3d64.cuh) under slab
    decomposition: its ghost-plane stores, fused and copy push, two slabs."""
    shape, order, world = (84, 58, 70), 8, 2
    cuts = [lo for lo, _ in slab.split_planes(shape[0], order // 2, world)][1:]
    kwargs = problem_kwargs(shape, order, False, 18, (2, 1, 2, 1, 0, 2), cuts)
    kwargs["dtype"] = "float64"
    case = {"math": math, "push": push, "problem": kwargs, "passes": 1}
    parts, infos, traces = run_slabs(case, world, tmp_path)
    u = slab.assemble_wavefield(parts, infos, shape[0])
    monkeypatch.setenv("SIMWAVE_CUDA_MATH", math)
    p = problems.make_problem(**kwargs)
    assert p["u"].dtype == np.float64
    single = problems.clone(p)
    cuda_forward(single)
    assert np.abs(single["u"]).max() > 0
    assert np.array_equal(u, single["u"])
    assert rel_l2(traces, single["receivers"]) <= 1e-12


def _device_count():
    from cuda_abi import core
    return core().simwave_cuda_device_count()


@pytest.mark.parametrize("math", ["strict", "fast"])
@pytest.mark.parametrize("shape,order,density,steps,bc,world", CASES)
def test_forward_over_several_devices_of_one_process(shape, order, density, steps, bc, world,
                                                     math, monkeypatch):
    """SIMWAVE_CUDA_NGPUS: the drop-in forward() itself cuts the problem into
    slabs, one device and one host thread each (no torchrun, no user-side
    reduction).  Needs as many GPUs as slabs; the wavefield must be
    bit-identical to the single-device call, the traces agree to rounding."""
    if _device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    monkeypatch.setenv("SIMWAVE_CUDA_MATH", math)
    r = order // 2
    cuts = [lo for lo, _ in slab.split_planes(shape[0], r, world)][1:]
    p = problems.make_problem(**problem_kwargs(shape, order, density, steps, bc, cuts))
    single = problems.clone(p)
    cuda_forward(single)
    monkeypatch.setenv("SIMWAVE_CUDA_NGPUS", str(world))
    multi = problems.clone(p)
    cuda_forward(multi)
    assert np.abs(single["u"]).max() > 0
    assert np.array_equal(multi["u"], single["u"])
    assert rel_l2(multi["receivers"], single["receivers"]) <= 2e-6
    # through the public API as well: nothing but the environment changes
    multi2 = problems.clone(p)
    multi2["begin_timestep"], multi2["end_timestep"] = 1, steps - 3
    part = problems.clone(p)
    part["end_timestep"] = steps - 3
    monkeypatch.delenv("SIMWAVE_CUDA_NGPUS")
    cuda_forward(part)
    monkeypatch.setenv("SIMWAVE_CUDA_DEVICES", ",".join(str(d) for d in range(world)))
    cuda_forward(multi2)
    assert np.array_equal(multi2["u"], part["u"])
