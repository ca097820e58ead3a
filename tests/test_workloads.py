"""
The synthetic workloads bench.py measures (workloads.py): shapes and metadata
of the five BASELINE.json configurations at reduced size, and the property the
weak-scaling slab leg rests on -- every rank builds only its own z-slab, and
the slabs are windows of ONE global model.
"""
import numpy as np

import workloads
from simwave_b200 import slab


def test_slab_workload_ranks_are_windows_of_one_global_model():
    n, planes, world, order = 112, 24, 3, 16
    r = order // 2
    whole = workloads.slab_3d(rank=0, world=1, planes_per_gpu=world * planes,
                              n=n, space_order=order, timesteps=3)
    nz = world * planes + 2 * r
    assert whole["velocity"].shape == (nz, n, n)
    assert whole["damp"].max() > 0 and whole["damp"][r:-r, 50, 50].min() == 0
    for rank, (lo, hi) in enumerate(slab.split_planes(nz, r, world)):
        part = workloads.slab_3d(rank=rank, world=world, planes_per_gpu=planes,
                                 n=n, space_order=order, timesteps=3)
        a, b = lo - r, hi + r
        assert part["global_shape"] == (nz, n, n)
        assert part["owned_planes"] == hi - lo == planes
        for key in ("velocity", "density", "damp"):
            assert np.array_equal(part[key], whole[key][a:b]), (rank, key)
        assert part["slab_up"] == int(rank > 0)
        assert part["slab_down"] == int(rank < world - 1)
        # inner faces carry no boundary condition of their own
        assert part["bc"][0] == (0 if rank > 0 else whole["bc"][0])
        assert part["bc"][1] == (0 if rank < world - 1 else whole["bc"][1])
        assert part["dt"] == whole["dt"]


def test_named_configurations_have_the_documented_shapes():
    c1 = workloads.readme_2d(timesteps=4)
    assert c1["velocity"].shape == (517, 517) and c1["space_order"] == 4
    c2 = workloads.marmousi_2d(timesteps=4)
    assert c2["velocity"].shape == (429, 1849) and c2["space_order"] == 8
    assert len(c2["rec_offsets"]) - 1 == 1700 and c2["damp"].max() > 0
    assert workloads.interior_points(c2) == 421 * 1841
    assert workloads.bytes_per_point(c2) == 20
    vd = workloads.slab_3d(planes_per_gpu=16, n=100, timesteps=2)
    assert workloads.bytes_per_point(vd) == 24


def test_bench_reads_the_profiled_traffic():
    import bench
    t = bench.profiled_traffic("overthrust_3d")
    pts = 207 * 801 * 801                                 # interior points of C3
    assert t is not None and 16 * pts < t < 20 * pts      # between compulsory and algorithmic
    assert bench.profiled_traffic("no_such_workload") is None


def test_survey_shots_share_the_model_and_move_the_acquisition():
    base = workloads.shot_3d(shot=0, n=48, timesteps=4)
    for shot in (1, 3):
        fresh = workloads.shot_3d(shot=shot, n=48, timesteps=4)
        again = workloads.reshoot(base, shot)
        for key in ("src_intervals", "src_values", "src_offsets",
                    "rec_intervals", "rec_values", "rec_offsets"):
            assert np.array_equal(fresh[key], again[key]), key
        assert again["velocity"] is base["velocity"] and again["damp"] is base["damp"]
        assert again["u"] is not base["u"] and not again["u"].any()
        assert again["shot"] == shot
    assert not np.array_equal(workloads.reshoot(base, 1)["src_intervals"],
                              base["src_intervals"])


def test_api_solvers_hand_the_kernel_the_arrays_of_the_abi_builders(monkeypatch):
    """bench.py's e2e_api leg (workloads.api_solver: the reference's benchmark
    scripts through simwave_b200's public API) runs the same problem as the
    ABI-level builders: every kernel argument agrees bit for bit."""
    from simwave_b200.kernel.backend import middleware
    seen = {}

    def capture(self, **kwargs):
        seen.update(kwargs)
        return kwargs['u_full'], kwargs['shot_record']
    monkeypatch.setattr(middleware.Middleware, "_exec_forward", capture)
    for name in ("readme_2d", "marmousi_2d"):
        p = workloads.WORKLOADS[name]()
        solver = workloads.api_solver(name)
        u, rec = solver.forward()
        assert rec.shape == p["receivers"].shape
        assert np.array_equal(seen["velocity_model"], p["velocity"])
        assert np.array_equal(seen["damping_mask"], p["damp"])
        assert np.array_equal(seen["wavelet"], p["wavelet"])
        assert seen["dt"] == p["dt"] and seen["end_timestep"] == p["end_timestep"]
        assert np.array_equal(seen["second_order_fd_coefficients"], p["coeff2"])
        for a, b in (("src_points_interval", "src_intervals"),
                     ("src_points_values", "src_values"),
                     ("rec_points_interval", "rec_intervals"),
                     ("rec_points_values", "rec_values"),
                     ("rec_points_values_offset", "rec_offsets")):
            assert np.array_equal(seen[a], p[b]), (name, a)
        assert list(seen["boundary_condition"]) == list(p["bc"])
    short = workloads.api_solver("readme_2d", timesteps=40)
    assert short.time_model.timesteps in (40, 41)
