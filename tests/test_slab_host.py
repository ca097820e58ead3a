"""
Host side of the slab decomposition (simwave_b200/slab.py), on CPU.

The partition logic -- plane ranges, ghost planes, clipped source / receiver
tables, boundary codes of inner faces, assembly of the results -- is checked by
running the CPU oracle as the per-slab "device": world_size ranks (gloo) each
step their own slab one time step at a time and exchange ghost planes after
every step, exactly the protocol the CUDA engine follows on the device.  The
assembled wavefield must be bit-identical to the oracle's single-domain run and
the summed traces equal within float rounding.
"""
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "oracle"))

import oracle  # noqa: E402
import problems  # noqa: E402
from simwave_b200 import slab  # noqa: E402


def test_split_planes_covers_the_interior():
    for nz, r, world in [(64, 4, 2), (100, 8, 3), (215, 4, 8), (1040, 8, 8)]:
        ranges = slab.split_planes(nz, r, world)
        assert ranges[0][0] == r and ranges[-1][1] == nz - r
        for (a, b), (c, d) in zip(ranges, ranges[1:]):
            assert b == c and b > a
        sizes = [b - a for a, b in ranges]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        slab.split_planes(40, 8, 4)


def test_partition_tables_add_up_to_the_global_tables():
    p = problems.make_problem(shape=(60, 30, 32), space_order=8, timesteps=5,
                              num_sources=5, num_receivers=9, seed=1,
                              multi_wavelet=True)
    world = 3
    for kind in ("src", "rec"):
        count = len(p[kind + "_offsets"]) - 1
        total = [np.zeros(60) for _ in range(count)]
        for rank in range(world):
            q, info = slab.partition(p, rank, world)
            assert len(q[kind + "_offsets"]) - 1 == count
            a = info["planes"][0]
            iv = q[kind + "_intervals"].reshape(count, 6)
            giv = p[kind + "_intervals"].reshape(count, 6)
            for i in range(count):
                zb, ze = int(iv[i, 0]), int(iv[i, 1])
                w = q[kind + "_values"][int(q[kind + "_offsets"][i]):
                                        int(q[kind + "_offsets"][i + 1])]
                total[i][a + zb:a + ze + 1] += w[:ze - zb + 1]
                # x / y windows and weights are untouched
                assert np.array_equal(iv[i, 2:], giv[i, 2:])
                rest = p[kind + "_values"][int(p[kind + "_offsets"][i]):
                                           int(p[kind + "_offsets"][i + 1])]
                nzw = int(giv[i, 1] - giv[i, 0]) + 1
                assert np.array_equal(w[ze - zb + 1:], rest[nzw:])
        for i in range(count):
            zb, ze = int(giv[i, 0]), int(giv[i, 1])
            g = p[kind + "_values"][int(p[kind + "_offsets"][i]):][:ze - zb + 1]
            expect = np.zeros(60)
            expect[zb:ze + 1] = g
            assert np.array_equal(total[i], expect)


def _slab_worker(rank, world, port, shape, order, density, steps, bc, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = problems.make_problem(shape=shape, space_order=order,
                                  density=density, timesteps=steps, bc=bc,
                                  nbl=((0, 3), (2, 2), (3, 2)), num_sources=3,
                                  num_receivers=8, src_radius=4, rec_radius=4,
                                  multi_wavelet=True, seed=7)
        q, info = slab.partition(p, rank, world)
        r = info["radius"]
        nS = q["velocity"].shape[0]
        for n in range(1, steps + 1):
            q["begin_timestep"] = q["end_timestep"] = n
            oracle.forward(q)
            nxt = q["u"][(n + 1) % 3]
            reqs = []
            if info["up"]:
                reqs.append(dist.isend(torch.from_numpy(nxt[r:2 * r].copy()), rank - 1))
                top = torch.empty_like(torch.from_numpy(nxt[:r]))
                reqs.append(dist.irecv(top, rank - 1))
            if info["down"]:
                reqs.append(dist.isend(torch.from_numpy(nxt[nS - 2 * r:nS - r].copy()),
                                       rank + 1))
                bot = torch.empty_like(torch.from_numpy(nxt[nS - r:]))
                reqs.append(dist.irecv(bot, rank + 1))
            for req in reqs:
                req.wait()
            if info["up"]:
                nxt[:r] = top.numpy()
            if info["down"]:
                nxt[nS - r:] = bot.numpy()
        rec = torch.from_numpy(q["receivers"].copy())
        dist.reduce(rec, 0)
        parts = [None] * world
        dist.all_gather_object(parts, (q["u"], info))
        if rank == 0:
            u = slab.assemble_wavefield([x[0] for x in parts],
                                        [x[1] for x in parts], shape[0])
            np.savez(out, u=u, receivers=rec.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,order,density,bc", [
    (2, (40, 24, 26), 4, False, (2, 1, 2, 1, 0, 2)),
    (2, (44, 22, 22), 8, True, (1, 2, 1, 1, 2, 2)),
    (3, (50, 20, 24), 4, False, (2, 2, 0, 1, 1, 0)),
])
def test_slab_protocol_matches_single_domain(world, shape, order, density, bc,
                                             tmp_path):
    import torch.multiprocessing as mp
    steps = 14
    out = str(tmp_path / "slab.npz")
    port = 29500 + (os.getpid() + world * 7 + order) % 2000
    mp.spawn(_slab_worker, args=(world, port, shape, order, density, steps,
                                 bc, out), nprocs=world, join=True)
    got = np.load(out)
    p = problems.make_problem(shape=shape, space_order=order, density=density,
                              timesteps=steps, bc=bc,
                              nbl=((0, 3), (2, 2), (3, 2)), num_sources=3,
                              num_receivers=8, src_radius=4, rec_radius=4,
                              multi_wavelet=True, seed=7)
    oracle.forward(p)
    assert np.abs(p["u"]).max() > 0
    assert np.array_equal(got["u"], p["u"])
    scale = np.abs(p["receivers"]).max()
    assert np.allclose(got["receivers"], p["receivers"], rtol=0,
                       atol=2e-6 * scale)
