#!/usr/bin/env python
"""
Multi-GPU parity check of the slab decomposition (run under torchrun, one rank
per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
        --master-addr 127.0.0.1 --master-port 29511 tools/slab_check.py

Every rank builds the same seeded global problem, takes its z-slab, runs the
time loop with device-side halo exchange; rank 0 assembles the wavefield, sums
the traces and compares with (a) the same problem run on one GPU through the
drop-in `forward` and (b) the CPU oracle.  In strict math mode (a) must be
bit-identical for the wavefield.
"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests"), os.path.join(REPO, "oracle")):
    sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import problems  # noqa: E402
from simwave_b200 import slab  # noqa: E402


def rel_l2(a, b):
    den = np.linalg.norm(b.astype(np.float64).ravel())
    return np.linalg.norm((a.astype(np.float64) - b).ravel()) / (den or 1.0)


CASES = [
    # shape, order, density, steps, bc
    ((80, 60, 150), 8, False, 40, (2, 1, 2, 1, 0, 2)),
    ((64, 40, 44), 4, True, 30, (1, 2, 1, 1, 2, 2)),
    ((150, 50, 70), 16, False, 25, (2, 1, 1, 1, 1, 1)),     # tall enough for 8 slabs
]


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ["SIMWAVE_CUDA_DEVICE"] = str(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    failures = 0

    def gather_bytes(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out

    for math, push in (("strict", "fused"), ("fast", "fused"), ("strict", "copy")):
        os.environ["SIMWAVE_CUDA_MATH"] = math
        os.environ["SIMWAVE_CUDA_SLAB_PUSH"] = push
        for shape, order, density, steps, bc in CASES:
            p = problems.make_problem(
                shape=shape, space_order=order, density=density,
                timesteps=steps, bc=bc, nbl=((0, 3), (2, 2), (3, 2)),
                num_sources=3, num_receivers=8, src_radius=4, rec_radius=4,
                multi_wavelet=True, seed=7)
            q, info = slab.partition(p, rank, world)
            plan = slab.Plan(q)
            slab.connect_neighbours(plan, rank, world, gather_bytes)
            dist.barrier()
            for rep in range(2):        # second pass exercises reset()
                plan.reset()
                torch.cuda.synchronize()
                dist.barrier()
                plan.run(1, steps)
            plan.download()
            rec = torch.from_numpy(q["receivers"].copy()).cuda()
            dist.reduce(rec, 0)
            parts = [None] * world
            dist.all_gather_object(parts, (q["u"], info))
            plan.destroy()
            if rank == 0:
                from cuda_abi import cuda_forward
                import oracle
                u = slab.assemble_wavefield([x[0] for x in parts],
                                            [x[1] for x in parts], shape[0])
                single = problems.clone(p)
                cuda_forward(single)
                cpu = problems.clone(p)
                oracle.forward(cpu)
                same = np.array_equal(u, single["u"])
                eu, er = rel_l2(u, cpu["u"]), rel_l2(rec.cpu().numpy(), cpu["receivers"])
                es = rel_l2(rec.cpu().numpy(), single["receivers"])
                ok = same and eu <= 1e-5 and er <= 1e-5
                failures += 0 if ok else 1
                print("%-6s/%-5s %s so%d %s T=%d on %d slabs: wavefield %s single-GPU; "
                      "vs CPU oracle rel-L2 u %.2e rec %.2e; rec vs single-GPU %.2e  %s"
                      % (math, push, "x".join(map(str, shape)), order,
                         "var" if density else "const", steps, world,
                         "bit-identical to" if same else "DIFFERS from",
                         eu, er, es, "OK" if ok else "FAIL"), flush=True)
            dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("slab_check:", "PASS" if failures == 0 else "FAIL (%d)" % failures)
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
