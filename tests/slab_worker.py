"""
One rank of a slab-decomposed run, as its own process (tests/test_slab_gpu.py
starts `world` of them on ONE GPU: CUDA IPC works between processes that share
a device, so the peer stores, the wait / publish kernels and the descriptor
exchange are exercised without a multi-GPU box).

    python tests/slab_worker.py <rank> <world> <case.json> <dir>

The ranks meet through files in <dir>: desc_<rank>.bin (the 512-byte slab
descriptor), ready_<rank> (barrier before the time loop), out_<rank>.npz.
"""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "tests")):
    sys.path.insert(0, p)

import problems  # noqa: E402
from simwave_b200 import slab  # noqa: E402


def wait_for(paths, timeout=120.0):
    t0 = time.time()
    while not all(os.path.exists(p) for p in paths):
        if time.time() - t0 > timeout:
            raise SystemExit("timed out waiting for %s" % paths)
        time.sleep(0.01)


def publish(path, data=b"1"):
    with open(path + ".tmp", "wb") as f:
        f.write(data)
    os.rename(path + ".tmp", path)


def main():
    rank, world = int(sys.argv[1]), int(sys.argv[2])
    with open(sys.argv[3]) as f:
        case = json.load(f)
    where = sys.argv[4]
    os.environ["SIMWAVE_CUDA_DEVICE"] = "0"
    os.environ["SIMWAVE_CUDA_MATH"] = case["math"]
    os.environ["SIMWAVE_CUDA_SLAB_PUSH"] = case["push"]
    p = problems.make_problem(**case["problem"])
    q, info = slab.partition(p, rank, world)
    plan = slab.Plan(q)

    def gather(b):
        publish(os.path.join(where, "desc_%d.bin" % rank), b)
        paths = [os.path.join(where, "desc_%d.bin" % k) for k in range(world)]
        wait_for(paths)
        out = []
        for path in paths:
            with open(path, "rb") as f:
                out.append(f.read())
        return out
    slab.connect_neighbours(plan, rank, world, gather)
    for rep in range(case.get("passes", 1)):      # a second pass exercises reset()
        plan.reset()
        publish(os.path.join(where, "ready_%d_%d" % (rep, rank)))
        wait_for([os.path.join(where, "ready_%d_%d" % (rep, k)) for k in range(world)])
        plan.run(1, p["end_timestep"])
    plan.download()
    np.savez(os.path.join(where, "out_%d.npz" % rank), u=q["u"], receivers=q["receivers"],
             planes=np.array(info["planes"]), owned=np.array(info["owned"]))
    # neighbours may still be reading my buffers through their mappings
    publish(os.path.join(where, "done_%d" % rank))
    wait_for([os.path.join(where, "done_%d" % k) for k in range(world)])
    plan.destroy()


if __name__ == "__main__":
    main()
