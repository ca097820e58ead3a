#!/bin/bash
# final check of the round: GPU tests, smoke, 2D probe, default bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 120 python tools/loop2d_probe.py 2>&1 | grep -v "^$" | head -3
timeout 600 python bench.py --no-slab > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cat gpurun_out/bench_final.json
