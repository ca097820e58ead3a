#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_convergence.py -m gpu -x -q > gpurun_out/pytest_conv.log 2>&1; echo "exit $?" >> gpurun_out/pytest_conv.log; tail -3 gpurun_out/pytest_conv.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step3d_tiled -s 3 -c 1 -o gpurun_out/prof_c4 -f python tools/sweep.py --workload slab_3d --timesteps 6 --cfgs 0 --math fast --repeat 1 > gpurun_out/ncu_c4.log 2>&1
tail -2 gpurun_out/ncu_c4.log
SIMWAVE_CUDA_VERBOSE=1 timeout 900 python bench.py --no-slab --no-cpu > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; cat gpurun_out/bench_e2e.json
grep "simwave_b200: forward\|upload phases" gpurun_out/bench_e2e.err | tail -4
