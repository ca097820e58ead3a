#!/bin/bash
# state check: GPU tests, default bench with the upload/teardown phase log, 2D workload
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/nvidia-smi.txt 2>&1
nproc >> gpurun_out/nvidia-smi.txt; free -g >> gpurun_out/nvidia-smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
SIMWAVE_CUDA_VERBOSE=1 timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
grep simwave_b200 gpurun_out/bench_default.err | tail -12
SIMWAVE_CUDA_VERBOSE=1 timeout 600 python bench.py --workload marmousi_2d --no-slab --no-cpu > gpurun_out/bench_marmousi.json 2> gpurun_out/bench_marmousi.err; cat gpurun_out/bench_marmousi.json
grep simwave_b200 gpurun_out/bench_marmousi.err | tail -4
