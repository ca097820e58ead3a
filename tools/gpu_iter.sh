#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/parity_report.py > gpurun_out/parity_fast.txt 2>&1; cat gpurun_out/parity_fast.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
SIMWAVE_CUDA_VERBOSE=1 timeout 600 python tools/sweep.py --timesteps 60 --cfgs 0,6,7 --math fast > gpurun_out/sweep_overthrust.txt 2>&1; cat gpurun_out/sweep_overthrust.txt
