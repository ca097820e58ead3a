#!/bin/bash
# iteration pass: tiled-kernel tests (hang-guarded), full GPU tests, sweep
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "tiled" > gpurun_out/pytest_tiled.log 2>&1; echo "exit $?" >> gpurun_out/pytest_tiled.log
tail -15 gpurun_out/pytest_tiled.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
SIMWAVE_CUDA_VERBOSE=1 timeout 600 python tools/sweep.py --timesteps 60 --cfgs ${CFGS:-0,1,2,3,4,5,6,7} --math ${MATHS:-fast,strict} > gpurun_out/sweep_overthrust.txt 2>&1; cat gpurun_out/sweep_overthrust.txt
