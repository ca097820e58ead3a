#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "variable_density or default_math or fast_math" > gpurun_out/pytest_tiled.log 2>&1; echo "exit $?" >> gpurun_out/pytest_tiled.log
tail -12 gpurun_out/pytest_tiled.log
SIMWAVE_CUDA_VERBOSE=1 timeout 600 python tools/sweep.py --workload slab_3d --timesteps 20 --cfgs 0,1,2,3 --math fast,strict > gpurun_out/sweep_slab3d_so16_vd.txt 2>&1; cat gpurun_out/sweep_slab3d_so16_vd.txt
timeout 600 python tools/parity_report.py > gpurun_out/parity_fast.txt 2>&1; cat gpurun_out/parity_fast.txt
