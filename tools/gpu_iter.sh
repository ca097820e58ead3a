#!/bin/bash
# kernel iteration: parity tests, tile sweeps on C3 and the C4-shaped slab, 2D step probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cuda_parity.py tests/test_convergence.py -m gpu -x -q > gpurun_out/pytest_parity.log 2>&1; echo "exit $?" >> gpurun_out/pytest_parity.log
tail -4 gpurun_out/pytest_parity.log
timeout 600 python tools/sweep.py --timesteps 60 --math fast --cfgs 0,1,2,3,4,5,6,7 > gpurun_out/sweep_overthrust.txt 2>&1; cat gpurun_out/sweep_overthrust.txt
timeout 600 python tools/sweep.py --workload slab_3d --timesteps 20 --cfgs 0,4,5,6,7,8 --math fast > gpurun_out/sweep_slab3d.txt 2>&1; cat gpurun_out/sweep_slab3d.txt
SIMWAVE_CUDA_LOOP2D_TRACE=1 timeout 300 python tools/loop2d_probe.py 2>&1 | grep -v "^$" | cut -c1-250
