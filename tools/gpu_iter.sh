#!/bin/bash
# kernel iteration: parity tests, then tile sweeps on C3 and the C4-shaped slab
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -x -q > gpurun_out/pytest_parity.log 2>&1; echo "exit $?" >> gpurun_out/pytest_parity.log
tail -6 gpurun_out/pytest_parity.log
timeout 600 python tools/sweep.py --timesteps 60 --math fast --cfgs 0,1,2,3,4,5,6,7 > gpurun_out/sweep_overthrust.txt 2>&1; cat gpurun_out/sweep_overthrust.txt
timeout 600 python tools/sweep.py --workload slab_3d --timesteps 20 --cfgs 0,1,2,3 --math fast > gpurun_out/sweep_slab3d.txt 2>&1; cat gpurun_out/sweep_slab3d.txt
timeout 300 python tools/sweep.py --timesteps 60 --math strict --cfgs 6 > gpurun_out/sweep_overthrust_strict.txt 2>&1; cat gpurun_out/sweep_overthrust_strict.txt
