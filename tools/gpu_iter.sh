#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step3d_tiled -s 4 -c 1 -o gpurun_out/prof_tiled_r8_vd -f python tools/sweep.py --workload slab_3d --timesteps 8 --cfgs 0 --math fast --repeat 0 > gpurun_out/ncu_vd.log 2>&1
tail -3 gpurun_out/ncu_vd.log
