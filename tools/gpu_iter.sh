#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --timesteps 40 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step3d_tiled -s 30 -c 2 -o gpurun_out/prof_tiled_r4_fast -f python bench.py --timesteps 40 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
