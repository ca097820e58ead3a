#!/bin/bash
# One GPU pass for the judged evidence: GPU tests, smoke, the default bench line,
# the ncu launch list and the full captures of the C3 and C4-shaped step kernels.
# Results land in gpurun_out/; export the summaries with tools/profile_export.py.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
SIMWAVE_CUDA_VERBOSE=1 timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
grep "simwave_b200: forward\|upload phases" gpurun_out/bench_default.err | tail -4
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --timesteps 40 --steps 1 --warmup 1 --no-cpu --no-e2e --no-slab > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step3d_tiled -s 30 -c 1 -o gpurun_out/prof_c3 -f python bench.py --timesteps 40 --steps 1 --warmup 1 --no-cpu --no-e2e --no-slab > gpurun_out/ncu_c3.log 2>&1
tail -2 gpurun_out/ncu_c3.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step3d_tiled -s 3 -c 1 -o gpurun_out/prof_c4 -f python tools/sweep.py --workload slab_3d --timesteps 6 --cfgs 5 --math fast --repeat 1 > gpurun_out/ncu_c4.log 2>&1
tail -2 gpurun_out/ncu_c4.log
ls -la gpurun_out
