#!/bin/bash
# One GPU pass for the judged evidence of a round (R=r02 ...): GPU tests, smoke,
# the default bench line with its wall time, the ncu launch list of the same
# command, full captures of the C3 and C4-shaped step kernels.  Results land
# in gpurun_out/; export summaries with tools/profile_export.py.
R=${R:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${R}_smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${R}_pytest_gpu.log
tail -3 gpurun_out/${R}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/${R}_smoke.txt
t0=$(date +%s)
SIMWAVE_CUDA_VERBOSE=1 timeout 1200 python bench.py > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err
echo "default bench wall $(( $(date +%s) - t0 )) s" | tee gpurun_out/${R}_bench_n1.wall
cat gpurun_out/${R}_bench_n1.json | cut -c1-1500
t0=$(date +%s)
timeout 900 python bench.py --impl reference > gpurun_out/${R}_bench_ref_n1.json 2> gpurun_out/${R}_bench_ref_n1.err
echo "reference arm wall $(( $(date +%s) - t0 )) s" | tee -a gpurun_out/${R}_bench_n1.wall
cat gpurun_out/${R}_bench_ref_n1.json | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_c3.csv python bench.py --timesteps 40 --steps 1 --warmup 1 --no-cpu --no-e2e --no-slab --no-survey --no-2d --no-api --no-f64 > gpurun_out/${R}_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step3d_tiled -s 30 -c 1 -o gpurun_out/${R}_prof_c3 -f python bench.py --timesteps 40 --steps 1 --warmup 1 --no-cpu --no-e2e --no-slab --no-survey --no-2d --no-api --no-f64 > gpurun_out/${R}_ncu_c3.log 2>&1
tail -2 gpurun_out/${R}_ncu_c3.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:step3d_tiled -s 3 -c 1 -o gpurun_out/${R}_prof_c4 -f python tools/sweep.py --workload slab_3d --kw planes_per_gpu=512 --timesteps 6 --cfgs 5 --math fast --repeat 1 --no-simple > gpurun_out/${R}_ncu_c4.log 2>&1
tail -2 gpurun_out/${R}_ncu_c4.log
ls -la gpurun_out | grep ${R}_ | tail -20
