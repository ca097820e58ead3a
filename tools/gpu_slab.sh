#!/bin/bash
mkdir -p gpurun_out
NG=${NG:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG > gpurun_out/bench_n$NG.json 2> gpurun_out/bench_n$NG.err; echo "exit $?"
cat gpurun_out/bench_n$NG.json; tail -5 gpurun_out/bench_n$NG.err
timeout 300 python bench.py --timesteps 300 --no-e2e --no-cpu > gpurun_out/bench_n1_short.json 2> gpurun_out/bench_n1_short.err; cat gpurun_out/bench_n1_short.json; tail -3 gpurun_out/bench_n1_short.err
