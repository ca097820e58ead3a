#!/bin/bash
# multi-GPU: slab parity check (CHECK=1), then the default bench under torchrun
# (shot-parallel C3 + slab leg), as the driver launches it
mkdir -p gpurun_out
NG=${NG:-2}
if [ "${CHECK:-1}" = "1" ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tools/slab_check.py > gpurun_out/slab_check_n$NG.log 2>&1; echo "exit $?" >> gpurun_out/slab_check_n$NG.log
  grep -v "^\*\|OMP_NUM\|^W1\|^$" gpurun_out/slab_check_n$NG.log | tail -12
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG --steps 2 --warmup 3 > gpurun_out/bench_n$NG.json 2> gpurun_out/bench_n$NG.err; cat gpurun_out/bench_n$NG.json; grep -i "error\|Traceback" gpurun_out/bench_n$NG.err | head -3
