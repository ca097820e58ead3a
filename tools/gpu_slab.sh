#!/bin/bash
# multi-GPU: slab parity check, then the default bench under torchrun (shot-parallel C3 + slab leg)
mkdir -p gpurun_out
NG=${NG:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tools/slab_check.py > gpurun_out/slab_check_n$NG.log 2>&1; echo "exit $?" >> gpurun_out/slab_check_n$NG.log
grep -v "^\*\|OMP_NUM\|^W1\|^$" gpurun_out/slab_check_n$NG.log | tail -12
if [ "$NG" = "2" ]; then
  timeout 600 python bench.py --slab-only > gpurun_out/slab_n1.json 2> gpurun_out/slab_n1.err; cut -c1-330 gpurun_out/slab_n1.json; tail -2 gpurun_out/slab_n1.err
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG --steps 2 --warmup 3 > gpurun_out/bench_n$NG.json 2> gpurun_out/bench_n$NG.err; cat gpurun_out/bench_n$NG.json; grep -i "error\|Traceback" gpurun_out/bench_n$NG.err | head -3
