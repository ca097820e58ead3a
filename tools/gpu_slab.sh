#!/bin/bash
mkdir -p gpurun_out
NG=${NG:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tools/slab_check.py > gpurun_out/slab_check.log 2>&1; echo "exit $?" >> gpurun_out/slab_check.log
grep -v "^\*\|OMP_NUM" gpurun_out/slab_check.log | tail -14
timeout 300 python bench.py --slab-only > gpurun_out/slab_n1.json 2> gpurun_out/slab_n1.err; cut -c1-160 gpurun_out/slab_n1.json; tail -2 gpurun_out/slab_n1.err
for n in 2 4 8; do
  if [ $n -le $NG ]; then
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --slab-only > gpurun_out/slab_n$n.json 2> gpurun_out/slab_n$n.err; cut -c1-160 gpurun_out/slab_n$n.json; grep -i "error\|Traceback" gpurun_out/slab_n$n.err | head -3
  fi
done
