#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29511 tools/slab_check.py > gpurun_out/slab_check.log 2>&1; echo "exit $?" >> gpurun_out/slab_check.log
tail -25 gpurun_out/slab_check.log
