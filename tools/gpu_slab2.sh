#!/bin/bash
# slab experiments: parity check of the early-publish mode, then slab-only bench with late and early halo publish
mkdir -p gpurun_out
NG=${NG:-2}
SLAB_CHECK_MODES=early timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 tools/slab_check.py > gpurun_out/slab_check_early_n$NG.log 2>&1; echo "exit $?" >> gpurun_out/slab_check_early_n$NG.log
grep -v "^\*\|OMP_NUM\|^W1\|^$" gpurun_out/slab_check_early_n$NG.log | tail -5
for mode in late early; do
SIMWAVE_CUDA_SLAB_PUBLISH=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $NG --slab-only --steps 3 > gpurun_out/slab_n${NG}_$mode.json 2> gpurun_out/slab_n${NG}_$mode.err; echo $mode; cut -c1-200 gpurun_out/slab_n${NG}_$mode.json; grep -i "error\|Traceback" gpurun_out/slab_n${NG}_$mode.err | head -3
done
