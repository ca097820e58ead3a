#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cuda_parity.py tests/test_convergence.py -m gpu -x -q > gpurun_out/pytest_parity.log 2>&1; echo "exit $?" >> gpurun_out/pytest_parity.log
tail -6 gpurun_out/pytest_parity.log
timeout 600 python bench.py --workload marmousi_2d --no-slab --no-cpu > gpurun_out/bench_marmousi.json 2> gpurun_out/bench_marmousi.err; cut -c1-260 gpurun_out/bench_marmousi.json
timeout 600 python bench.py --workload readme_2d --no-slab --no-cpu > gpurun_out/bench_readme.json 2>/dev/null; cut -c1-260 gpurun_out/bench_readme.json
