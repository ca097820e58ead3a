#!/bin/bash
# compute-sanitizer memcheck over a small, fast subset of the GPU parity tests:
# every kernel family (tiled 3D constant / variable density, plain kernels, both
# 2D loops, sources / receivers / boundaries, snapshot drain) on grids of a few
# thousand points, through the drop-in `forward`.  memcheck only: racecheck does
# not model mbarrier / TMA completion and flags every ring slot of the tiled
# kernel.  Writes gpurun_out/<R>_sanitizer.log; exit code = the sanitizer's.
R=${R:-r02}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
K='test_strict_mode_bit_identical_to_oracle or test_tiled_kernel_every_configuration or test_persistent_2d_loop_every_radius or test_u_saving or test_sources_in_the_halo or test_tiled_variable_density_matches_oracle'
SIMWAVE_CUDA_CACHE=0 timeout 3000 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "$K" > gpurun_out/${R}_sanitizer.log 2>&1
rc=$?
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/${R}_sanitizer.log | tail -8
echo "sanitizer exit $rc" | tee -a gpurun_out/${R}_sanitizer.log
exit $rc
