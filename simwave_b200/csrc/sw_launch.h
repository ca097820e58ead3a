// Host-side launchers; each (type, dimension, density) variant is instantiated
// in its own translation unit (sw_inst.cu compiled 8 times) so the build
// parallelises.
#pragma once

#include "sw_common.h"

namespace sw {

// one thread per point, operands from global memory; any radius 1..10
template <typename T, int NDIM, bool VARDEN>
void launch_step_simple(int math, const StepArgs<T> &a, cudaStream_t stream);

}  // namespace sw
