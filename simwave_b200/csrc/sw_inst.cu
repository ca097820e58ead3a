// Instantiates the step kernels of ONE variant, chosen on the command line:
//   -DSW_T=float|double  -DSW_NDIM=2|3  -DSW_VARDEN=0|1
#include "sw_launch.h"
#include "sw_step_simple.cuh"
#if SW_NDIM == 2
#include <algorithm>

#include "sw_loop2d.cuh"
#include "sw_loop2d_resident.cuh"
#endif

namespace sw {

template <typename T, int NDIM, bool VARDEN, int R>
static void launch_simple_r(int math, const StepArgs<T> &a, cudaStream_t stream)
{
    const Grid &g = a.g;
    dim3 block(64, 4, 1);
    dim3 grid((g.nF - 2 * R + block.x - 1) / block.x, (g.nM - 2 * R + block.y - 1) / block.y,
              NDIM == 3 ? g.nS - 2 * R : 1);
    if (math == MATH_STRICT)
        step_simple_kernel<T, NDIM, VARDEN, R, MATH_STRICT><<<grid, block, 0, stream>>>(a);
    else
        step_simple_kernel<T, NDIM, VARDEN, R, MATH_FAST><<<grid, block, 0, stream>>>(a);
}

template <typename T, int NDIM, bool VARDEN>
void launch_step_simple(int math, const StepArgs<T> &a, cudaStream_t stream)
{
    switch (a.g.r) {
#define SW_CASE(R) case R: launch_simple_r<T, NDIM, VARDEN, R>(math, a, stream); break;
        SW_CASE(1) SW_CASE(2) SW_CASE(3) SW_CASE(4) SW_CASE(5)
        SW_CASE(6) SW_CASE(7) SW_CASE(8) SW_CASE(9) SW_CASE(10)
#undef SW_CASE
    default:
        throw Error("stencil radius " + std::to_string(a.g.r) + " not supported (1..10)");
    }
    SW_CUDA(cudaGetLastError());
}

template void launch_step_simple<SW_T, SW_NDIM, (SW_VARDEN != 0)>(int, const StepArgs<SW_T> &,
                                                                   cudaStream_t);

#if SW_NDIM == 2
template bool launch_loop2d<SW_T, (SW_VARDEN != 0)>(int, const LoopArgs<SW_T> &, cudaStream_t);
template bool loop2d_resident_tiling<SW_T, (SW_VARDEN != 0)>(int, const Grid &, int, int,
                                                             Loop2dTiling *);
template bool launch_loop2d_resident<SW_T, (SW_VARDEN != 0)>(int, const LoopArgs<SW_T> &,
                                                             const Loop2dTiling &, cudaStream_t);
#endif

}  // namespace sw
