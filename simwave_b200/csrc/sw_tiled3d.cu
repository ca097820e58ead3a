// Host side of the tiled 3D kernel: tensor maps, configuration table, launch.
// Compiled once per radius (-DSW_RADIUS=1..10) so the build parallelises.
#include "sw_launch.h"
#include "sw_step_tiled3d.cuh"

namespace sw {

template <int R, int PM, int TX, int TY, int PF, int MINB>
static void launch_cfg(int math, const StepArgs<float> &a, const CUtensorMap &map, int zChunk,
                       cudaStream_t stream)
{
    using TL = Tile3D<R, PM, TX, TY, PF>;
    const Grid &g = a.g;
    dim3 grid((g.nF - 2 * R + TL::BY - 1) / TL::BY, (g.nM - 2 * R + TL::BX - 1) / TL::BX,
              (g.nS - 2 * R + zChunk - 1) / zChunk);
    auto kStrict = step3d_tiled_kernel<R, PM, TX, TY, PF, MATH_STRICT, MINB>;
    auto kFast = step3d_tiled_kernel<R, PM, TX, TY, PF, MATH_FAST, MINB>;
    auto k = (math == MATH_STRICT) ? kStrict : kFast;
    static bool configured[2] = {false, false};
    if (!configured[math == MATH_STRICT]) {
        SW_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     TL::SMEM_BYTES));
        configured[math == MATH_STRICT] = true;
    }
    k<<<grid, TL::THREADS, TL::SMEM_BYTES, stream>>>(a, map, zChunk);
}

#define SW_CFG(ID, PM, TX, TY, PF, MINB)                                                   \
    case ID:                                                                               \
        if (query) { *query = {PM, TX, TY, PF, Tile3D<R, PM, TX, TY, PF>::SMEM_BYTES}; return true; } \
        launch_cfg<R, PM, TX, TY, PF, MINB>(math, a, *map, zChunk, stream);                \
        return true;

template <int R>
static bool dispatch(int cfg, TiledInfo *query, int math, const StepArgs<float> &a,
                     const CUtensorMap *map, int zChunk, cudaStream_t stream)
{
    if constexpr (R <= 5) {
        switch (cfg) {
            SW_CFG(0, 2, 16, 8, 2, 3)     // 16 x 64 tile, 128 threads, <= 168 regs
            SW_CFG(1, 2, 16, 16, 2, 1)    // 32 x 64 tile, 256 threads, <= 255 regs
            SW_CFG(2, 1, 16, 16, 2, 2)    // 16 x 64 tile, 256 threads, <= 128 regs
            SW_CFG(3, 2, 32, 8, 2, 1)     // 16 x 128 tile, 256 threads, <= 255 regs
            SW_CFG(4, 1, 32, 8, 2, 2)     // 8 x 128 tile, 256 threads
            SW_CFG(5, 1, 32, 16, 2, 1)    // 16 x 128 tile, 512 threads
            SW_CFG(6, 1, 16, 8, 2, 4)     // 8 x 64 tile, 128 threads
        default: return false;
        }
    } else {
        switch (cfg) {
            SW_CFG(0, 1, 16, 16, 1, 2)    // 16 x 64 tile, 256 threads
            SW_CFG(1, 1, 32, 8, 1, 2)     // 8 x 128 tile, 256 threads
            SW_CFG(2, 1, 32, 16, 1, 1)    // 16 x 128 tile, 512 threads
        default: return false;
        }
    }
}

}  // namespace sw

// one exported pair per radius
#define SW_CAT2(a, b) a##b
#define SW_CAT(a, b) SW_CAT2(a, b)

namespace sw {
bool SW_CAT(tiled3d_query_r, SW_RADIUS)(int cfg, TiledInfo *info)
{
    StepArgs<float> dummy{};
    return dispatch<SW_RADIUS>(cfg, info, 0, dummy, nullptr, 1, nullptr);
}
bool SW_CAT(tiled3d_launch_r, SW_RADIUS)(int cfg, int math, const StepArgs<float> &a,
                                         const CUtensorMap &map, int zChunk, cudaStream_t stream)
{
    const bool ok = dispatch<SW_RADIUS>(cfg, nullptr, math, a, &map, zChunk, stream);
    if (ok)
        SW_CUDA(cudaGetLastError());
    return ok;
}
}  // namespace sw
