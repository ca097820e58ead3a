// Host side of the float64 tiled 3D kernel: one configuration per radius
// class.  Compiled once per radius (-DSW_RADIUS=1..10) like sw_tiled3d.cu.
#include <atomic>

#include "sw_launch.h"
#include "sw_step_tiled3d64.cuh"

namespace sw {

// R <= 5: 14 x 32 tile (256-byte rows), 7 consumer warps + the producer warp,
// two planes of u_cur and three stream stages in flight, 2 CTAs per SM.
// Larger radii: the ring of R+1+PF halo planes fills shared memory, one CTA
// per SM, shallower rings.
template <int R>
struct Cfg64 {
    static constexpr int TX = 16, TY = 14;
    static constexpr int PF = R <= 5 ? 2 : 1;
    static constexpr int PS = R <= 5 ? 3 : 2;
    static constexpr int MINB = R <= 5 ? 2 : 1;
    using TL = Tile3D64<R, TX, TY, PF, PS>;
};

template <int R>
static bool query64(TiledInfo *info)
{
    using C = Cfg64<R>;
    *info = {1, C::TX, C::TY, C::PF, C::PS, C::MINB, C::TL::SMEM_BYTES, C::TL::VW};
    return true;
}

template <int R>
static bool launch64(int math, const StepArgs<double> &a, const StepMaps &maps,
                     const unsigned char *qflags, int zChunk, cudaStream_t stream)
{
    using C = Cfg64<R>;
    using TL = typename C::TL;
    const Grid &g = a.g;
    dim3 grid((g.nF - 2 * R + TL::BY - 1) / TL::BY, (g.nM - 2 * R + TL::BX - 1) / TL::BX,
              (g.nS - 2 * R + zChunk - 1) / zChunk);
    auto kStrict = step3d_tiled64_kernel<R, C::TX, C::TY, C::PF, C::PS, MATH_STRICT, C::MINB>;
    auto kFast = step3d_tiled64_kernel<R, C::TX, C::TY, C::PF, C::PS, MATH_FAST, C::MINB>;
    auto k = (math == MATH_STRICT) ? kStrict : kFast;
    static std::atomic<unsigned long long> configured[2];
    int dev = 0;
    SW_CUDA(cudaGetDevice(&dev));
    std::atomic<unsigned long long> &mask = configured[math == MATH_STRICT];
    if (!(mask.load(std::memory_order_acquire) >> (dev & 63) & 1ull)) {
        SW_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     TL::SMEM_BYTES));
        mask.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
    k<<<grid, TL::THREADS, TL::SMEM_BYTES, stream>>>(a, maps, qflags, zChunk);
    SW_CUDA(cudaGetLastError());
    return true;
}

}  // namespace sw

#define SW_CAT2(a, b) a##b
#define SW_CAT(a, b) SW_CAT2(a, b)

namespace sw {
bool SW_CAT(tiled3d64_query_r, SW_RADIUS)(TiledInfo *info) { return query64<SW_RADIUS>(info); }
bool SW_CAT(tiled3d64_launch_r, SW_RADIUS)(int math, const StepArgs<double> &a,
                                           const StepMaps &maps, const unsigned char *qflags,
                                           int zChunk, cudaStream_t stream)
{
    return launch64<SW_RADIUS>(math, a, maps, qflags, zChunk, stream);
}
}  // namespace sw
