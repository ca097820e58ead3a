// Host side of the float64 tiled 3D kernel: one configuration per radius
// class and density variant.  Compiled once per radius (-DSW_RADIUS=1..10) like sw_tiled3d.cu.
#include <atomic>

#include "sw_launch.h"
#include "sw_step_tiled3d64.cuh"

namespace sw {

// R <= 5: 14 x 32 tile (256-byte rows), 7 consumer warps + the producer warp,
// two planes of u_cur and three stream stages in flight, 2 CTAs per SM.
// Larger radii: the ring of R+1+PF halo planes fills shared memory, one CTA
// per SM, shallower rings.
// Variable density adds three (STRICT: four) stream tiles per stage: one CTA
// per SM, two stream stages.
template <int R, bool VARDEN>
struct Cfg64 {
    static constexpr int TX = 16, TY = 14;
    static constexpr int PF = R <= 5 ? 2 : 1;
    static constexpr int PS = (R <= 5 && !VARDEN) ? 3 : 2;
    static constexpr int MINB = (R <= 5 && !VARDEN) ? 2 : 1;
    using TLS = Tile3D64<R, TX, TY, PF, PS, VARDEN, VARDEN>;     // STRICT layout
    using TLF = Tile3D64<R, TX, TY, PF, PS, VARDEN, false>;      // FAST layout
};

template <int R, bool VARDEN>
static bool query64(int math, TiledInfo *info)
{
    using C = Cfg64<R, VARDEN>;
    *info = {1, C::TX, C::TY, C::PF, C::PS, C::MINB,
             math == MATH_STRICT ? C::TLS::SMEM_BYTES : C::TLF::SMEM_BYTES, C::TLS::VW};
    return true;
}

template <int R, bool VARDEN>
static bool launch64(int math, const StepArgs<double> &a, const StepMaps &maps,
                     const unsigned char *qflags, int zChunk, cudaStream_t stream)
{
    using C = Cfg64<R, VARDEN>;
    using TL = typename C::TLS;
    const int smemBytes = math == MATH_STRICT ? C::TLS::SMEM_BYTES : C::TLF::SMEM_BYTES;
    const Grid &g = a.g;
    dim3 grid((g.nF - 2 * R + TL::BY - 1) / TL::BY, (g.nM - 2 * R + TL::BX - 1) / TL::BX,
              (g.nS - 2 * R + zChunk - 1) / zChunk);
    auto kStrict =
        step3d_tiled64_kernel<R, C::TX, C::TY, C::PF, C::PS, MATH_STRICT, C::MINB, VARDEN>;
    auto kFast = step3d_tiled64_kernel<R, C::TX, C::TY, C::PF, C::PS, MATH_FAST, C::MINB, VARDEN>;
    auto k = (math == MATH_STRICT) ? kStrict : kFast;
    static std::atomic<unsigned long long> configured[2];
    int dev = 0;
    SW_CUDA(cudaGetDevice(&dev));
    std::atomic<unsigned long long> &mask = configured[math == MATH_STRICT];
    if (!(mask.load(std::memory_order_acquire) >> (dev & 63) & 1ull)) {
        SW_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smemBytes));
        mask.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
    k<<<grid, TL::THREADS, smemBytes, stream>>>(a, maps, qflags, zChunk);
    SW_CUDA(cudaGetLastError());
    return true;
}

}  // namespace sw

#define SW_CAT2(a, b) a##b
#define SW_CAT(a, b) SW_CAT2(a, b)

namespace sw {
bool SW_CAT(tiled3d64_query_r, SW_RADIUS)(bool varden, int math, TiledInfo *info)
{
    return varden ? query64<SW_RADIUS, true>(math, info) : query64<SW_RADIUS, false>(math, info);
}
bool SW_CAT(tiled3d64_launch_r, SW_RADIUS)(bool varden, int math, const StepArgs<double> &a,
                                           const StepMaps &maps, const unsigned char *qflags,
                                           int zChunk, cudaStream_t stream)
{
    return varden ? launch64<SW_RADIUS, true>(math, a, maps, qflags, zChunk, stream)
                  : launch64<SW_RADIUS, false>(math, a, maps, qflags, zChunk, stream);
}
}  // namespace sw
