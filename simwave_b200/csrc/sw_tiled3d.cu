// Host side of the tiled 3D kernel: configuration table and launch.
// Compiled once per radius (-DSW_RADIUS=1..10) so the build parallelises.
#include <atomic>

#include "sw_launch.h"
#include "sw_step_tiled3d.cuh"

// planes per trip of the unrolled plane loop
#ifndef SW_UNROLL
#define SW_UNROLL 3
#endif

namespace sw {

template <int R, int PM, int TX, int TY, int PF, int PS, int MINB, bool VARDEN, int UNR>
static void launch_cfg(int math, const StepArgs<float> &a, const StepMaps &maps,
                       const unsigned char *qflags, int zChunk, cudaStream_t stream)
{
    using TLS = Tile3D<R, PM, TX, TY, PF, PS, VARDEN, VARDEN>;      // STRICT layout
    using TLF = Tile3D<R, PM, TX, TY, PF, PS, VARDEN, false>;       // FAST layout
    using TL = TLS;
    const int smemBytes = (math == MATH_STRICT) ? TLS::SMEM_BYTES : TLF::SMEM_BYTES;
    const Grid &g = a.g;
    dim3 grid((g.nF - 2 * R + TL::BY - 1) / TL::BY, (g.nM - 2 * R + TL::BX - 1) / TL::BX,
              (g.nS - 2 * R + zChunk - 1) / zChunk);
    auto kStrict = step3d_tiled_kernel<R, PM, TX, TY, PF, PS, MATH_STRICT, MINB, VARDEN, UNR>;
    auto kFast = step3d_tiled_kernel<R, PM, TX, TY, PF, PS, MATH_FAST, MINB, VARDEN, UNR>;
    auto k = (math == MATH_STRICT) ? kStrict : kFast;
    // the shared-memory opt-in is per device: one bit per ordinal (atomic: one
    // host thread per device may be launching, simwave_cuda_set_slab_devices)
    static std::atomic<unsigned long long> configured[2];
    int dev = 0;
    SW_CUDA(cudaGetDevice(&dev));
    std::atomic<unsigned long long> &mask = configured[math == MATH_STRICT];
    if (!(mask.load(std::memory_order_acquire) >> (dev & 63) & 1ull)) {
        SW_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     smemBytes));
        mask.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
    k<<<grid, TL::THREADS, smemBytes, stream>>>(a, maps, qflags, zChunk);
}

#define SW_CFG(ID, PM, TX, TY, PF, PS, MINB) SW_CFGU(ID, PM, TX, TY, PF, PS, MINB, SW_UNROLL)
#define SW_CFGU(ID, PM, TX, TY, PF, PS, MINB, UNR)                                         \
    case ID:                                                                               \
        if (query) {                                                                       \
            *query = {PM, TX, TY, PF, PS, MINB,                                            \
                      (math == MATH_STRICT)                                                \
                          ? Tile3D<R, PM, TX, TY, PF, PS, VARDEN, VARDEN>::SMEM_BYTES      \
                          : Tile3D<R, PM, TX, TY, PF, PS, VARDEN, false>::SMEM_BYTES};     \
            return true;                                                                   \
        }                                                                                  \
        launch_cfg<R, PM, TX, TY, PF, PS, MINB, VARDEN, UNR>(math, a, *maps, qflags,       \
                                                             zChunk, stream);              \
        return true;

// PM, TX, TY = points per thread along M, thread columns, thread rows
// (tile = TY*PM rows x 4*TX columns); PF / PS = u_cur planes / stream
// stages in flight; MINB = CTAs per SM the register budget is held to.
// Thread counts are mostly kept at multiples of 4 warps (register allocation
// granularity): 7 consumer warps + the producer warp, etc.
template <int R, bool VARDEN>
static bool dispatch(int cfg, TiledInfo *query, int math, const StepArgs<float> &a,
                     const StepMaps *maps, const unsigned char *qflags, int zChunk,
                     cudaStream_t stream)
{
    if constexpr (!VARDEN && R <= 5) {
        switch (cfg) {
            SW_CFG(0, 1, 16, 14, 3, 3, 2)    // 14 x 64 tile, 7+1 warps
            SW_CFG(1, 2, 16, 14, 2, 3, 1)    // 28 x 64 tile, 7+1 warps
            SW_CFG(2, 1, 32, 7, 3, 3, 2)     // 7 x 128 tile, 7+1 warps
            SW_CFG(3, 2, 32, 7, 2, 3, 1)     // 14 x 128 tile, 7+1 warps
            SW_CFG(4, 1, 16, 22, 3, 3, 1)    // 22 x 64 tile, 11+1 warps
            SW_CFG(5, 1, 16, 6, 3, 3, 4)     // 6 x 64 tile, 3+1 warps
            SW_CFG(6, 1, 16, 16, 3, 3, 2)    // 16 x 64 tile, 8+1 warps
            SW_CFG(7, 1, 16, 30, 2, 3, 1)    // 30 x 64 tile, 15+1 warps
        default: return false;
        }
    } else if constexpr (!VARDEN) {
        switch (cfg) {
            SW_CFG(0, 1, 16, 14, 2, 3, 1)    // 14 x 64 tile, 7+1 warps
            SW_CFG(1, 1, 16, 22, 2, 2, 1)    // 22 x 64 tile, 11+1 warps
            SW_CFG(2, 1, 32, 7, 2, 3, 1)     // 7 x 128 tile, 7+1 warps
            SW_CFG(3, 1, 16, 30, 1, 2, 1)    // 30 x 64 tile, 15+1 warps
        default: return false;
        }
    } else if constexpr (R <= 5) {
        switch (cfg) {
            SW_CFG(0, 1, 16, 16, 2, 2, 2)    // 16 x 64 tile, 8+1 warps
            SW_CFG(1, 1, 16, 14, 2, 2, 2)    // 14 x 64 tile, 7+1 warps
            SW_CFG(2, 1, 16, 30, 2, 2, 1)    // 30 x 64 tile, 15+1 warps
            SW_CFG(3, 1, 16, 22, 2, 3, 1)    // 22 x 64 tile, 11+1 warps
        default: return false;
        }
    } else {
        switch (cfg) {
            SW_CFG(0, 1, 16, 22, 1, 2, 1)    // 22 x 64 tile, 11+1 warps
            SW_CFG(1, 1, 16, 14, 2, 2, 1)    // 14 x 64 tile, 7+1 warps
            SW_CFG(2, 1, 16, 16, 1, 2, 1)    // 16 x 64 tile, 8+1 warps
            SW_CFG(3, 1, 32, 7, 1, 2, 1)     // 7 x 128 tile, 7+1 warps
            SW_CFG(4, 1, 16, 22, 2, 2, 1)    // 22 x 64 tile, deeper u_cur prefetch
            SW_CFG(5, 1, 16, 22, 1, 3, 1)    // 22 x 64 tile, deeper stream prefetch
            SW_CFG(6, 1, 16, 26, 1, 2, 1)    // 26 x 64 tile, 13+1 warps
            SW_CFG(7, 1, 16, 20, 2, 3, 1)    // 20 x 64 tile, both rings deeper (FAST layout only)
            SW_CFG(8, 1, 16, 18, 3, 3, 1)    // 18 x 64 tile, 3 u_cur planes in flight (FAST only)
            SW_CFGU(9, 1, 16, 22, 1, 3, 1, 4)    // configuration 5, plane loop unrolled 4x
            SW_CFGU(10, 1, 16, 22, 1, 3, 1, 6)   // ... 6x (fewer queue moves, longer body)
            SW_CFGU(11, 1, 16, 26, 1, 2, 1, 6)   // configuration 6, 6x
            SW_CFGU(12, 1, 16, 22, 1, 3, 1, 2)   // configuration 5, 2x
        default: return false;
        }
    }
}

}  // namespace sw

// one exported pair per radius
#define SW_CAT2(a, b) a##b
#define SW_CAT(a, b) SW_CAT2(a, b)

namespace sw {
bool SW_CAT(tiled3d_query_r, SW_RADIUS)(int cfg, bool varden, int math, TiledInfo *info)
{
    StepArgs<float> dummy{};
    return varden
               ? dispatch<SW_RADIUS, true>(cfg, info, math, dummy, nullptr, nullptr, 1, nullptr)
               : dispatch<SW_RADIUS, false>(cfg, info, math, dummy, nullptr, nullptr, 1, nullptr);
}
bool SW_CAT(tiled3d_launch_r, SW_RADIUS)(int cfg, bool varden, int math,
                                         const StepArgs<float> &a, const StepMaps &maps,
                                         const unsigned char *qflags, int zChunk,
                                         cudaStream_t stream)
{
    const bool ok =
        varden ? dispatch<SW_RADIUS, true>(cfg, nullptr, math, a, &maps, qflags, zChunk, stream)
               : dispatch<SW_RADIUS, false>(cfg, nullptr, math, a, &maps, qflags, zChunk, stream);
    if (ok)
        SW_CUDA(cudaGetLastError());
    return ok;
}
}  // namespace sw
