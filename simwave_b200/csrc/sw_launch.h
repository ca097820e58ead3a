// Host-side launchers; each (type, dimension, density) variant is instantiated
// in its own translation unit (sw_inst.cu compiled 8 times) so the build
// parallelises.
#pragma once

#include "sw_common.h"

namespace sw {

// one thread per point, operands from global memory; any radius 1..10
template <typename T, int NDIM, bool VARDEN>
void launch_step_simple(int math, const StepArgs<T> &a, cudaStream_t stream);

// persistent 2D time loop (sw_loop2d.cuh): the whole range [begin, end] of time
// steps in one cooperative launch; false = not available, use per-step launches
template <typename T>
struct LoopArgs {
    StepArgs<T> a;            // prev / cur / next are taken from `slot`
    T *slot[3];               // the three rotating wavefield slots
    PointTables<T> src, rec;
    const T *wavelet;
    int waveletCount, srcMode, srcMaxPoints;
    int fuseSources;          // sources added by the thread that owns the cell
    int srcLoM, srcHiM, srcLoF, srcHiF;   // bounding box of all source windows
    T *recOut;                // [wavelet_size][rec.count]
    long long begin, end;     // time steps, inclusive
    unsigned *barrier;        // grid barrier counter, zero at launch
    unsigned long long *trace;   // phase time stamps of CTA 0 (development aid) or nullptr
    // tile-resident loop (loop2d_resident_kernel) only: one CTA per tile of
    // tm x tf interior points for the whole range of time steps
    int tm, tf, tilesM, tilesF;
    unsigned *flags;          // [tilesM * tilesF] steps completed per tile, zero at launch
    const int *recStart;      // [tiles + 1] receivers grouped by the tile that owns the
    const int *recIndex;      //   first cell of their window
};
template <typename T, bool VARDEN>
bool launch_loop2d(int math, const LoopArgs<T> &L, cudaStream_t stream);

// Tile-resident variant: every CTA keeps its tile of the wavefields and of the
// model in shared memory across the time loop and exchanges only halo strips
// with its neighbours (per-tile step flags instead of a grid barrier).
struct Loop2dTiling {
    int tm, tf, tilesM, tilesF;
    size_t smemBytes;
};
// picks the tiling for grid g (false: this problem does not fit -- too many
// tiles for one co-resident wave, or not enough shared memory); a tile is at
// least minTm x minTf points (the largest receiver window, so that a window
// reaches no further than a neighbouring tile)
template <typename T, bool VARDEN>
bool loop2d_resident_tiling(int math, const Grid &g, int minTm, int minTf, Loop2dTiling *out);
template <typename T, bool VARDEN>
bool launch_loop2d_resident(int math, const LoopArgs<T> &L, const Loop2dTiling &tiling,
                            cudaStream_t stream);

// ---- tiled 3D kernel (float32, constant density) ---------------------------
struct TiledInfo {
    int pm, tx, ty, pf, ps; // points per thread along M, thread columns / rows, prefetch depths
    int minBlocks;          // CTAs per SM the kernel was compiled for
    int smemBytes;
    int vw = 4;             // points per thread along F (4 floats / 2 doubles: 16 bytes)
    int tileM() const { return ty * pm; }
    int tileF() const { return tx * vw; }
};

}  // namespace sw

#include <cuda.h>

namespace sw {
// tensor maps of one time step: u_cur with halo; u_prev, c0, q halo-free
struct StepMaps {
    CUtensorMap cur, prev, c0, q;
    CUtensorMap rho, frF, frM, frS;     // variable density only
    // planes by which the producer warp runs ahead with L2 prefetches of every
    // input tile (cp.async.bulk.prefetch.tensor; 0 = none, at most 32)
    int prefetch;
    // Sources added by the thread that owns the cell, right after its stencil
    // value and before the boundary-aware store (the reference's own order:
    // section 1, section 2, section 3 of the loop): used when every window lies
    // among the interior points and the sources are few.  srcLo / srcHi: the
    // bounding box of all windows, (S,M,F).
    int srcFused;
    PointTables<float> src;
    const float *wavelet;
    int waveletCount;
    long long step;
    int srcLo[3], srcHi[3];
};
#define SW_DECL_TILED(R)                                                              \
    bool tiled3d_query_r##R(int cfg, bool varden, int math, TiledInfo *info);                   \
    bool tiled3d_launch_r##R(int cfg, bool varden, int math, const StepArgs<float> &a, \
                             const StepMaps &maps, const unsigned char *qflags,       \
                             int zChunk, cudaStream_t stream);
SW_DECL_TILED(1) SW_DECL_TILED(2) SW_DECL_TILED(3) SW_DECL_TILED(4) SW_DECL_TILED(5)
SW_DECL_TILED(6) SW_DECL_TILED(7) SW_DECL_TILED(8) SW_DECL_TILED(9) SW_DECL_TILED(10)
#undef SW_DECL_TILED
// float64 (sw_step_tiled3d64.cuh)
#define SW_DECL_TILED64(R)                                                            \
    bool tiled3d64_query_r##R(bool varden, int math, TiledInfo *info);                \
    bool tiled3d64_launch_r##R(bool varden, int math, const StepArgs<double> &a,      \
                               const StepMaps &maps, const unsigned char *qflags,     \
                               int zChunk, cudaStream_t stream);
SW_DECL_TILED64(1) SW_DECL_TILED64(2) SW_DECL_TILED64(3) SW_DECL_TILED64(4) SW_DECL_TILED64(5)
SW_DECL_TILED64(6) SW_DECL_TILED64(7) SW_DECL_TILED64(8) SW_DECL_TILED64(9) SW_DECL_TILED64(10)
#undef SW_DECL_TILED64
}  // namespace sw
