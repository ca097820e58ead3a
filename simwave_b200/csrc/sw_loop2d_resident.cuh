// Tile-resident time-loop kernel for 2D problems.
//
// The persistent loop of sw_loop2d.cuh still fetches every operand of every
// step from L2 and closes every step with a grid-wide barrier: a Marmousi-sized
// step is ~2 us of L2 round trips plus ~1.6 us of barrier.  A 2D model of that
// size (0.8 M points; u_prev, u_cur, c0, q: 13 MB) fits in the shared memory
// of the 148 SMs (33 MB).  Here every CTA owns ONE tile of tm x tf interior
// points for the whole range of time steps and keeps its part of the problem
// on chip:
//
//   * shared memory holds the tile of u_cur and of u_prev with a ring of R
//     (RP along F) halo cells, and the tile of c0, q (and rho).  u_next is
//     written over u_prev in place -- a point's previous value is read by
//     nobody but the point itself -- so the two wavefield tiles just swap roles;
//   * a step is: stencil from shared memory -> the new values go to the tile
//     in shared memory AND to the global u_next slot (the global slots stay
//     exactly what the per-step kernels would have written: boundary mirrors,
//     the final wavefields, the cells the receivers sample);
//   * no grid barrier: a CTA publishes "step n done" in its own flag word
//     (st.release.gpu), waits for the flags of its (up to 8) neighbouring tiles
//     (ld.acquire.gpu) and then fetches only the halo ring of the new field
//     from the neighbours' global stores (L2).  Tiles drift apart by at most
//     one step per tile of distance;
//   * receivers are sampled from the global u_cur slot by the CTA that owns the
//     first cell of their window (tiles are at least one window wide, so every
//     window cell belongs to this tile or to one of the 8 neighbours whose
//     flags were awaited); sources are added by the thread that owns the cell,
//     as in the other loop kernel.
//
// Overwriting a global slot is safe with the three rotating slots: tile X
// writes slot (n+1)%3 in step n; the last readers of what it held (u^{n-2})
// were the neighbours' halo fetches at the end of step n-3 and their receivers
// at the start of step n-2, and X cannot start step n before every neighbour
// has published step n-1.
//
// Same device functions, same arithmetic, same order of section 1 / 2 / 3 of
// the reference loop as the other kernels (2d/wave.c:113-464): bit-identical
// to the per-step launches in either math mode.
//
// MEASURED (B200, round 2, profiles/r02_loop2d_probe.txt): NOT faster than the
// grid-barrier loop -- Marmousi-shaped C2 9.6 us per step against 6.2.  The
// stencil from shared memory takes 2.2 us of the step, but publishing a step
// costs a gpu-scope release (MEMBAR.ALL.GPU: 2-3 us on this part even for a
// single tile with nothing to wait for), then the neighbours' polls and the
// ring fetch are two more L2 round trips in series (ncu: 46 % of the warp
// samples wait at the barrier behind the flag poll, 12 % behind the ring
// fetch).  The grid-barrier loop pays one such release per step and overlaps
// its operand loads over 24 warps.  Kept as an opt-in
// (SIMWAVE_CUDA_LOOP2D=resident) with its parity tests; what would make it
// win is an exchange that needs no gpu-scope fence (cluster-scope barriers
// over distributed shared memory), not a faster stencil.
#pragma once

#include "sw_loop2d.cuh"

namespace sw {

// neighbourhood of a point inside a shared-memory tile (one point per thread)
template <typename T>
struct TileNeighbours {
    const T *p;
    int pitch;
    __device__ __forceinline__ T C() const { return p[0]; }
    __device__ __forceinline__ T F(int k) const { return p[k]; }
    __device__ __forceinline__ T M(int k) const { return p[k * pitch]; }
    __device__ __forceinline__ T M1(int k) const { return M(k); }
    __device__ __forceinline__ T S(int) const { return T(0); }
};

// what store_with_boundaries leaves in the cell of interior point (m,f) itself
template <typename T>
__device__ __forceinline__ T own_cell_value(const StepArgs<T> &a, int m, int f, T val)
{
    if (!a.fuse_bc)
        return val;
    const int r = a.g.r;
    const bool z = ((a.bc[4] == 1) & (f == r)) | ((a.bc[5] == 1) & (f == a.g.nF - r - 1)) |
                   ((a.bc[2] == 1) & (m == r)) | ((a.bc[3] == 1) & (m == a.g.nM - r - 1));
    return z ? T(0) : val;
}

// four elements from global memory through L2 (another SM wrote them)
__device__ __forceinline__ void load4_cg(const float *p, float out[4])
{
    const float4 v = __ldcg(reinterpret_cast<const float4 *>(p));
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
__device__ __forceinline__ void load4_cg(const double *p, double out[4])
{
    const double2 a = __ldcg(reinterpret_cast<const double2 *>(p));
    const double2 b = __ldcg(reinterpret_cast<const double2 *>(p + 2));
    out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y;
}

template <typename T, bool VARDEN, int R, int MATH>
__global__ void __launch_bounds__((Loop2dShape<T, R, VARDEN>::RESIDENT_THREADS), 1)
loop2d_resident_kernel(const __grid_constant__ LoopArgs<T> L)
{
    // strips of four points (float32); float64 takes the one-point-per-thread path
    constexpr int ROWS = Loop2dShape<T, R, VARDEN>::VW == 4 ? Loop2dShape<T, R, VARDEN>::ROWS : 0;
    constexpr int RP = (R + 3) / 4 * 4;
    const StepArgs<T> &a = L.a;
    const Grid &g = a.g;
    const int lastF = g.nF - R - 1, lastM = g.nM - R - 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nwarps = blockDim.x >> 5;

    // ---- my tile -------------------------------------------------------------
    const int tile = blockIdx.x;
    const int tM = tile / L.tilesF, tF = tile % L.tilesF;
    const int m0t = R + tM * L.tm, f0t = R + tF * L.tf;
    const int tmv = min(L.tm, lastM + 1 - m0t);          // valid rows / columns
    const int tfv = min(L.tf, lastF + 1 - f0t);
    const int PS = L.tf + 2 * RP;                        // shared row pitch
    const int rowsS = L.tm + 2 * R;
    const int fieldElems = rowsS * PS;

    extern __shared__ __align__(16) unsigned char smemRaw[];
    T *sCur = reinterpret_cast<T *>(smemRaw);
    T *sPrev = sCur + fieldElems;
    T *sC0 = sPrev + fieldElems;
    T *sQ = sC0 + fieldElems;
    T *sRho = sQ + fieldElems;                           // VARDEN only

    // cell (m,f) of the grid lives at tile[(m - mTop) * PS + (f - fLeft)]
    const int mTop = m0t - R, fLeft = f0t - RP;
    Grid gs = g;
    gs.pitch = PS;
    gs.planeStride = 0;

    // first column of the window that exists in a row of the pitched layout,
    // one past the last: [-lpad, pitch - lpad)
    const int fLimit = (int)g.pitch - g.lpad;

    // copy the window rows [mA, mB) x 4-element chunks [cA, cB) (chunk c = columns
    // f0t - RP + 4c ..) from a global field into a shared tile
    auto fetch = [&](T *tileS, const T *field, int mA, int mB, int cA, int cB) {
        const int nc = cB - cA, total = (mB - mA) * nc;
        for (int i = tid; i < total; i += blockDim.x) {
            const int m = mA + i / nc, c = cA + i % nc;
            const int f = f0t - RP + 4 * c;
            T v[4] = {T(0), T(0), T(0), T(0)};
            if (f + 3 < fLimit)
                load4_cg(field + g.at(0, m, f), v);
            store4(tileS + (m - mTop) * PS + 4 * c, v);
        }
    };
    const int mBot = m0t + tmv + R;                      // window rows [mTop, mBot)
    const int chunksAll = PS / 4, chunksHalo = RP / 4;
    // first chunk right of my last valid column (a partial tile keeps the
    // domain's own halo columns inside its column range)
    const int chunkRight = (RP + tfv) / 4;
    auto fetch_all = [&](T *tileS, const T *field) { fetch(tileS, field, mTop, mBot, 0, chunksAll); };
    // the ring only: R rows above and below my columns, RP columns left and right
    // of my rows (a star stencil never reads the corners)
    auto fetch_ring = [&](T *tileS, const T *field) {
        fetch(tileS, field, mTop, m0t, chunksHalo, chunksAll - chunksHalo);
        fetch(tileS, field, m0t + tmv, mBot, chunksHalo, chunksAll - chunksHalo);
        fetch(tileS, field, m0t, m0t + tmv, 0, chunksHalo);
        fetch(tileS, field, m0t, m0t + tmv, chunkRight, chunksAll);
    };

    fetch_all(sCur, L.slot[L.begin % 3]);
    fetch_all(sPrev, L.slot[(L.begin - 1) % 3]);
    fetch_all(sC0, a.c0);
    fetch_all(sQ, a.q);
    if (VARDEN)
        fetch_all(sRho, a.rho);
    __syncthreads();

    // neighbouring tiles whose flags I wait for (lanes 0..7 of warp 0)
    int nbTile = -1;
    if (tid < 8) {
        const int k = tid < 4 ? tid : tid + 1;           // skip the centre of the 3x3
        const int dm = k / 3 - 1, df = k % 3 - 1;
        const int nm = tM + dm, nf = tF + df;
        if (nm >= 0 && nm < L.tilesM && nf >= 0 && nf < L.tilesF)
            nbTile = nm * L.tilesF + nf;
    }
    const int recLo = L.recStart ? L.recStart[tile] : 0;
    const int recHi = L.recStart ? L.recStart[tile + 1] : 0;

    int stamp = 0;
    auto mark = [&]() {
        if (L.trace && blockIdx.x == 0 && tid == 0 && stamp < 256) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            L.trace[stamp++] = t;
        }
    };

    // strips of ROWS x 4 points, dealt to the threads round robin; a thread
    // walks its strips with increments instead of dividing by the row length
    const int stripsPerRow = L.tf / 4;
    const int strip0Row = tid / stripsPerRow, strip0Col = tid % stripsPerRow;
    const int stripDRow = (int)blockDim.x / stripsPerRow, stripDCol = (int)blockDim.x % stripsPerRow;

    for (long long n = L.begin; n <= L.end; n++) {
        const T *curG = L.slot[n % 3];
        T *nextG = L.slot[(n + 1) % 3];

        mark();
        // receivers of this tile: trace row n-1, from the global slot
        for (int i = recLo + warp; i < recHi; i += nwarps) {
            const int rec = L.recIndex[i];
            const T sum = (MATH == MATH_STRICT)
                              ? receiver_sample<T, 2>(g, curG, L.rec, rec, lane)
                              : receiver_sample_fast<T, 2>(g, curG, L.rec, rec, lane);
            if (lane == 0)
                L.recOut[(n - 1) * L.rec.count + rec] = sum;
        }

        mark();
        if constexpr (ROWS > 0) {
            const int stripRows = (tmv + ROWS - 1) / ROWS;
            for (int srow = strip0Row, scol = strip0Col; srow < stripRows;) {
                const int f0 = f0t + 4 * scol;
                const int m0 = m0t + srow * ROWS;
                srow += stripDRow;
                scol += stripDCol;
                if (scol >= stripsPerRow) {
                    scol -= stripsPerRow;
                    srow++;
                }
                const int rowsValid = min(ROWS, m0t + tmv - m0);
                const int colsValid = min(4, lastF - f0 + 1);
                if (colsValid <= 0)
                    continue;
                Strip<T, R, ROWS> su, sd;
                su.load(gs, sCur, m0 - mTop, f0 - fLeft, rowsValid);
                if (VARDEN)
                    sd.load(gs, sRho, m0 - mTop, f0 - fLeft, rowsValid);
                T out[ROWS][4];
#pragma unroll
                for (int i = 0; i < ROWS; i++) {
                    if (i >= rowsValid)
                        continue;
                    const int ps = (m0 + i - mTop) * PS + (f0 - fLeft);
                    T pv[4], cv[4], qv[4];
                    load4<T>(sPrev + ps, pv);
                    load4<T>(sC0 + ps, cv);
                    load4<T>(sQ + ps, qv);
                    if constexpr (Loop2dShape<T, R, VARDEN>::PACKED && MATH == MATH_FAST) {
                        strip_values_packed<R>(a, su, pv, cv, qv, out[i]);
                    } else {
#pragma unroll
                        for (int c = 0; c < 4; c++) {
                            const StripNeighbours<T, R, ROWS> nu{su, i, c};
                            const StripNeighbours<T, R, ROWS> nd{VARDEN ? sd : su, i, c};
                            out[i][c] = value_from_neighbours<T, 2, VARDEN, R, MATH>(
                                a, nu, nd, pv[c], cv[c], qv[c]);
                        }
                    }
                }
                const bool clearF = !a.fuse_bc || (f0 > 2 * R && f0 + 3 < lastF - R);
#pragma unroll
                for (int i = 0; i < ROWS; i++) {
                    if (i >= rowsValid)
                        continue;
                    const int m = m0 + i;
                    const bool inBox = L.fuseSources && m >= L.srcLoM && m <= L.srcHiM &&
                                       f0 + 3 >= L.srcLoF && f0 <= L.srcHiF;
                    if (inBox) {
#pragma unroll
                        for (int c = 0; c < 4; c++)
                            if (c < colsValid)
                                out[i][c] = loop2d_add_sources<T>(L, n, m, f0 + c, out[i][c]);
                    }
                    const bool clearM = !a.fuse_bc || (m > 2 * R && m < lastM - R);
                    T own[4];
                    if (clearF && clearM && colsValid == 4) {
                        store4(nextG + g.at(0, m, f0), out[i]);
#pragma unroll
                        for (int c = 0; c < 4; c++)
                            own[c] = out[i][c];
                    } else {
#pragma unroll
                        for (int c = 0; c < 4; c++) {
                            own[c] = T(0);
                            if (c < colsValid) {
                                simple_store<T, 2>(a, nextG, 0, m, f0 + c, out[i][c]);
                                own[c] = own_cell_value<T>(a, m, f0 + c, out[i][c]);
                            }
                        }
                    }
                    // in place over u_prev: nobody else reads this cell's old value
                    store4(sPrev + (m - mTop) * PS + (f0 - fLeft), own);
                }
            }
        } else {
            // one point per thread, operands one by one from the shared tiles
            const int total = tmv * tfv;
            for (int idx = tid; idx < total; idx += blockDim.x) {
                const int m = m0t + idx / tfv, f = f0t + idx % tfv;
                const int ps = (m - mTop) * PS + (f - fLeft);
                const TileNeighbours<T> nu{sCur + ps, PS};
                const TileNeighbours<T> nd{VARDEN ? sRho + ps : sCur + ps, PS};
                T v = value_from_neighbours<T, 2, VARDEN, R, MATH>(a, nu, nd, sPrev[ps], sC0[ps],
                                                                   sQ[ps]);
                if (L.fuseSources && m >= L.srcLoM && m <= L.srcHiM && f >= L.srcLoF &&
                    f <= L.srcHiF)
                    v = loop2d_add_sources<T>(L, n, m, f, v);
                simple_store<T, 2>(a, nextG, 0, m, f, v);
                sPrev[ps] = own_cell_value<T>(a, m, f, v);
            }
        }

        mark();
        // publish step n, wait for the neighbours' step n, fetch the new ring
        __syncthreads();
        const unsigned done = (unsigned)(n - L.begin + 1);
        if (tid == 0)
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(L.flags + tile), "r"(done)
                         : "memory");
        if (n < L.end) {
            if (nbTile >= 0) {
                unsigned seen;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];"
                                 : "=r"(seen)
                                 : "l"(L.flags + nbTile)
                                 : "memory");
                } while ((int)(seen - done) < 0);
            }
            __syncthreads();
            fetch_ring(sPrev, nextG);
            __syncthreads();
        }
        T *t = sCur; sCur = sPrev; sPrev = t;
        mark();
    }
}

// ---- host side ---------------------------------------------------------------
template <typename T, bool VARDEN, int R>
static bool resident_tiling_r(int math, const Grid &g, int minTm, int minTf, Loop2dTiling *out)
{
    constexpr int RP = (R + 3) / 4 * 4;
    constexpr int ROWS = (Loop2dShape<T, R, VARDEN>::VW == 4 && Loop2dShape<T, R, VARDEN>::ROWS)
                             ? Loop2dShape<T, R, VARDEN>::ROWS : 1;
    auto k = (math == MATH_STRICT) ? loop2d_resident_kernel<T, VARDEN, R, MATH_STRICT>
                                   : loop2d_resident_kernel<T, VARDEN, R, MATH_FAST>;
    int dev = 0, sms = 0, coop = 0, maxSmem = 0;
    SW_CUDA(cudaGetDevice(&dev));
    SW_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop)
        return false;
    SW_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    SW_CUDA(cudaDeviceGetAttribute(&maxSmem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const int nMi = g.nM - 2 * R, nFi = g.nF - 2 * R;
    const int fields = VARDEN ? 5 : 4;
    minTm = std::max(minTm, R);
    minTf = std::max(minTf, RP);
    // one CTA per tile, one CTA per SM: the fewest points per CTA wins, the
    // halo ring a CTA fetches per step breaks ties
    double best = 1e300;
    Loop2dTiling bt{};
    for (int tf = 32; tf <= 1024; tf += 32) {
        if (tf < minTf)
            continue;
        const int tilesF = (nFi + tf - 1) / tf;
        if (tilesF > sms)
            continue;
        const int tilesMmax = sms / tilesF;
        for (int tilesM = 1; tilesM <= tilesMmax; tilesM++) {
            int tm = (nMi + tilesM - 1) / tilesM;
            tm = (tm + ROWS - 1) / ROWS * ROWS;
            if (tm < minTm)
                break;
            const int realM = (nMi + tm - 1) / tm;
            const size_t smem = (size_t)fields * (tm + 2 * R) * (tf + 2 * RP) * sizeof(T);
            if (smem > (size_t)maxSmem)
                continue;
            const double cost = (double)tm * tf + 0.5 * (2.0 * R * tf + 2.0 * RP * tm);
            if (cost < best) {
                best = cost;
                bt = {tm, tf, realM, tilesF, smem};
            }
        }
    }
    if (best == 1e300)
        return false;
    SW_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bt.smemBytes));
    int perSm = 0;
    SW_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
        &perSm, k, Loop2dShape<T, R, VARDEN>::RESIDENT_THREADS, bt.smemBytes));
    if ((long long)perSm * sms < (long long)bt.tilesM * bt.tilesF)
        return false;
    *out = bt;
    return true;
}

template <typename T, bool VARDEN, int R>
static bool launch_resident_r(int math, const LoopArgs<T> &L, const Loop2dTiling &tl,
                              cudaStream_t stream)
{
    auto k = (math == MATH_STRICT) ? loop2d_resident_kernel<T, VARDEN, R, MATH_STRICT>
                                   : loop2d_resident_kernel<T, VARDEN, R, MATH_FAST>;
    void *args[] = {(void *)&L};
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)k, dim3((unsigned)(tl.tilesM * tl.tilesF)),
                                                dim3(Loop2dShape<T, R, VARDEN>::RESIDENT_THREADS), args,
                                                tl.smemBytes, stream);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported) {
        cudaGetLastError();
        return false;
    }
    SW_CUDA(e);
    return true;
}

template <typename T, bool VARDEN>
bool loop2d_resident_tiling(int math, const Grid &g, int minTm, int minTf, Loop2dTiling *out)
{
    switch (g.r) {
#define SW_CASE(R) case R: return resident_tiling_r<T, VARDEN, R>(math, g, minTm, minTf, out);
        SW_CASE(1) SW_CASE(2) SW_CASE(3) SW_CASE(4) SW_CASE(5)
        SW_CASE(6) SW_CASE(7) SW_CASE(8) SW_CASE(9) SW_CASE(10)
#undef SW_CASE
    default:
        return false;
    }
}

template <typename T, bool VARDEN>
bool launch_loop2d_resident(int math, const LoopArgs<T> &L, const Loop2dTiling &tl,
                            cudaStream_t stream)
{
    switch (L.a.g.r) {
#define SW_CASE(R) case R: return launch_resident_r<T, VARDEN, R>(math, L, tl, stream);
        SW_CASE(1) SW_CASE(2) SW_CASE(3) SW_CASE(4) SW_CASE(5)
        SW_CASE(6) SW_CASE(7) SW_CASE(8) SW_CASE(9) SW_CASE(10)
#undef SW_CASE
    default:
        return false;
    }
}

}  // namespace sw
