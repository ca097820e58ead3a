// Persistent time-loop kernel for 2D problems.
//
// A 2D model of the sizes simwave's benchmarks use (Marmousi: 0.8 M points,
// 3 MB per field) lives in L2, and one time step moves a few microseconds of
// data: launched as three kernels per step (receivers, stencil, sources) the
// loop is bound by launch latency, not by bandwidth.  This kernel is launched
// ONCE for a range of time steps with one CTA set resident on every SM
// (cooperative launch) and separates the phases of a step with grid-wide
// barriers instead of kernel boundaries:
//
//   step n:  [receivers sample u_cur]  [stencil + fused boundaries -> u_next]
//            ---- grid barrier ----
//            [source injection into u_next]
//            ---- grid barrier ----
//
// When every source window lies among the interior points (the usual case) the
// thread that owns a cell adds its source terms before the boundary-aware
// store -- stencil, then sources in index order, then boundary conditions, the
// reference's own order -- and the first barrier goes away.
//
// Same device functions, same arithmetic and the same order of the phases as
// the three-kernel path (constant_density/2d/wave.c:113-464), so the results
// are bit-identical to it in either math mode.  Only the three rotating slots
// of saving_stride == 0 are handled here; snapshot runs keep the per-step
// launches, which is also the fallback when a cooperative launch is refused.
#pragma once

#include "sw_launch.h"
#include "sw_points.cuh"
#include "sw_step_simple.cuh"

namespace sw {

constexpr int kLoop2dThreads = 384;
constexpr int kLoop2dCols = 128;                              // columns of a tile
constexpr int kLoop2dGroups = kLoop2dThreads / (kLoop2dCols / 4);   // row groups of a tile

// rows a thread updates per trip (0: one point per thread, scalar loads),
// bounded by the registers the neighbourhood of a strip takes: (ROWS + 2R) x 4
// elements of the column block plus ROWS x 2 x RP at its sides, twice that
// with a density field
template <typename T, int R, bool VARDEN>
struct Loop2dShape {
    static constexpr bool F32 = sizeof(T) == 4;
    // a strip is one 128-bit vector wide: 4 floats or 2 doubles
    static constexpr int VW = 16 / (int)sizeof(T);
    static constexpr int ROWS =
        F32 ? (!VARDEN ? (R <= 5 ? 1 : 0) : (R <= 2 ? 1 : 0)) : (!VARDEN ? (R <= 4 ? 1 : 0) : (R <= 2 ? 1 : 0));
    static constexpr int TILE_ROWS = kLoop2dGroups * (ROWS ? ROWS : 1);
    // columns of a tile: a warp covers one row of strips (the one-point-per-
    // thread path keeps 128 columns)
    static constexpr int COLS = ROWS ? 32 * VW : kLoop2dCols;
    // two-wide float32 arithmetic for the four points of a strip (FAST mode)
    static constexpr bool PACKED = F32 && !VARDEN && ROWS == 1;
    // threads of a CTA of the tile-resident loop (one CTA per SM): as many warps
    // as the registers of the strip allow
    static constexpr int RESIDENT_THREADS = PACKED ? 640 : 384;
};

// four consecutive elements through 128-bit loads / stores (16-byte aligned
// for float, 32-byte for double: f == r mod 4 in the pitched layout)
template <typename T>
__device__ __forceinline__ void load4(const T *p, T out[4]);
template <>
__device__ __forceinline__ void load4<float>(const float *p, float out[4])
{
    const float4 v = *reinterpret_cast<const float4 *>(p);
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
template <>
__device__ __forceinline__ void load4<double>(const double *p, double out[4])
{
    const double2 a = *reinterpret_cast<const double2 *>(p);
    const double2 b = *reinterpret_cast<const double2 *>(p + 2);
    out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y;
}
__device__ __forceinline__ void store4(float *p, const float v[4])
{
    *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(double *p, const double v[4])
{
    *reinterpret_cast<double2 *>(p) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2 *>(p + 2) = make_double2(v[2], v[3]);
}

// one 128-bit vector of elements: 4 floats or 2 doubles (16-byte aligned:
// f == r mod VW in the pitched layout)
__device__ __forceinline__ void loadv(const float *p, float out[4]) { load4<float>(p, out); }
__device__ __forceinline__ void loadv(const double *p, double out[2])
{
    const double2 a = *reinterpret_cast<const double2 *>(p);
    out[0] = a.x; out[1] = a.y;
}
__device__ __forceinline__ void storev(float *p, const float v[4]) { store4(p, v); }
__device__ __forceinline__ void storev(double *p, const double v[2])
{
    *reinterpret_cast<double2 *>(p) = make_double2(v[0], v[1]);
}

// Neighbourhood of a strip of ROWS x VW points (rows m0.., columns f0..f0+VW-1;
// VW = one 128-bit vector) held in registers: `col` = my columns over rows
// m0-R .. m0+ROWS+R-1 (vertically adjacent points share them), `side` = the RP
// elements left and right of my columns on my own rows.
template <typename T, int R, int ROWS>
struct Strip {
    static constexpr int VW = 16 / (int)sizeof(T);
    static constexpr int RP = (R + VW - 1) / VW * VW;
    T col[ROWS + 2 * R][VW];
    T side[ROWS][2][RP];

    __device__ __forceinline__ void load(const Grid &g, const T *field, int m0, int f0,
                                         int rowsValid)
    {
#pragma unroll
        for (int j = 0; j < ROWS + 2 * R; j++) {
            if (j < rowsValid + 2 * R) {
                loadv(field + g.at(0, m0 - R + j, f0), col[j]);
            } else {
#pragma unroll
                for (int e = 0; e < VW; e++)
                    col[j][e] = T(0);
            }
        }
#pragma unroll
        for (int i = 0; i < ROWS; i++)
#pragma unroll
            for (int b = 0; b < RP / VW; b++) {
                if (i < rowsValid) {
                    loadv(field + g.at(0, m0 + i, f0 - RP + VW * b), &side[i][0][VW * b]);
                    loadv(field + g.at(0, m0 + i, f0 + VW + VW * b), &side[i][1][VW * b]);
                } else {
#pragma unroll
                    for (int e = 0; e < VW; e++)
                        side[i][0][VW * b + e] = side[i][1][VW * b + e] = T(0);
                }
            }
    }
};

// value_from_neighbours' view of point (row i, column c) of a strip
template <typename T, int R, int ROWS>
struct StripNeighbours {
    const Strip<T, R, ROWS> &t;
    int i, c;
    static constexpr int RP = Strip<T, R, ROWS>::RP;
    static constexpr int VW = Strip<T, R, ROWS>::VW;
    __device__ __forceinline__ T C() const { return t.col[i + R][c]; }
    __device__ __forceinline__ T F(int k) const
    {
        const int idx = c + k;
        if (idx < 0)
            return t.side[i][0][RP + idx];
        if (idx > VW - 1)
            return t.side[i][1][idx - VW];
        return t.col[i + R][idx];
    }
    __device__ __forceinline__ T M(int k) const { return t.col[i + R + k][c]; }
    __device__ __forceinline__ T M1(int k) const { return M(k); }
    __device__ __forceinline__ T S(int) const { return T(0); }
};

// The four points of a one-row strip on the two-wide float32 instructions
// (FAST mode, constant density): every lane goes through the operations of
// value_from_neighbours in the same order -- Stencil<float, 2, FAST>::begin /
// ring / laplacian, then update_point -- so the result is bit-identical to the
// scalar path at half the arithmetic instructions.
template <int R>
__device__ __forceinline__ void strip_values_packed(const StepArgs<float> &a,
                                                    const Strip<float, R, 1> &t,
                                                    const float pv[4], const float cv[4],
                                                    const float qv[4], float out[4])
{
    constexpr int RP = Strip<float, R, 1>::RP;
    // the F window of the strip: left side | my four columns | right side
    float w[4 + 2 * RP];
#pragma unroll
    for (int k = 0; k < RP; k++) {
        w[k] = t.side[0][0][k];
        w[RP + 4 + k] = t.side[0][1][k];
    }
#pragma unroll
    for (int c = 0; c < 4; c++)
        w[RP + c] = t.col[R][c];
    const bool damped = (qv[0] != 0.0f) | (qv[1] != 0.0f) | (qv[2] != 0.0f) | (qv[3] != 0.0f);
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int c = 2 * h;
        const float2 u = make_float2(w[RP + c], w[RP + c + 1]);
        float2 sF = Pair::mul(Pair::bc(a.c2[0]), u), sM = sF;
#pragma unroll
        for (int ir = 1; ir <= R; ir++) {
            sF = ring_sum2<MATH_FAST>(sF, a.c2[ir],
                                      make_float2(w[RP + c + ir], w[RP + c + 1 + ir]),
                                      make_float2(w[RP + c - ir], w[RP + c + 1 - ir]));
            sM = ring_sum2<MATH_FAST>(sM, a.c2[ir],
                                      make_float2(t.col[R + ir][c], t.col[R + ir][c + 1]),
                                      make_float2(t.col[R - ir][c], t.col[R - ir][c + 1]));
        }
        // Stencil<float, 2, FAST>::laplacian
        float2 lo = Pair::mul(sF, Pair::bc(a.inv_h2_lo[AX_F]));
        lo = Pair::fma(sM, Pair::bc(a.inv_h2_lo[AX_M]), lo);
        float2 lap = Pair::fma(sF, Pair::bc(a.inv_h2[AX_F]), lo);
        lap = Pair::fma(sM, Pair::bc(a.inv_h2[AX_M]), lap);
        const float2 prev = make_float2(pv[c], pv[c + 1]);
        const float2 c0 = make_float2(cv[c], cv[c + 1]);
        const float2 q = make_float2(qv[c], qv[c + 1]);
        const float2 o = damped ? update_pair<MATH_FAST, true>(lap, u, prev, c0, q)
                                : update_pair<MATH_FAST, false>(lap, u, prev, c0, q);
        out[c] = o.x;
        out[c + 1] = o.y;
    }
}

// Grid-wide barrier for co-resident CTAs (cooperative launch): one
// release-increment per CTA on a monotonic counter, thread 0 spins with
// acquire loads until every CTA of this round has arrived.  `target` is the
// counter value that ends the round.
__device__ __forceinline__ void loop2d_barrier(unsigned *counter, unsigned target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        unsigned seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter)
                         : "memory");
        } while ((int)(seen - target) < 0);
    }
    __syncthreads();
}

// sources of one cell, in index order, added to the stencil value `v`
// (section 2 of the loop, 2d/wave.c:199-271; every window is interior)
template <typename T>
__device__ __forceinline__ T loop2d_add_sources(const LoopArgs<T> &L, long long n, int m, int f,
                                                T v)
{
    const StepArgs<T> &a = L.a;
    const long long p = a.g.at(0, m, f);
    for (int sid = 0; sid < L.src.count; sid++) {
        const Window<T, 2> win(L.src, sid);
        if (!win.contains(0, m, f))
            continue;
        const long long wo = L.waveletCount > 1 ? (n - 1) * L.src.count + sid : n - 1;
        const T wt = L.wavelet[wo];
        if (wt == T(0))
            continue;
        const T kws = win.weight(0, m - win.lo[1], f - win.lo[2]);
        v = Ops<T>::add(v, source_term<T>(a.c0[p], a.q[p], kws, wt));
    }
    return v;
}

template <typename T, bool VARDEN, int R, int MATH>
__global__ void __launch_bounds__(kLoop2dThreads, 2)
loop2d_persistent_kernel(const __grid_constant__ LoopArgs<T> L)
{
    constexpr int ROWS = Loop2dShape<T, R, VARDEN>::ROWS;
    constexpr int TILE_ROWS = Loop2dShape<T, R, VARDEN>::TILE_ROWS;
    constexpr int VW = Loop2dShape<T, R, VARDEN>::VW;
    constexpr int COLS = Loop2dShape<T, R, VARDEN>::COLS;
    const StepArgs<T> &a = L.a;
    const Grid &g = a.g;

    const int nFi = g.nF - 2 * R, nMi = g.nM - 2 * R;   // interior extents
    const int lastF = g.nF - R - 1, lastM = g.nM - R - 1;
    const int tilesF = (nFi + COLS - 1) / COLS;
    const int tilesM = (nMi + TILE_ROWS - 1) / TILE_ROWS;
    const long long tiles = (long long)tilesF * tilesM;
    const int warpsPerBlock = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned round = 0;
    auto barrier = [&]() { loop2d_barrier(L.barrier, ++round * gridDim.x); };
    // development aid (SIMWAVE_CUDA_LOOP2D_TRACE): CTA 0 stamps the phases of
    // the first steps with the nanosecond timer
    int stamp = 0;
    auto mark = [&]() {
        if (L.trace && blockIdx.x == 0 && threadIdx.x == 0 && stamp < 256) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            L.trace[stamp++] = t;
        }
    };
    // receivers are dealt to the warps from the LAST CTA downwards: with more
    // CTAs than tiles those have no stencil work, otherwise the least
    const long long nwarps = (long long)gridDim.x * warpsPerBlock;
    const long long gwarp = (long long)(gridDim.x - 1 - blockIdx.x) * warpsPerBlock + warp;

    // window of this warp's first receiver: origin, extents, and the lane's
    // weights along M (row = lane) and along F (column = lane)
    int recLoM = 0, recLoF = 0, recNM = 0, recNF = 0;
    T recWM = T(0), recWF = T(0);
    if (gwarp < L.rec.count) {
        const Window<T, 2> win(L.rec, (int)gwarp);
        recLoM = win.lo[1]; recLoF = win.lo[2];
        recNM = win.n[1]; recNF = win.n[2];
        if (lane < recNM) recWM = win.w[1][lane];
        if (lane < recNF) recWF = win.w[2][lane];
    }
    const bool recCached = gwarp < L.rec.count && recNM <= 32;

    for (long long n = L.begin; n <= L.end; n++) {
        // slot rotation of saving_stride == 0 (2d/wave.c:113-117)
        const T *prev = L.slot[(n - 1) % 3];
        const T *cur = L.slot[n % 3];
        T *next = L.slot[(n + 1) % 3];

        mark();
        // receivers: one warp per receiver, trace row n-1.  The warp's first
        // receiver has its window and weights in registers (set up before the
        // loop): per step only the wavefield samples are loaded
        if (L.rec.count) {
            T *row = L.recOut + (n - 1) * L.rec.count;
            if (recCached && MATH != MATH_STRICT) {
                // receiver_sample_fast with the weights already in registers
                T acc = T(0);
                for (int im = 0; im < recNM; im++) {
                    const T wm = __shfl_sync(0xffffffffu, recWM, im & 31);
                    if (lane < recNF)
                        acc = Ops<T>::fma(cur[g.at(0, recLoM + im, recLoF + lane)], wm, acc);
                }
                acc = (lane < recNF) ? Ops<T>::mul(Ops<T>::fma(acc, T(1), T(0)), recWF) : T(0);
                for (int d = 16; d > 0; d >>= 1)
                    acc = Ops<T>::add(acc, __shfl_xor_sync(0xffffffffu, acc, d));
                if (lane == 0)
                    row[gwarp] = acc;
            } else if (recCached) {
                T sum = T(0);
                constexpr int BATCH = 4;
                for (int r0 = 0; r0 < recNM; r0 += BATCH) {
                    T prod[BATCH];
#pragma unroll
                    for (int b = 0; b < BATCH; b++) {
                        const int im = r0 + b;
                        // weight = w_m * w_f, as Window::weight (2d/wave.c:243)
                        const T kws = Ops<T>::mul(__shfl_sync(0xffffffffu, recWM, im & 31), recWF);
                        prod[b] = T(0);
                        if (im < recNM && lane < recNF)
                            prod[b] = Ops<T>::mul(cur[g.at(0, recLoM + im, recLoF + lane)], kws);
                    }
#pragma unroll
                    for (int b = 0; b < BATCH; b++)
                        if (r0 + b < recNM)
                            for (int l = 0; l < recNF; l++)
                                sum = Ops<T>::add(sum, __shfl_sync(0xffffffffu, prod[b], l));
                }
                if (lane == 0)
                    row[gwarp] = sum;
            }
            for (long long rec = recCached ? gwarp + nwarps : gwarp; rec < L.rec.count;
                 rec += nwarps) {
                const T sum = (MATH == MATH_STRICT)
                                  ? receiver_sample<T, 2>(g, cur, L.rec, (int)rec, lane)
                                  : receiver_sample_fast<T, 2>(g, cur, L.rec, (int)rec, lane);
                if (lane == 0)
                    row[rec] = sum;
            }
        }

        mark();
        // stencil: tiles of TILE_ROWS x kLoop2dCols points, block-strided
        for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
            if constexpr (ROWS > 0) {
                // a thread owns ROWS x VW points; operands through 128-bit loads
                const int f0 = R + (int)(t % tilesF) * COLS + VW * (threadIdx.x & 31);
                const int m0 = R + (int)(t / tilesF) * TILE_ROWS + (threadIdx.x >> 5) * ROWS;
                const int rowsValid = min(ROWS, lastM - m0 + 1);
                const int colsValid = min(VW, lastF - f0 + 1);
                if (rowsValid <= 0 || colsValid <= 0)
                    continue;
                Strip<T, R, ROWS> su, sd;
                su.load(g, cur, m0, f0, rowsValid);
                if (VARDEN)
                    sd.load(g, a.rho, m0, f0, rowsValid);
                T out[ROWS][VW];
#pragma unroll
                for (int i = 0; i < ROWS; i++) {
                    if (i >= rowsValid)
                        continue;
                    const long long p = g.at(0, m0 + i, f0);
                    T pv[VW], cv[VW], qv[VW];
                    loadv(prev + p, pv);
                    loadv(a.c0 + p, cv);
                    loadv(a.q + p, qv);
                    if constexpr (Loop2dShape<T, R, VARDEN>::PACKED && MATH == MATH_FAST) {
                        strip_values_packed<R>(a, su, pv, cv, qv, out[i]);
                    } else {
#pragma unroll
                        for (int c = 0; c < VW; c++) {
                            const StripNeighbours<T, R, ROWS> nu{su, i, c};
                            const StripNeighbours<T, R, ROWS> nd{VARDEN ? sd : su, i, c};
                            out[i][c] = value_from_neighbours<T, 2, VARDEN, R, MATH>(
                                a, nu, nd, pv[c], cv[c], qv[c]);
                        }
                    }
                }
                // rows / columns clear of every face region: plain vector stores
                const bool clearF = !a.fuse_bc || (f0 > 2 * R && f0 + VW - 1 < lastF - R);
#pragma unroll
                for (int i = 0; i < ROWS; i++) {
                    if (i >= rowsValid)
                        continue;
                    const int m = m0 + i;
                    const bool inBox = L.fuseSources && m >= L.srcLoM && m <= L.srcHiM &&
                                       f0 + VW - 1 >= L.srcLoF && f0 <= L.srcHiF;
                    if (inBox) {
#pragma unroll
                        for (int c = 0; c < VW; c++)
                            if (c < colsValid)
                                out[i][c] = loop2d_add_sources<T>(L, n, m, f0 + c, out[i][c]);
                    }
                    const bool clearM = !a.fuse_bc || (m > 2 * R && m < lastM - R);
                    if (clearF && clearM && colsValid == VW) {
                        storev(next + g.at(0, m, f0), out[i]);
                    } else {
#pragma unroll
                        for (int c = 0; c < VW; c++)
                            if (c < colsValid)
                                simple_store<T, 2>(a, next, 0, m, f0 + c, out[i][c]);
                    }
                }
            } else {
                // one point per thread and row group, operands one by one
                const int f = R + (int)(t % tilesF) * kLoop2dCols + (threadIdx.x % kLoop2dCols);
                for (int m = R + (int)(t / tilesF) * TILE_ROWS + threadIdx.x / kLoop2dCols;
                     m < min(R + (int)(t / tilesF + 1) * TILE_ROWS, g.nM - R);
                     m += kLoop2dThreads / kLoop2dCols) {   // 3 rows per pass
                    if (f > lastF)
                        continue;
                    T v = simple_value<T, 2, VARDEN, R, MATH>(a, prev, cur, 0, m, f);
                    if (L.fuseSources && m >= L.srcLoM && m <= L.srcHiM && f >= L.srcLoF &&
                        f <= L.srcHiF)
                        v = loop2d_add_sources<T>(L, n, m, f, v);
                    simple_store<T, 2>(a, next, 0, m, f, v);
                }
            }
        }

        mark();
        if (L.src.count && !L.fuseSources) {
            barrier();
            // sources: blocks striped over (cells of a window) x (sources)
            const int gx = (L.srcMaxPoints + blockDim.x - 1) / blockDim.x;
            for (long long w = blockIdx.x; w < (long long)gx * L.src.count; w += gridDim.x)
                source_apply<T, 2>(a, next, L.src, L.wavelet, L.waveletCount, n, L.srcMode,
                                   (int)(w % gx), gx, (int)(w / gx), L.src.count);
        }
        if (!a.fuse_bc) {
            // literal boundary passes, F axis then M axis (2d/wave.c:285-393)
            for (int axis = AX_F; axis >= AX_M; axis--) {
                barrier();
                const int before = a.bc[2 * axis], after = a.bc[2 * axis + 1];
                const long long lines = (axis == AX_F ? g.nM : g.nF) - 2 * R;
                if (before | after)
                    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
                         i < lines; i += (long long)gridDim.x * blockDim.x)
                        boundary_line<T>(g, next, axis, before, after, i);
            }
        }
        barrier();
        mark();
    }
}

template <typename T, bool VARDEN, int R>
static bool launch_loop2d_r(int math, const LoopArgs<T> &L, cudaStream_t stream)
{
    auto kStrict = loop2d_persistent_kernel<T, VARDEN, R, MATH_STRICT>;
    auto kFast = loop2d_persistent_kernel<T, VARDEN, R, MATH_FAST>;
    auto k = (math == MATH_STRICT) ? kStrict : kFast;
    int dev = 0, sms = 0, coop = 0, perSm = 0;
    SW_CUDA(cudaGetDevice(&dev));
    SW_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop)
        return false;
    SW_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    SW_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k, kLoop2dThreads, 0));
    if (perSm < 1)
        return false;
    // no more CTAs than tiles of work; a barrier costs more the wider it is
    const Grid &g = L.a.g;
    constexpr int TILE_ROWS = Loop2dShape<T, R, VARDEN>::TILE_ROWS;
    constexpr int COLS = Loop2dShape<T, R, VARDEN>::COLS;
    const long long tiles = (long long)((g.nM - 2 * R + TILE_ROWS - 1) / TILE_ROWS) *
                            ((g.nF - 2 * R + COLS - 1) / COLS);
    long long blocks = (long long)sms * std::min(perSm, 2);
    blocks = std::max<long long>(1, std::min(blocks, tiles));
    void *args[] = {(void *)&L};
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)k, dim3((unsigned)blocks),
                                                dim3(kLoop2dThreads),
                                                args, 0, stream);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported) {
        cudaGetLastError();
        return false;
    }
    SW_CUDA(e);
    return true;
}

template <typename T, bool VARDEN>
bool launch_loop2d(int math, const LoopArgs<T> &L, cudaStream_t stream)
{
    switch (L.a.g.r) {
#define SW_CASE(R) case R: return launch_loop2d_r<T, VARDEN, R>(math, L, stream);
        SW_CASE(1) SW_CASE(2) SW_CASE(3) SW_CASE(4) SW_CASE(5)
        SW_CASE(6) SW_CASE(7) SW_CASE(8) SW_CASE(9) SW_CASE(10)
#undef SW_CASE
    default:
        return false;
    }
}

}  // namespace sw
