// Small kernels around the stencil: model preparation, layout conversion,
// source injection, receiver sampling and the stand-alone boundary passes.
#pragma once

#include "sw_math.cuh"

namespace sw {

// ---------------------------------------------------------------------------
// layout conversion / model preparation (run once per forward call)
// ---------------------------------------------------------------------------

// dense C-order (nS,nM,nF) -> pitched
template <typename T>
__global__ void pack_kernel(Grid g, const T *__restrict__ dense, T *__restrict__ pitched)
{
    const long long rows = (long long)g.nS * g.nM;
    for (long long row = blockIdx.y; row < rows; row += gridDim.y)
        for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < g.nF; f += gridDim.x * blockDim.x)
            pitched[row * g.pitch + f] = dense[row * g.nF + f];
}

// velocity, damping (dense) -> c0 = dt^2/slowness, q = damp*dt/(2*slowness) (pitched)
// with the reference's roundings (constant_density/3d/wave.c:177-183)
template <typename T>
__global__ void model_kernel(Grid g, const T *__restrict__ velocity, const T *__restrict__ damp,
                             T dt, T dtsq, T *__restrict__ c0, T *__restrict__ q)
{
    const long long rows = (long long)g.nS * g.nM;
    for (long long row = blockIdx.y; row < rows; row += gridDim.y)
        for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < g.nF; f += gridDim.x * blockDim.x) {
            T a, b;
            model_coefficients<T>(velocity[row * g.nF + f], damp[row * g.nF + f], dt, dtsq, a, b);
            c0[row * g.pitch + f] = a;
            q[row * g.pitch + f] = b;
        }
}

// First derivatives of the density along every axis at the interior points,
// accumulated ring by ring exactly as the step kernels do
// (variable_density/3d/wave.c:180-194).  They are constant in time; the tiled
// kernel streams them instead of keeping a halo of the density.  In fast math
// mode the time-invariant factor 1/(4 h^2 rho) is folded in as well
// (density_weight), so the density itself need not be streamed.
template <typename T, int MATH>
__global__ void rho_gradient_kernel(const __grid_constant__ StepArgs<T> a, T *__restrict__ frF,
                                    T *__restrict__ frM, T *__restrict__ frS)
{
    const Grid &g = a.g;
    const int r = g.r;
    const int f = r + blockIdx.x * blockDim.x + threadIdx.x;
    const int m = r + blockIdx.y;
    const int s = r + blockIdx.z;
    if (f >= g.nF - r)
        return;
    const long long p = g.at(s, m, f);
    const T *d = a.rho + p;
    T gF = T(0), gM = T(0), gS = T(0);
    for (int ir = 1; ir <= r; ir++) {
        const long long oM = (long long)ir * g.pitch, oS = (long long)ir * g.planeStride;
        gF = ring_diff<T, MATH>(gF, a.c1[ir], d[ir], d[-ir]);
        gM = ring_diff<T, MATH>(gM, a.c1[ir], d[oM], d[-oM]);
        gS = ring_diff<T, MATH>(gS, a.c1[ir], d[oS], d[-oS]);
    }
    if (MATH != MATH_STRICT) {
        const T rho = d[0];
        gF = density_weight<T>(gF, a.inv_four_h2[AX_F], rho);
        gM = density_weight<T>(gM, a.inv_four_h2[AX_M], rho);
        gS = density_weight<T>(gS, a.inv_four_h2[AX_S], rho);
    }
    frF[p] = gF;
    frM[p] = gM;
    frS[p] = gS;
}

// flags[(s*tilesM + tm)*tilesF + tf] = 1 where q != 0 somewhere among the
// interior points of tile (tm,tf) on plane s.  One block per (tile, plane);
// blockIdx.z counts interior planes.
template <typename T>
__global__ void qflag_kernel(Grid g, const T *__restrict__ q, int tileM, int tileF,
                             unsigned char *__restrict__ flags)
{
    const int s = g.r + blockIdx.z;
    const int m0 = g.r + blockIdx.y * tileM, f0 = g.r + blockIdx.x * tileF;
    const int m1 = min(m0 + tileM, g.nM - g.r), f1 = min(f0 + tileF, g.nF - g.r);
    int any = 0;
    for (int idx = threadIdx.x; idx < tileM * tileF; idx += blockDim.x) {
        const int m = m0 + idx / tileF, f = f0 + idx % tileF;
        if (m < m1 && f < f1 && q[g.at(s, m, f)] != T(0))
            any = 1;
    }
    if (__syncthreads_or(any) && threadIdx.x == 0)
        flags[((long long)s * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = 1;
}

// ---------------------------------------------------------------------------
// sources
// ---------------------------------------------------------------------------

// window of one source/receiver; axis order S,M,F (tables: z|x|y resp. z|x)
template <typename T, int NDIM>
struct Window {
    int lo[3], n[3];
    const T *w[3];
    __device__ Window(const PointTables<T> &t, int id)
    {
        const unsigned long long *iv = t.intervals + (size_t)id * 2 * NDIM;
        const T *v = t.values + t.offsets[id];
        lo[0] = 0; n[0] = 1; w[0] = nullptr;
        for (int a = 0; a < NDIM; a++) {
            const int ax = a + (3 - NDIM);
            lo[ax] = (int)iv[2 * a];
            n[ax] = (int)(iv[2 * a + 1] - iv[2 * a]) + 1;
            w[ax] = v;
            v += n[ax];
        }
    }
    __device__ int points() const { return n[0] * n[1] * n[2]; }
    // (vz*vx)*vy in 3D, vz*vx in 2D (3d/wave.c:265, 2d/wave.c:243)
    __device__ T weight(int is, int im, int jf) const
    {
        if (NDIM == 3)
            return Ops<T>::mul(Ops<T>::mul(w[0][is], w[1][im]), w[2][jf]);
        return Ops<T>::mul(w[1][im], w[2][jf]);
    }
    __device__ bool contains(int s, int m, int f) const
    {
        return s >= lo[0] && s < lo[0] + n[0] && m >= lo[1] && m < lo[1] + n[1] &&
               f >= lo[2] && f < lo[2] + n[2];
    }
};

// What the boundary passes would do to an increment `t` added at (s,m,f):
// the linear image of store_with_boundaries for a delta.  `fused` selects
// between that and a plain add (separate boundary kernels follow).
template <typename T, int NDIM>
__device__ __forceinline__ void add_with_boundaries(const StepArgs<T> &a, T *next, int s, int m,
                                                    int f, T t, bool atomic)
{
    const Grid &g = a.g;
    const int r = g.r;
    const long long p = g.at(s, m, f);
    // slab decomposition: a cell of my outermost owned planes is mirrored into
    // the neighbour's ghost plane (atomic mode: the ghost copy is refreshed by
    // the engine's plane copy instead, see Plan::slab_push)
    T *alt = (NDIM == 3 && !atomic) ? a.ghost_copy(s) : nullptr;
    // `samePlane`: the target lies on plane s (everything but S mirrors, which
    // exist only at outer faces and have no ghost copy)
    auto add = [&](long long idx, bool samePlane = true) {
        if (atomic) {
            atomicAdd(next + idx, t);
        } else {
            const T sum = Ops<T>::add(next[idx], t);
            next[idx] = sum;
            if (alt && samePlane)
                alt[idx] = sum;
        }
    };
    if (!a.fuse_bc) {
        add(p);
        return;
    }
    const int firstF = r, lastF = g.nF - r - 1;
    const int firstM = r, lastM = g.nM - r - 1;
    const int firstS = r, lastS = g.nS - r - 1;
    const bool inF = f >= firstF && f <= lastF;
    const bool inM = m >= firstM && m <= lastM;
    const bool inS = (NDIM == 2) || (s >= firstS && s <= lastS);

    if (!(inF && inM && inS)) {
        // halo cell: the reference adds, then a Neumann pass overwrites the
        // cell if it is a mirror target (interior on the other axes, in the
        // halo of a Neumann face); otherwise the sum stays.
        bool target = false;
        if (inM && inS && !inF)
            target = (f < firstF) ? a.bc[4] == 2 : a.bc[5] == 2;
        if (inF && inS && !inM)
            target = (m < firstM) ? a.bc[2] == 2 : a.bc[3] == 2;
        if (NDIM == 3 && inF && inM && !inS)
            target = (s < firstS) ? a.bc[0] == 2 : a.bc[1] == 2;
        if (!target)
            add(p, inS);
        return;
    }

    const bool zFb = (a.bc[4] == 1) & (f == firstF);
    const bool zFa = (a.bc[5] == 1) & (f == lastF);
    const bool zMb = (a.bc[2] == 1) & (m == firstM);
    const bool zMa = (a.bc[3] == 1) & (m == lastM);
    bool zSb = false, zSa = false;
    if (NDIM == 3) {
        zSb = (a.bc[0] == 1) & (s == firstS);
        zSa = (a.bc[1] == 1) & (s == lastS);
    }
    if (!(zFb | zFa | zMb | zMa | zSb | zSa))
        add(p);
    if (a.bc[4] == 2 && f > firstF && f <= firstF + r)
        add(p - 2 * (f - firstF));
    if (a.bc[5] == 2 && f < lastF && f >= lastF - r && !zFb)
        add(p + 2 * (lastF - f));
    const bool zF = zFb | zFa;
    if (a.bc[2] == 2 && m > firstM && m <= firstM + r && !zF)
        add(p - 2 * (long long)(m - firstM) * g.pitch);
    if (a.bc[3] == 2 && m < lastM && m >= lastM - r && !zF && !zMb)
        add(p + 2 * (long long)(lastM - m) * g.pitch);
    if (NDIM == 3) {
        const bool zFM = zF | zMb | zMa;
        if (a.bc[0] == 2 && s > firstS && s <= firstS + r && !zFM)
            add(p - 2 * (long long)(s - firstS) * g.planeStride, false);
        if (a.bc[1] == 2 && s < lastS && s >= lastS - r && !zFM && !zSb)
            add(p + 2 * (long long)(lastS - s) * g.planeStride, false);
    }
}

enum SourceMode {
    SRC_DISJOINT = 0,  // no two windows share a cell: one thread per (source, cell)
    SRC_ORDERED = 1,   // windows overlap: the lowest source covering a cell adds
                       // every covering source in index order (sequential-C order)
    SRC_ATOMIC = 2     // very many overlapping sources: atomics (the order the
                       // reference's OpenMP build has: unspecified)
};

// Section 2 of the reference loop (3d/wave.c:208-295, 2d/wave.c:199-271).
// Work is spread like a 2D launch: (bx, gx) = block index / count along the
// cells of a window, (by, gy) along the sources.
template <typename T, int NDIM>
__device__ __forceinline__ void source_apply(const StepArgs<T> &a, T *next,
                                             const PointTables<T> &tab,
                                             const T *__restrict__ wavelet, int waveletCount,
                                             long long step, int mode, int bx, int gx, int by,
                                             int gy)
{
    auto wavelet_of = [&](int sid) {
        long long wo = step - 1;
        if (waveletCount > 1)
            wo = (step - 1) * tab.count + sid;
        return wavelet[wo];
    };

    for (int src = by; src < tab.count; src += gy) {
        const Window<T, NDIM> win(tab, src);
        const int total = win.points();
        const T w = wavelet_of(src);
        if (mode != SRC_ORDERED && w == T(0))
            continue;

        for (int idx = bx * blockDim.x + threadIdx.x; idx < total; idx += gx * blockDim.x) {
            const int jf = idx % win.n[2];
            const int im = (idx / win.n[2]) % win.n[1];
            const int is = idx / (win.n[2] * win.n[1]);
            const int s = win.lo[0] + is, m = win.lo[1] + im, f = win.lo[2] + jf;
            const long long p = a.g.at(s, m, f);

            if (mode == SRC_ORDERED) {
                // am I the first source covering this cell?
                bool owner = true;
                for (int t = 0; t < src && owner; t++)
                    owner = !Window<T, NDIM>(tab, t).contains(s, m, f);
                if (!owner)
                    continue;
                for (int t = src; t < tab.count; t++) {
                    const Window<T, NDIM> other(tab, t);
                    if (t != src && !other.contains(s, m, f))
                        continue;
                    const T wt = wavelet_of(t);
                    if (wt == T(0))
                        continue;
                    const T kws =
                        other.weight(s - other.lo[0], m - other.lo[1], f - other.lo[2]);
                    add_with_boundaries<T, NDIM>(a, next, s, m, f,
                                                 source_term<T>(a.c0[p], a.q[p], kws, wt), false);
                }
            } else {
                const T kws = win.weight(is, im, jf);
                add_with_boundaries<T, NDIM>(a, next, s, m, f,
                                             source_term<T>(a.c0[p], a.q[p], kws, w),
                                             mode == SRC_ATOMIC);
            }
        }
    }
}

template <typename T, int NDIM>
__global__ void source_kernel(const __grid_constant__ StepArgs<T> a, PointTables<T> tab,
                              const T *__restrict__ wavelet, int waveletCount, long long step,
                              int mode)
{
    source_apply<T, NDIM>(a, a.next, tab, wavelet, waveletCount, step, mode, blockIdx.x,
                          gridDim.x, blockIdx.y, gridDim.y);
}

// ---------------------------------------------------------------------------
// receivers: Section 4 of the reference loop (3d/wave.c:499-566,
// 2d/wave.c:412-464).  One warp per receiver; the lanes fetch a row of the
// window (F axis) together, then every lane folds the row into the running
// sum in the reference's order (S outer, F inner, one rounding per product and
// per add), so the trace is bit-identical to the sequential C loop.
// ---------------------------------------------------------------------------
template <typename T, int NDIM>
__device__ __forceinline__ T receiver_sample(const Grid &g, const T *cur,
                                             const PointTables<T> &tab, int rec, int lane)
{
    const Window<T, NDIM> win(tab, rec);
    const int rows = win.n[0] * win.n[1];
    const int nf = win.n[2];     // <= 2*10+1 < 32

    T sum = T(0);
    constexpr int BATCH = 4;
    for (int r0 = 0; r0 < rows; r0 += BATCH) {
        T prod[BATCH];
#pragma unroll
        for (int b = 0; b < BATCH; b++) {
            const int rr = r0 + b;
            prod[b] = T(0);
            if (rr < rows && lane < nf) {
                const int is = rr / win.n[1], im = rr % win.n[1];
                const T kws = win.weight(is, im, lane);
                prod[b] = Ops<T>::mul(cur[g.at(win.lo[0] + is, win.lo[1] + im, win.lo[2] + lane)],
                                      kws);
            }
        }
#pragma unroll
        for (int b = 0; b < BATCH; b++) {
            if (r0 + b < rows)
                for (int l = 0; l < nf; l++)
                    sum = Ops<T>::add(sum, __shfl_sync(0xffffffffu, prod[b], l));
        }
    }
    return sum;
}

// FAST mode: every lane sums its own column of the window (F offset = lane)
// over the rows with FMAs, then the lanes are added pairwise.  The serial
// chain above costs one shuffle and one dependent add per window point (729
// for a half-width of 4: ~35 us per launch); this one 81 FMAs and 5 shuffles.
// The sum is the same up to the rounding of a different association.
template <typename T, int NDIM>
__device__ __forceinline__ T receiver_sample_fast(const Grid &g, const T *cur,
                                                  const PointTables<T> &tab, int rec, int lane)
{
    const Window<T, NDIM> win(tab, rec);
    T acc = T(0);
    if (lane < win.n[2]) {
        const T wf = win.w[2][lane];
        for (int is = 0; is < win.n[0]; is++) {
            const T ws = (NDIM == 3) ? win.w[0][is] : T(1);
            T part = T(0);
            for (int im = 0; im < win.n[1]; im++)
                part = Ops<T>::fma(cur[g.at(win.lo[0] + is, win.lo[1] + im, win.lo[2] + lane)],
                                   win.w[1][im], part);
            acc = Ops<T>::fma(part, ws, acc);
        }
        acc = Ops<T>::mul(acc, wf);
    }
    for (int d = 16; d > 0; d >>= 1)
        acc = Ops<T>::add(acc, __shfl_xor_sync(0xffffffffu, acc, d));
    return acc;
}

template <typename T, int NDIM, int MATH>
__global__ void receiver_kernel(Grid g, const T *__restrict__ cur, PointTables<T> tab,
                                T *__restrict__ row)
{
    const int lane = threadIdx.x & 31;
    const int rec = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (rec >= tab.count)
        return;
    const T sum = (MATH == MATH_STRICT) ? receiver_sample<T, NDIM>(g, cur, tab, rec, lane)
                                        : receiver_sample_fast<T, NDIM>(g, cur, tab, rec, lane);
    if (lane == 0)
        row[rec] = sum;
}

// ---------------------------------------------------------------------------
// stand-alone boundary pass over one axis: the literal per-line sequence of
// the reference (3d/wave.c:324-366 etc.).  Used for grids too small for the
// fused form and as an independent check of it (SIMWAVE_CUDA_BC=separate).
// `axis` is AX_S / AX_M / AX_F; one thread per line.
// ---------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void boundary_line(const Grid &g, T *next, int axis, int before,
                                              int after, long long tid)
{
    const int r = g.r;
    int n[3] = {g.nS, g.nM, g.nF};
    long long stride[3] = {g.planeStride, g.pitch, 1};
    int oa = -1, ob = -1;   // the looped axes, ob the faster one
    for (int a = 0; a < 3; a++) {
        if (a == axis || (g.ndim == 2 && a == AX_S))
            continue;
        if (oa < 0) oa = a; else ob = a;
    }
    // 2D: only one other axis
    const int nb = (ob >= 0) ? n[ob] - 2 * r : 1;
    const int na = n[oa] - 2 * r;
    if (tid >= (long long)na * nb)
        return;
    const int ia = (int)(tid / nb) + r;
    const int ib = (ob >= 0) ? (int)(tid % nb) + r : 0;
    T *line = next + ia * stride[oa] + ((ob >= 0) ? ib * stride[ob] : 0);
    const long long sa = stride[axis];
    const int first = r, last = n[axis] - r - 1;

    if (before == 1)
        line[first * sa] = T(0);
    if (before == 2)
        for (int ir = 1; ir <= r; ir++)
            line[(first - ir) * sa] = line[(first + ir) * sa];
    if (after == 1)
        line[last * sa] = T(0);
    if (after == 2)
        for (int ir = 1; ir <= r; ir++)
            line[(last + ir) * sa] = line[(last - ir) * sa];
}

template <typename T>
__global__ void boundary_axis_kernel(Grid g, T *next, int axis, int before, int after)
{
    boundary_line<T>(g, next, axis, before, after,
                     (long long)blockIdx.x * blockDim.x + threadIdx.x);
}

}  // namespace sw
