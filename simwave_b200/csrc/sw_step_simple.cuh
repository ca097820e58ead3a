// Reference-order update of one grid point and the plain (one thread per
// point, operands straight from global memory) step kernel built on it.
//
// The plain kernel is the correctness baseline on the device: it serves every
// variant (2D/3D, constant/variable density, float/double, any radius, the
// bug-compatible strides of variable_density/3d/wave.c:185-186) and is what
// the tiled kernels are validated against.
#pragma once

#include "sw_math.cuh"

namespace sw {

// ---------------------------------------------------------------------------
// Boundary conditions fused into the store of an interior point.
//
// The reference applies them to the finished field in three passes, F axis
// first, then M, then S (3d/wave.c:311-480: y, x, z; 2d/wave.c:285-393: x, z),
// each pass visiting only interior indices of the other axes and doing, per
// line: before-Dirichlet, before-Neumann, after-Dirichlet, after-Neumann.
// Dirichlet (code 1) zeroes the first/last interior plane, Neumann (code 2)
// copies interior plane first+ir to halo plane first-ir (last-ir to last+ir).
// Every cell written by those passes is a function of exactly one interior
// value, so the thread that owns that value can write all of them, provided
// it reproduces what the earlier passes had already done to the value by the
// time the reference reads it: zeroing by Dirichlet planes of axes processed
// earlier (and of the `before` face of the same axis for an `after` mirror).
// Valid when every extent is >= 3r+2 (mirror sources are then interior
// cells); smaller grids take the separate boundary kernels.
// ---------------------------------------------------------------------------
template <typename T, int NDIM>
__device__ __forceinline__ void store_with_boundaries(const StepArgs<T> &a, T *next, long long p,
                                                      int s, int m, int f, T val)
{
    const Grid &g = a.g;
    const int r = g.r;
    const int firstF = r, lastF = g.nF - r - 1;
    const int firstM = r, lastM = g.nM - r - 1;
    const int firstS = r, lastS = g.nS - r - 1;
    // slab decomposition: cells of my outermost owned planes are written into
    // the neighbour's ghost planes as well (F/M mirror cells included; S
    // mirrors only exist at outer faces, which have no neighbour)
    T *alt = (NDIM == 3) ? a.ghost_copy(s) : nullptr;
    auto put = [&](long long idx, T v) {
        next[idx] = v;
        if (alt)
            alt[idx] = v;
    };

    const bool zFb = (a.bc[4] == 1) & (f == firstF);
    const bool zFa = (a.bc[5] == 1) & (f == lastF);
    const bool zMb = (a.bc[2] == 1) & (m == firstM);
    const bool zMa = (a.bc[3] == 1) & (m == lastM);
    bool zSb = false, zSa = false;
    if (NDIM == 3) {
        zSb = (a.bc[0] == 1) & (s == firstS);
        zSa = (a.bc[1] == 1) & (s == lastS);
    }

    put(p, (zFb | zFa | zMb | zMa | zSb | zSa) ? T(0) : val);

    // F pass
    if (a.bc[4] == 2 && f > firstF && f <= firstF + r)
        put(p - 2 * (f - firstF), val);
    if (a.bc[5] == 2 && f < lastF && f >= lastF - r)
        put(p + 2 * (lastF - f), zFb ? T(0) : val);

    // M pass sees the F pass' zeroing
    const T vM = (zFb | zFa) ? T(0) : val;
    if (a.bc[2] == 2 && m > firstM && m <= firstM + r)
        put(p - 2 * (long long)(m - firstM) * g.pitch, vM);
    if (a.bc[3] == 2 && m < lastM && m >= lastM - r)
        put(p + 2 * (long long)(lastM - m) * g.pitch, zMb ? T(0) : vM);

    // S pass sees the F and M passes' zeroing
    if (NDIM == 3) {
        const T vS = (zFb | zFa | zMb | zMa) ? T(0) : val;
        if (a.bc[0] == 2 && s > firstS && s <= firstS + r)
            next[p - 2 * (long long)(s - firstS) * g.planeStride] = vS;
        if (a.bc[1] == 2 && s < lastS && s >= lastS - r)
            next[p + 2 * (long long)(lastS - s) * g.planeStride] = zSb ? T(0) : vS;
    }
}

// Pitched index of the cell that lies `delta` elements away from (s,m,f) in
// the caller's dense C-order array.  Only used to stay bug-compatible with
// variable_density/3d/wave.c:185-186 when nx != ny.
__device__ __forceinline__ long long dense_shift(const Grid &g, int denseNx, int denseNy, int s,
                                                 int m, int f, long long delta)
{
    long long flat = ((long long)s * denseNx + m) * denseNy + f + delta;
    long long row = flat / denseNy;
    int ff = (int)(flat - row * denseNy);
    long long ss = row / denseNx;
    int mm = (int)(row - ss * denseNx);
    return g.at((int)ss, mm, ff);
}

template <typename T, int NDIM, bool VARDEN, int R, int MATH>
__global__ void __launch_bounds__(256)
step_simple_kernel(const __grid_constant__ StepArgs<T> a)
{
    const Grid &g = a.g;
    const int f = R + blockIdx.x * blockDim.x + threadIdx.x;
    const int m = R + blockIdx.y * blockDim.y + threadIdx.y;
    const int s = (NDIM == 3) ? R + blockIdx.z : 0;
    if (f >= g.nF - R || m >= g.nM - R)
        return;

    const long long p = g.at(s, m, f);
    const T *u = a.cur + p;
    const T uc = u[0];

    // second derivatives: c[0]*u + sum_ir c[ir]*(u[+ir] + u[-ir]), per axis
    Stencil<T, NDIM, MATH> acc;
    acc.begin(a, uc);
    T fpF = T(0), fpM = T(0), fpS = T(0);
    T frF = T(0), frM = T(0), frS = T(0);
    const T *d = VARDEN ? a.rho + p : nullptr;

#pragma unroll
    for (int ir = 1; ir <= R; ir++) {
        const long long oM = (long long)ir * g.pitch;
        const long long oS = (long long)ir * g.planeStride;
        acc.ring(a, ir, u[ir], u[-ir], u[oM], u[-oM], NDIM == 3 ? u[oS] : T(0),
                 NDIM == 3 ? u[-oS] : T(0));
        if (VARDEN) {
            fpF = ring_diff<T, MATH>(fpF, a.c1[ir], u[ir], u[-ir]);
            frF = ring_diff<T, MATH>(frF, a.c1[ir], d[ir], d[-ir]);
            if (NDIM == 3 && a.quirk) {
                // reference steps by ir*nx dense elements here
                const long long hi = dense_shift(g, a.denseNx, a.denseNy, s, m, f,
                                                 (long long)ir * a.denseNx);
                const long long lo = dense_shift(g, a.denseNx, a.denseNy, s, m, f,
                                                 -(long long)ir * a.denseNx);
                fpM = ring_diff<T, MATH>(fpM, a.c1[ir], a.cur[hi], a.cur[lo]);
                frM = ring_diff<T, MATH>(frM, a.c1[ir], a.rho[hi], a.rho[lo]);
            } else {
                fpM = ring_diff<T, MATH>(fpM, a.c1[ir], u[oM], u[-oM]);
                frM = ring_diff<T, MATH>(frM, a.c1[ir], d[oM], d[-oM]);
            }
            if (NDIM == 3) {
                fpS = ring_diff<T, MATH>(fpS, a.c1[ir], u[oS], u[-oS]);
                frS = ring_diff<T, MATH>(frS, a.c1[ir], d[oS], d[-oS]);
            }
        }
    }

    T lap = acc.laplacian(a);
    if (VARDEN) {
        if (MATH == MATH_STRICT) {
            lap = density_term<T, NDIM>(lap, fpS, frS, fpM, frM, fpF, frF, a.four_h2, d[0]);
        } else {
            const T rho = d[0];
            lap = fast_density_term<T, NDIM>(
                lap, fpS, NDIM == 3 ? density_weight<T>(frS, a.inv_four_h2[AX_S], rho) : T(0),
                fpM, density_weight<T>(frM, a.inv_four_h2[AX_M], rho), fpF,
                density_weight<T>(frF, a.inv_four_h2[AX_F], rho));
        }
    }

    const T val = update_point<T, MATH>(lap, uc, a.prev[p], a.c0[p], a.q[p]);

    if (a.fuse_bc) {
        store_with_boundaries<T, NDIM>(a, a.next, p, s, m, f, val);
    } else {
        a.next[p] = val;
        if (NDIM == 3) {
            if (T *alt = a.ghost_copy(s))
                alt[p] = val;
        }
    }
}

}  // namespace sw
