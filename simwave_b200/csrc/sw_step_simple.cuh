// Reference-order update of one grid point and the plain (one thread per
// point, operands straight from global memory) step kernel built on it.
//
// The plain kernel is the correctness baseline on the device: it serves every
// variant (2D/3D, constant/variable density, float/double, any radius, the
// bug-compatible strides of variable_density/3d/wave.c:185-186) and is what
// the tiled kernels are validated against.
#pragma once

#include <type_traits>

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

// New value of one interior point from its neighbourhood (before sources and
// boundary conditions): the arithmetic of section 1 of the reference loop
// (constant_density/3d/wave.c:142-188, variable_density/3d/wave.c:145-214 and
// the 2D files), written once for every kernel.  `U` / `D` give the wavefield
// / density at signed offsets along an axis: C() centre, F(k), M(k), S(k), and
// M1(k) = the neighbour the FIRST derivative along M uses (it differs from M(k)
// only under the bug-compatible strides of variable_density/3d/wave.c:185-186).
// A density accessor may instead carry the three first derivatives of the
// density already summed (rho_gradient_kernel: same ring order and roundings;
// in FAST mode already weighted by 1 / (4 h^2 rho)): it then declares
// `static constexpr bool kDerivatives = true` and offers frF() / frM() / frS().
template <class D, class = void>
struct has_derivatives : std::false_type {};
template <class D>
struct has_derivatives<D, std::void_t<decltype(D::kDerivatives)>> : std::true_type {};

template <typename T, int NDIM, bool VARDEN, int R, int MATH, class U, class D>
__device__ __forceinline__ T value_from_neighbours(const StepArgs<T> &a, const U &u, const D &d,
                                                   T prevv, T c0v, T qv)
{
    constexpr bool kPre = VARDEN && has_derivatives<D>::value;
    const T uc = u.C();

    // second derivatives: c[0]*u + sum_ir c[ir]*(u[+ir] + u[-ir]), per axis
    Stencil<T, NDIM, MATH> acc;
    acc.begin(a, uc);
    T fpF = T(0), fpM = T(0), fpS = T(0);
    T frF = T(0), frM = T(0), frS = T(0);

    // F-axis sums of u in split order where the tiled kernel uses it
    constexpr bool SPLIT = kSplitF<T, NDIM, MATH>;
    if constexpr (SPLIT)
        split_f_sums<R, VARDEN>(a, u, acc.sF, fpF);

#pragma unroll
    for (int ir = 1; ir <= R; ir++) {
        if constexpr (SPLIT)
            acc.ring_ms(a, ir, u.M(ir), u.M(-ir), u.S(ir), u.S(-ir));
        else
            acc.ring(a, ir, u.F(ir), u.F(-ir), u.M(ir), u.M(-ir), NDIM == 3 ? u.S(ir) : T(0),
                     NDIM == 3 ? u.S(-ir) : T(0));
        if (VARDEN) {
            if constexpr (!SPLIT)
                fpF = ring_diff<T, MATH>(fpF, a.c1[ir], u.F(ir), u.F(-ir));
            fpM = ring_diff<T, MATH>(fpM, a.c1[ir], u.M1(ir), u.M1(-ir));
            if (NDIM == 3)
                fpS = ring_diff<T, MATH>(fpS, a.c1[ir], u.S(ir), u.S(-ir));
            if constexpr (!kPre) {
                frF = ring_diff<T, MATH>(frF, a.c1[ir], d.F(ir), d.F(-ir));
                frM = ring_diff<T, MATH>(frM, a.c1[ir], d.M1(ir), d.M1(-ir));
                if (NDIM == 3)
                    frS = ring_diff<T, MATH>(frS, a.c1[ir], d.S(ir), d.S(-ir));
            }
        }
    }

    T lap = acc.laplacian(a);
    if constexpr (kPre) {
        if (MATH == MATH_STRICT)
            lap = density_term<T, NDIM>(lap, fpS, d.frS(), fpM, d.frM(), fpF, d.frF(), a.four_h2,
                                        d.C());
        else
            lap = fast_density_term<T, NDIM>(lap, fpS, d.frS(), fpM, d.frM(), fpF, d.frF());
    } else if (VARDEN) {
        if (MATH == MATH_STRICT) {
            lap = density_term<T, NDIM>(lap, fpS, frS, fpM, frM, fpF, frF, a.four_h2, d.C());
        } else {
            const T rho = d.C();
            lap = fast_density_term<T, NDIM>(
                lap, fpS, NDIM == 3 ? density_weight<T>(frS, a.inv_four_h2[AX_S], rho) : T(0),
                fpM, density_weight<T>(frM, a.inv_four_h2[AX_M], rho), fpF,
                density_weight<T>(frF, a.inv_four_h2[AX_F], rho));
        }
    }

    return update_point<T, MATH>(lap, uc, prevv, c0v, qv);
}

// neighbourhood of (s,m,f) straight from a pitched field in global memory
template <typename T>
struct GlobalNeighbours {
    const T *base;      // element (0,0,0) of the field
    const T *p;         // element (s,m,f)
    const Grid *g;
    int s, m, f, quirk, denseNx, denseNy;
    __device__ __forceinline__ T C() const { return p[0]; }
    __device__ __forceinline__ T F(int k) const { return p[k]; }
    __device__ __forceinline__ T M(int k) const { return p[(long long)k * g->pitch]; }
    __device__ __forceinline__ T S(int k) const { return p[(long long)k * g->planeStride]; }
    __device__ __forceinline__ T M1(int k) const
    {
        if (quirk)   // the reference steps by k*nx dense elements here
            return base[dense_shift(*g, denseNx, denseNy, s, m, f, (long long)k * denseNx)];
        return M(k);
    }
};

// New value of interior point (s,m,f) from `cur` / `prev`, operands straight
// from global memory.
template <typename T, int NDIM, bool VARDEN, int R, int MATH>
__device__ __forceinline__ T simple_value(const StepArgs<T> &a, const T *prev, const T *cur,
                                          int s, int m, int f)
{
    const Grid &g = a.g;
    const long long p = g.at(s, m, f);
    const int quirk = (NDIM == 3 && VARDEN) ? a.quirk : 0;
    const GlobalNeighbours<T> u{cur, cur + p, &g, s, m, f, quirk, a.denseNx, a.denseNy};
    const GlobalNeighbours<T> d{a.rho, VARDEN ? a.rho + p : nullptr, &g, s, m, f, quirk,
                                a.denseNx, a.denseNy};
    return value_from_neighbours<T, NDIM, VARDEN, R, MATH>(a, u, d, prev[p], a.c0[p], a.q[p]);
}

// Stores the value of interior point (s,m,f) into `next`, with the boundary
// cells that depend on it when the conditions are fused.
template <typename T, int NDIM>
__device__ __forceinline__ void simple_store(const StepArgs<T> &a, T *next, int s, int m, int f,
                                             T val)
{
    const Grid &g = a.g;
    const long long p = g.at(s, m, f);
    if (a.fuse_bc) {
        store_with_boundaries<T, NDIM>(a, next, p, s, m, f, val);
    } else {
        next[p] = val;
        if (NDIM == 3) {
            if (T *alt = a.ghost_copy(s))
                alt[p] = val;
        }
    }
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
    simple_store<T, NDIM>(a, a.next, s, m, f,
                          simple_value<T, NDIM, VARDEN, R, MATH>(a, a.prev, a.cur, s, m, f));
}

}  // namespace sw
