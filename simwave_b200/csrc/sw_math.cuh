// Arithmetic of one grid-point update, written once for every kernel.
//
// The reference kernel is C99 compiled without contraction; parts of its
// update are promoted to double by double literals (constant_density/3d/
// wave.c:177-185).  Two math modes are offered:
//
//   STRICT  every operation is rounded exactly where the reference rounds it
//           (no FMA contraction, true divisions, the same float/double mix).
//           With identical inputs the result is bit-identical to the
//           reference's sequential C kernel.
//   FAST    what a GPU compiler does with the same source by default (nvcc
//           contracts to FMA; the reference's own benchmark flags add fast
//           math): the stencil coefficients are pre-divided by h^2 and
//           accumulated with FMA, and the leapfrog combination is evaluated in
//           the working precision.  Agrees with the reference within the
//           float32 noise floor the reference shows between its own builds
//           (rel-L2 ~1e-6); this is the default.
#pragma once

#include "sw_common.h"

namespace sw {

enum MathMode { MATH_FAST = 0, MATH_STRICT = 1 };


template <typename T>
struct Ops;

template <>
struct Ops<float> {
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return fmaf(a, b, c); }
};

template <>
struct Ops<double> {
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return ::fma(a, b, c); }
};

// acc + c * (a + b)     -- one ring of a second-derivative stencil
template <typename T, int MATH>
__device__ __forceinline__ T ring_sum(T acc, T c, T a, T b)
{
    if (MATH == MATH_STRICT)
        return Ops<T>::add(acc, Ops<T>::mul(c, Ops<T>::add(a, b)));
    return Ops<T>::fma(c, Ops<T>::add(a, b), acc);
}

// acc + c * (a - b)     -- one ring of a first-derivative stencil
template <typename T, int MATH>
__device__ __forceinline__ T ring_diff(T acc, T c, T a, T b)
{
    if (MATH == MATH_STRICT)
        return Ops<T>::add(acc, Ops<T>::mul(c, Ops<T>::sub(a, b)));
    return Ops<T>::fma(c, Ops<T>::sub(a, b), acc);
}

// x / h2
template <typename T, int MATH>
__device__ __forceinline__ T over_h2(T x, T h2, T inv_h2)
{
    if (MATH == MATH_STRICT)
        return Ops<T>::div(x, h2);
    return Ops<T>::mul(x, inv_h2);
}

// sum of the per-axis second derivatives, F first, S last
// (3d/wave.c:174: sum_y/dy2 + sum_x/dx2 + sum_z/dz2; 2d/wave.c:167)
template <typename T, int NDIM, int MATH>
__device__ __forceinline__ T laplacian(T sdS, T sdM, T sdF, const T *h2, const T *inv_h2)
{
    T v = Ops<T>::add(over_h2<T, MATH>(sdF, h2[AX_F], inv_h2[AX_F]),
                      over_h2<T, MATH>(sdM, h2[AX_M], inv_h2[AX_M]));
    if (NDIM == 3)
        v = Ops<T>::add(v, over_h2<T, MATH>(sdS, h2[AX_S], inv_h2[AX_S]));
    // the reference starts from `value = 0.0; value += ...`
    return Ops<T>::add(T(0), v);
}

// variable-density correction: value -= (tF + tM + tS) / rho
// with t = (fd_p * fd_rho) / (4 * h2)   (variable_density/3d/wave.c:196-200)
template <typename T, int NDIM>
__device__ __forceinline__ T density_term(T value, T fpS, T frS, T fpM, T frM, T fpF, T frF,
                                          const T *four_h2, T rho)
{
    T tF = Ops<T>::div(Ops<T>::mul(fpF, frF), four_h2[AX_F]);
    T tM = Ops<T>::div(Ops<T>::mul(fpM, frM), four_h2[AX_M]);
    T t = Ops<T>::add(tF, tM);
    if (NDIM == 3) {
        T tS = Ops<T>::div(Ops<T>::mul(fpS, frS), four_h2[AX_S]);
        t = Ops<T>::add(t, tS);
    }
    return Ops<T>::sub(value, Ops<T>::div(t, rho));
}

// FAST-mode form of the same correction: the density factor of every axis is
// folded once per point into g = fr / (4 h^2 rho) (it is constant in time),
// leaving value - (fpF gF + fpM gM + fpS gS).
template <typename T>
__device__ __forceinline__ T density_weight(T fr, T inv_four_h2, T rho)
{
    return Ops<T>::mul(Ops<T>::mul(fr, inv_four_h2), Ops<T>::div(T(1), rho));
}
template <typename T, int NDIM>
__device__ __forceinline__ T fast_density_term(T value, T fpS, T gS, T fpM, T gM, T fpF, T gF)
{
    T t = Ops<T>::mul(fpF, gF);
    t = Ops<T>::fma(fpM, gM, t);
    if (NDIM == 3)
        t = Ops<T>::fma(fpS, gS, t);
    return Ops<T>::sub(value, t);
}

// Damping factors of a point inside an absorbing layer (q != 0):
//   D = 1.0 + q, N = 1.0 - q, the literal 1.0 making the add a double add
//   (3d/wave.c:180-181).
template <typename T>
__device__ __forceinline__ void damping_factors(T q, T &D, T &N)
{
    D = (T)__dadd_rn(1.0, (double)q);
    N = (T)__dsub_rn(1.0, (double)q);
}

// u_next = 2.0/D * u - (N/D) * u_prev + value * (c0/D)        (3d/wave.c:183-185)
// `lap` is the spatial operator before the dt^2 v^2 scaling.
template <typename T>
__device__ __forceinline__ T leapfrog(T lap, T u, T prev, T c0, T q)
{
    if (q == T(0)) {
        // D == N == 1: 2.0/D*u and (N/D)*prev are exact, value = lap*c0
        T value = Ops<T>::mul(lap, c0);
        double d = __dsub_rn((double)Ops<T>::mul(T(2), u), (double)prev);
        return (T)__dadd_rn(d, (double)value);
    }
    T D, N;
    damping_factors(q, D, N);
    T value = Ops<T>::mul(lap, Ops<T>::div(c0, D));
    double a = __dmul_rn(__ddiv_rn(2.0, (double)D), (double)u);
    T b = Ops<T>::mul(Ops<T>::div(N, D), prev);
    return (T)__dadd_rn(__dsub_rn(a, (double)b), (double)value);
}

// FAST-mode update.  `lap` already carries the 1/h^2 factors.
//   q == 0:  u_next = (2u - prev) + lap*c0
//   q != 0:  u_next = (2u - (1-q) prev + lap*c0) * (1/(1+q))
// (the same expression as above multiplied out: 2/D u - N/D prev + lap c0/D).
// 2u - prev is exact in the working precision wherever u and prev are of the
// same sign and size (the usual case: the field varies slowly from one step
// to the next), so rounding it separately costs nothing measurable against
// the reference's double accumulation (measured: tools/parity_report.py).
// 1/d for d in [1, 2): float32 goes through the hardware reciprocal
// (rcp.approx, 1 ulp) and one Newton step instead of an IEEE division -- the
// division's range checks and slow path cost more than the rest of the update
// inside an absorbing layer.  float64 divides.
template <typename T>
__device__ __forceinline__ T fast_reciprocal(T d)
{
    return Ops<T>::div(T(1), d);
}
template <>
__device__ __forceinline__ float fast_reciprocal<float>(float d)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    return Ops<float>::mul(r, Ops<float>::fma(-d, r, 2.0f));
}

template <typename T>
__device__ __forceinline__ T fast_leapfrog(T lap, T u, T prev, T c0, T q)
{
    if (q == T(0))
        return Ops<T>::fma(lap, c0, Ops<T>::fma(T(2), u, -prev));
    const T N = Ops<T>::sub(T(1), q);
    const T t = Ops<T>::fma(lap, c0, Ops<T>::fma(-N, prev, Ops<T>::add(u, u)));
    return Ops<T>::mul(t, fast_reciprocal<T>(Ops<T>::add(T(1), q)));
}

// One interface for both modes: `lap` is what Stencil<..>::laplacian returns.
template <typename T, int MATH>
__device__ __forceinline__ T update_point(T lap, T u, T prev, T c0, T q)
{
    if (MATH == MATH_STRICT)
        return leapfrog<T>(lap, u, prev, c0, q);
    return fast_leapfrog<T>(lap, u, prev, c0, q);
}

// Accumulator of the second-derivative stencil of one point.
//   STRICT: one sum per axis, every operation rounded as in the reference,
//           divided by h^2 at the end (laplacian<>).
//   FAST:   the same sums as FMA chains, scaled by the rounded 1/h^2.
template <typename T, int NDIM, int MATH>
struct Stencil {
    T sF, sM, sS;
    __device__ __forceinline__ void begin(const StepArgs<T> &a, T u)
    {
        // both modes start every axis from c[0]*u, as the reference does: the
        // half stencil sums to zero only with the coefficients it was built
        // from, and that cancellation must survive (a constant field has no
        // Laplacian)
        sF = sM = sS = Ops<T>::mul(a.c2[0], u);
    }
    // ring `ir`: neighbours at +-ir along F, M and S
    __device__ __forceinline__ void ring(const StepArgs<T> &a, int ir, T fp, T fm, T mp, T mm,
                                         T sp, T sm)
    {
        sF = ring_sum<T, MATH>(sF, a.c2[ir], fp, fm);
        sM = ring_sum<T, MATH>(sM, a.c2[ir], mp, mm);
        if (NDIM == 3)
            sS = ring_sum<T, MATH>(sS, a.c2[ir], sp, sm);
    }
    // the same ring without the F axis (see split_f_sums)
    __device__ __forceinline__ void ring_ms(const StepArgs<T> &a, int ir, T mp, T mm, T sp, T sm)
    {
        sM = ring_sum<T, MATH>(sM, a.c2[ir], mp, mm);
        if (NDIM == 3)
            sS = ring_sum<T, MATH>(sS, a.c2[ir], sp, sm);
    }
    __device__ __forceinline__ T laplacian(const StepArgs<T> &a) const
    {
        if (MATH == MATH_STRICT)
            return sw::laplacian<T, NDIM, MATH>(sS, sM, sF, a.h2, a.inv_h2);
        // sum over axes of s/h^2 with 1/h^2 = inv_h2 + inv_h2_lo carried to
        // twice the working precision: a one-sided relative error in the
        // scale of the Laplacian acts like a velocity error and shows up as
        // a phase drift that grows with the number of time steps
        T lo = Ops<T>::mul(sF, a.inv_h2_lo[AX_F]);
        lo = Ops<T>::fma(sM, a.inv_h2_lo[AX_M], lo);
        if (NDIM == 3)
            lo = Ops<T>::fma(sS, a.inv_h2_lo[AX_S], lo);
        T t = Ops<T>::fma(sF, a.inv_h2[AX_F], lo);
        t = Ops<T>::fma(sM, a.inv_h2[AX_M], t);
        if (NDIM == 3)
            t = Ops<T>::fma(sS, a.inv_h2[AX_S], t);
        return t;
    }
};
template <typename T, int MATH>
using Stencil3 = Stencil<T, 3, MATH>;

// ---------------------------------------------------------------------------
// "Split" order of the F-axis sums (FAST mode, 3D, float32).
//
// The tiled kernel works on aligned pairs of F-neighbours (two-wide float32
// instructions).  A ring c[ir]*(u[+ir] + u[-ir]) with odd ir needs the pair
// (u[f+ir], u[f+1+ir]), which straddles two aligned pairs and costs register
// moves (about 60 per thread and plane at radius 8: a fifth of all issue
// slots).  Instead every value of the F window gets its own FMA: the even
// offsets accumulate in one chain with the pair in place, the odd offsets in
// a second chain whose two lanes are exchanged (lane 0 gathers the sum of the
// ODD point, lane 1 of the even one) with a coefficient PAIR per window pair
// (StepArgs::c2odd / c1odd); one lane-swapping add joins the chains.  Same
// number of arithmetic instructions, no moves.  Per point this reads:
//
//   E = c[0]*u;  for even offsets o != 0 ascending:  E = fma(c[|o|], u[o], E)
//   O = 0;       for odd offsets o ascending:        O = fma(c[|o|], u[o], O)
//   sum = E + O
//
// (first derivative: the same with sign(o)*c1[|o|], both chains from 0),
// independent of the lane a point sits in, so the plain kernel evaluates
// exactly this and the two kernels stay bit-identical.
// ---------------------------------------------------------------------------
template <typename T, int NDIM, int MATH>
constexpr bool kSplitF = false;
template <>
constexpr bool kSplitF<float, 3, MATH_FAST> = true;

template <int R, bool VARDEN, class U>
__device__ __forceinline__ void split_f_sums(const StepArgs<float> &a, const U &u, float &sF,
                                             float &fpF)
{
    float E = sF, O = 0.0f, dE = 0.0f, dO = 0.0f;
#pragma unroll
    for (int o = -R; o <= R; o++) {
        if (o == 0)
            continue;
        const int k = o < 0 ? -o : o;
        const float w = u.F(o);
        const float c1 = o < 0 ? -a.c1[k] : a.c1[k];
        if (k % 2 == 0) {
            E = fmaf(a.c2[k], w, E);
            if (VARDEN)
                dE = fmaf(c1, w, dE);
        } else {
            O = fmaf(a.c2[k], w, O);
            if (VARDEN)
                dO = fmaf(c1, w, dO);
        }
    }
    sF = __fadd_rn(E, O);
    fpF = __fadd_rn(dE, dO);
}

// ---------------------------------------------------------------------------
// Packed pairs.  sm_100a has two-wide float32 instructions (FADD2 / FMUL2 /
// FFMA2 on an aligned register pair): each lane is an ordinary IEEE
// round-to-nearest operation, so a pair of points goes through exactly the
// scalar helpers above, lane by lane, at half the issue slots.  Operations
// without a packed form (divisions, the double-promoted leapfrog of STRICT
// mode) are done per lane with the scalar helpers.
// ---------------------------------------------------------------------------
struct Pair {
    static __device__ __forceinline__ float2 bc(float s) { return make_float2(s, s); }
    static __device__ __forceinline__ float2 mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
    // ptxas (12.9) contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 although
    // both carry an explicit rounding; the scalar forms are left alone.
    // STRICT mode therefore multiplies lane by lane.
    static __device__ __forceinline__ float2 mul_exact(float2 a, float2 b)
    {
        return make_float2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
    }
    static __device__ __forceinline__ float2 add(float2 a, float2 b) { return __fadd2_rn(a, b); }
    // a - b: b * -1 is exact, so this is the one rounding of a scalar subtract
    static __device__ __forceinline__ float2 sub(float2 a, float2 b)
    {
        return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a);
    }
    static __device__ __forceinline__ float2 fma(float2 a, float2 b, float2 c)
    {
        return __ffma2_rn(a, b, c);
    }
};

template <int MATH>
__device__ __forceinline__ float2 ring_sum2(float2 acc, float c, float2 a, float2 b)
{
    if (MATH == MATH_STRICT)
        return Pair::add(acc, Pair::mul_exact(Pair::bc(c), Pair::add(a, b)));
    return Pair::fma(Pair::bc(c), Pair::add(a, b), acc);
}

template <int MATH>
__device__ __forceinline__ float2 ring_diff2(float2 acc, float c, float2 a, float2 b)
{
    if (MATH == MATH_STRICT)
        return Pair::add(acc, Pair::mul_exact(Pair::bc(c), Pair::sub(a, b)));
    return Pair::fma(Pair::bc(c), Pair::sub(a, b), acc);
}

// Stencil<float, 3, MATH> on two points at once
template <int MATH>
struct Stencil3x2 {
    float2 sF, sM, sS;
    __device__ __forceinline__ void begin(const StepArgs<float> &a, float2 u)
    {
        sF = sM = sS = (MATH == MATH_STRICT) ? Pair::mul_exact(Pair::bc(a.c2[0]), u)
                                             : Pair::mul(Pair::bc(a.c2[0]), u);
    }
    __device__ __forceinline__ void ringF(const StepArgs<float> &a, int ir, float2 p, float2 m)
    {
        sF = ring_sum2<MATH>(sF, a.c2[ir], p, m);
    }
    __device__ __forceinline__ void ringM(const StepArgs<float> &a, int ir, float2 p, float2 m)
    {
        sM = ring_sum2<MATH>(sM, a.c2[ir], p, m);
    }
    __device__ __forceinline__ void ringS(const StepArgs<float> &a, int ir, float2 p, float2 m)
    {
        sS = ring_sum2<MATH>(sS, a.c2[ir], p, m);
    }
    __device__ __forceinline__ float2 laplacian(const StepArgs<float> &a) const
    {
        if (MATH == MATH_STRICT)
            return make_float2(
                sw::laplacian<float, 3, MATH>(sS.x, sM.x, sF.x, a.h2, a.inv_h2),
                sw::laplacian<float, 3, MATH>(sS.y, sM.y, sF.y, a.h2, a.inv_h2));
        float2 lo = Pair::mul(sF, Pair::bc(a.inv_h2_lo[AX_F]));
        lo = Pair::fma(sM, Pair::bc(a.inv_h2_lo[AX_M]), lo);
        lo = Pair::fma(sS, Pair::bc(a.inv_h2_lo[AX_S]), lo);
        float2 t = Pair::fma(sF, Pair::bc(a.inv_h2[AX_F]), lo);
        t = Pair::fma(sM, Pair::bc(a.inv_h2[AX_M]), t);
        t = Pair::fma(sS, Pair::bc(a.inv_h2[AX_S]), t);
        return t;
    }
};

// fast_density_term<float, 3> on two points
__device__ __forceinline__ float2 fast_density_term2(float2 value, float2 fpS, float2 gS,
                                                     float2 fpM, float2 gM, float2 fpF, float2 gF)
{
    float2 t = Pair::mul(fpF, gF);
    t = Pair::fma(fpM, gM, t);
    t = Pair::fma(fpS, gS, t);
    return Pair::sub(value, t);
}

// update_point<float, MATH> on two points.  DAMPED == false: the caller knows
// q == 0 for both points (no absorbing layer inside this tile and plane).
template <int MATH, bool DAMPED>
__device__ __forceinline__ float2 update_pair(float2 lap, float2 u, float2 prev, float2 c0,
                                              float2 q)
{
    if (MATH == MATH_STRICT)
        return make_float2(update_point<float, MATH>(lap.x, u.x, prev.x, c0.x, DAMPED ? q.x : 0.0f),
                           update_point<float, MATH>(lap.y, u.y, prev.y, c0.y, DAMPED ? q.y : 0.0f));
    if (!DAMPED) {
        // fma(lap, c0, fma(2, u, -prev)), both lanes
        const float2 t = Pair::fma(Pair::bc(2.0f), u, make_float2(-prev.x, -prev.y));
        return Pair::fma(lap, c0, t);
    }
    // fast_leapfrog's damped form, lane for lane (for q == 0 it reduces to the
    // undamped expression bit for bit: N == 1, 1/(1+q) == 1)
    const float2 N = Pair::sub(Pair::bc(1.0f), q);
    const float2 t = Pair::fma(lap, c0, Pair::fma(make_float2(-N.x, -N.y), prev, Pair::add(u, u)));
    const float2 D = Pair::add(Pair::bc(1.0f), q);
    float2 r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(D.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(D.y));
    r = Pair::mul(r, Pair::fma(make_float2(-D.x, -D.y), r, Pair::bc(2.0f)));
    return Pair::mul(t, r);
}

// source increment: dt^2/slowness * kws * wavelet / D           (3d/wave.c:277)
template <typename T>
__device__ __forceinline__ T source_term(T c0, T q, T kws, T w)
{
    T D = T(1);
    if (q != T(0)) {
        T N;
        damping_factors(q, D, N);
    }
    return Ops<T>::div(Ops<T>::mul(Ops<T>::mul(c0, kws), w), D);
}

// per-point model coefficients from velocity and damping
// slowness = 1.0/(v*v) (double divide), c0 = dt^2/slowness, q = damp*dt/(2*slowness)
template <typename T>
__device__ __forceinline__ void model_coefficients(T v, T damp, T dt, T dtsq, T &c0, T &q)
{
    T slowness = (T)__ddiv_rn(1.0, (double)Ops<T>::mul(v, v));
    c0 = Ops<T>::div(dtsq, slowness);
    q = Ops<T>::div(Ops<T>::mul(damp, dt), Ops<T>::mul(T(2), slowness));
}

}  // namespace sw
