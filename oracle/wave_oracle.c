/*
 * wave_oracle.c -- CPU restatement of simwave's acoustic forward time loop.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under simwave_b200/ may import, link or
 * execute this file.  It exists so that the CUDA path can be checked on a box
 * where /root/reference is not available, and it is itself pinned bit-for-bit
 * against the reference's own wave.c (built into oracle/_ref/ by the Makefile)
 * by tests/test_oracle.py.
 *
 * One source, eight variants, selected at compile time:
 *     -DNDIM=2|3   -DVARDEN=0|1   -DFLOAT | -DDOUBLE   [-fopenmp -DORACLE_OMP]
 * Each variant exports a single symbol `forward` with the argument list of the
 * matching reference kernel:
 *     NDIM=2 VARDEN=0  simwave/kernel/backend/c_code/forward/constant_density/2d/wave.c:23-36
 *     NDIM=3 VARDEN=0  .../constant_density/3d/wave.c:23-36
 *     NDIM=2 VARDEN=1  .../variable_density/2d/wave.c:23-36
 *     NDIM=3 VARDEN=1  .../variable_density/3d/wave.c:23-36
 *
 * The code is organised around axis strides instead of the reference's
 * per-dimension copy-and-paste, but every floating-point expression keeps the
 * reference's operand types and association order (including the places where
 * a double literal promotes part of an expression), so with the same compiler
 * flags (-O3 -std=c99, i.e. no contraction, no fast-math) the output is
 * bit-identical.  The layout is C order, last axis contiguous: (z,x) or (z,x,y).
 */
#include <stddef.h>
#include <sys/time.h>

#if defined(FLOAT)
typedef float real;
#elif defined(DOUBLE)
typedef double real;
#else
#error "compile with -DFLOAT or -DDOUBLE"
#endif

#ifndef NDIM
#error "compile with -DNDIM=2 or -DNDIM=3"
#endif
#ifndef VARDEN
#define VARDEN 0
#endif

#if defined(ORACLE_OMP)
#define PARALLEL_FOR _Pragma("omp parallel for")
#else
#define PARALLEL_FOR
#endif

/* Grid description shared by all phases.  Axis 0 is z (slowest). */
typedef struct {
    size_t n[3];      /* extent per axis (n[2] == 1 in 2D)          */
    size_t s[3];      /* element stride per axis                     */
    size_t cells;     /* n[0]*n[1]*n[2]                              */
    size_t r;         /* stencil radius = space_order / 2            */
} grid_t;

static size_t cell(const grid_t *g, size_t i, size_t j, size_t k)
{
    return i * g->s[0] + j * g->s[1] + k * g->s[2];
}

/*
 * Phase 1: leapfrog update of every interior point.
 * constant density: 3d/wave.c:142-188, 2d/wave.c:140-180
 * variable density: 3d/wave.c:145-214, 2d/wave.c:140-202
 */
static void update_interior(const grid_t *g, const real *prev, const real *cur,
                            real *next, const real *vel, const real *rho,
                            const real *damp, const real *c2, const real *c1,
                            const real *hsq, real dt, real dtsq)
{
    const size_t r = g->r;
    const size_t k_lo = (NDIM == 3) ? r : 0;
    const size_t k_hi = (NDIM == 3) ? g->n[2] - r : 1;

#if VARDEN && NDIM == 3
    /* variable_density/3d/wave.c:185-186 steps the x first derivatives by
     * ir*nx elements, not ir*ny.  Kept as is: the reference is the oracle. */
    const size_t s_fd_x = g->n[1];
#elif VARDEN
    const size_t s_fd_x = g->s[NDIM - 1];
#endif

    PARALLEL_FOR
    for (size_t i = r; i < g->n[0] - r; i++) {
        for (size_t j = r; j < g->n[1] - r; j++) {
            for (size_t k = k_lo; k < k_hi; k++) {
                const size_t p = cell(g, i, j, k);
                const real *u = cur + p;

                real value = 0.0;

                /* second derivative, one accumulator per axis;
                 * axis index a counts from the contiguous axis backwards
                 * only in the final sum, see below */
                real sd[3];
                for (int a = 0; a < NDIM; a++)
                    sd[a] = c2[0] * u[0];
#if VARDEN
                real fp[3] = {0.0, 0.0, 0.0};
                real fr[3] = {0.0, 0.0, 0.0};
                const real *d = rho + p;
#endif
                for (size_t ir = 1; ir <= r; ir++) {
                    for (int a = 0; a < NDIM; a++) {
                        const size_t o = ir * g->s[a];
                        sd[a] += c2[ir] * (u[o] + u[-(ptrdiff_t)o]);
#if VARDEN
                        size_t of = o;
                        if (NDIM == 3 && a == 1)
                            of = ir * s_fd_x;
                        if (NDIM == 2 && a == 1)
                            of = ir * s_fd_x;
                        fp[a] += c1[ir] * (u[of] - u[-(ptrdiff_t)of]);
                        fr[a] += c1[ir] * (d[of] - d[-(ptrdiff_t)of]);
#endif
                    }
                }

                /* contiguous axis first, z last (3d/wave.c:174, 2d/wave.c:167) */
#if NDIM == 3
                value += sd[2] / hsq[2] + sd[1] / hsq[1] + sd[0] / hsq[0];
#else
                value += sd[1] / hsq[1] + sd[0] / hsq[0];
#endif

#if VARDEN
                {
#if NDIM == 3
                    real ty = (fp[2] * fr[2]) / (4 * hsq[2]);
                    real tx = (fp[1] * fr[1]) / (4 * hsq[1]);
                    real tz = (fp[0] * fr[0]) / (4 * hsq[0]);
                    value -= (ty + tx + tz) / d[0];
#else
                    real tx = (fp[1] * fr[1]) / (4 * hsq[1]);
                    real tz = (fp[0] * fr[0]) / (4 * hsq[0]);
                    value -= (tx + tz) / d[0];
#endif
                }
#endif
                /* double literal: the divide happens in double, the result is
                 * stored in `real` (3d/wave.c:177) */
                real slowness = 1.0 / (vel[p] * vel[p]);

                /* float quotient, double add, stored in `real` (:180-181) */
                real den = (1.0 + damp[p] * dt / (2 * slowness));
                real num = (1.0 - damp[p] * dt / (2 * slowness));

                value *= (dtsq / slowness) / den;

                /* 2.0/den*u in double, (num/den)*prev in `real`, sum in
                 * double, one rounding on store (:185) */
                next[p] = 2.0 / den * u[0] - (num / den) * prev[p] + value;
            }
        }
    }
}

/* Box of grid points and separable weights of one source / receiver.
 * Table layout: kernel/frontend/source.py:124-159 and kws.py:138-183. */
typedef struct {
    size_t lo[3], hi[3];      /* inclusive index range per axis          */
    const real *w[3];         /* weights per axis, w[a][0] at index lo[a] */
} window_t;

static window_t window_of(const grid_t *g, const size_t *intervals,
                          const real *values, const size_t *offsets, size_t id)
{
    window_t win;
    const size_t *iv = intervals + id * 2 * NDIM;
    const real *v = values + offsets[id];
    for (int a = 0; a < 3; a++) {
        win.lo[a] = 0;
        win.hi[a] = 0;
        win.w[a] = 0;
    }
    for (int a = 0; a < NDIM; a++) {
        win.lo[a] = iv[2 * a];
        win.hi[a] = iv[2 * a + 1];
        win.w[a] = v;
        v += win.hi[a] - win.lo[a] + 1;
    }
    (void)g;
    return win;
}

/*
 * Phase 2: add the source term to the new field.
 * 3d/wave.c:208-295, 2d/wave.c:199-271
 */
static void inject_sources(const grid_t *g, real *next, const real *vel,
                           const real *damp, const real *wavelet,
                           size_t wavelet_count, size_t n,
                           const size_t *intervals, const real *values,
                           const size_t *offsets, size_t num_sources,
                           real dt, real dtsq)
{
    for (size_t src = 0; src < num_sources; src++) {
        size_t wo = n - 1;
        if (wavelet_count > 1)
            wo = (n - 1) * num_sources + src;
        if (wavelet[wo] == 0.0)
            continue;

        window_t win = window_of(g, intervals, values, offsets, src);
        for (size_t i = win.lo[0]; i <= win.hi[0]; i++) {
            for (size_t j = win.lo[1]; j <= win.hi[1]; j++) {
#if NDIM == 3
                for (size_t k = win.lo[2]; k <= win.hi[2]; k++) {
                    real kws = win.w[0][i - win.lo[0]] * win.w[1][j - win.lo[1]]
                               * win.w[2][k - win.lo[2]];
                    size_t p = cell(g, i, j, k);
#else
                {
                    real kws = win.w[0][i - win.lo[0]] * win.w[1][j - win.lo[1]];
                    size_t p = cell(g, i, j, 0);
#endif
                    real slowness = 1.0 / (vel[p] * vel[p]);
                    real den = (1.0 + damp[p] * dt / (2 * slowness));
                    real value = dtsq / slowness * kws * wavelet[wo] / den;
#if defined(ORACLE_OMP)
                    /* the reference serialises overlapping boxes with an
                     * atomic; this loop is sequential, nothing to do */
#endif
                    next[p] += value;
                }
            }
        }
    }
}

/*
 * Phase 3: boundary conditions on one axis of the new field.
 * code 1: zero the first / last interior plane; code 2: mirror r planes into
 * the halo.  Only interior indices of the other axes are visited.
 * 3d/wave.c:304-480 (order y, x, z), 2d/wave.c:280-393 (order x, z).
 */
static void boundary_axis(const grid_t *g, real *next, int axis,
                          size_t before, size_t after)
{
    const size_t r = g->r;
    /* the two axes that are looped over */
    int oa[2], m = 0;
    for (int a = 0; a < NDIM; a++)
        if (a != axis)
            oa[m++] = a;
    if (m == 1)
        oa[1] = -1;

    const size_t n0 = g->n[oa[0]];
    const size_t n1 = (oa[1] >= 0) ? g->n[oa[1]] : 0;
    const size_t s0 = g->s[oa[0]];
    const size_t s1 = (oa[1] >= 0) ? g->s[oa[1]] : 0;
    const size_t sa = g->s[axis];
    const size_t first = r;
    const size_t last = g->n[axis] - r - 1;
    const size_t b_lo = (oa[1] >= 0) ? r : 0;
    const size_t b_hi = (oa[1] >= 0) ? n1 - r : 1;

    PARALLEL_FOR
    for (size_t a = r; a < n0 - r; a++) {
        for (size_t b = b_lo; b < b_hi; b++) {
            real *line = next + a * s0 + b * s1;
            if (before == 1)
                line[first * sa] = 0.0;
            if (before == 2)
                for (size_t ir = 1; ir <= r; ir++)
                    line[(first - ir) * sa] = line[(first + ir) * sa];
            if (after == 1)
                line[last * sa] = 0.0;
            if (after == 2)
                for (size_t ir = 1; ir <= r; ir++)
                    line[(last + ir) * sa] = line[(last - ir) * sa];
        }
    }
}

/*
 * Phase 4: sample the current field at the receivers.
 * 3d/wave.c:499-566, 2d/wave.c:412-464
 */
static void sample_receivers(const grid_t *g, const real *cur, real *row,
                             const size_t *intervals, const real *values,
                             const size_t *offsets, size_t num_receivers)
{
    PARALLEL_FOR
    for (size_t rec = 0; rec < num_receivers; rec++) {
        window_t win = window_of(g, intervals, values, offsets, rec);
        real sum = 0.0;
        for (size_t i = win.lo[0]; i <= win.hi[0]; i++) {
            for (size_t j = win.lo[1]; j <= win.hi[1]; j++) {
#if NDIM == 3
                for (size_t k = win.lo[2]; k <= win.hi[2]; k++) {
                    real kws = win.w[0][i - win.lo[0]] * win.w[1][j - win.lo[1]]
                               * win.w[2][k - win.lo[2]];
                    sum += cur[cell(g, i, j, k)] * kws;
                }
#else
                real kws = win.w[0][i - win.lo[0]] * win.w[1][j - win.lo[1]];
                sum += cur[cell(g, i, j, 0)] * kws;
#endif
            }
        }
        row[rec] = sum;
    }
}

static void swap_slots(real *a, real *b, size_t cells)
{
    PARALLEL_FOR
    for (size_t p = 0; p < cells; p++) {
        real t = a[p];
        a[p] = b[p];
        b[p] = t;
    }
}

double forward(real *u, real *velocity,
#if VARDEN
               real *density,
#endif
               real *damp,
               real *wavelet, size_t wavelet_size, size_t wavelet_count,
#if VARDEN
               real *coeff_order2, real *coeff_order1,
#else
               real *coeff,
#endif
               size_t *boundary_conditions,
               size_t *src_points_interval, size_t src_points_interval_size,
               real *src_points_values, size_t src_points_values_size,
               size_t *src_points_values_offset,
               size_t *rec_points_interval, size_t rec_points_interval_size,
               real *rec_points_values, size_t rec_points_values_size,
               size_t *rec_points_values_offset,
               real *receivers, size_t num_sources, size_t num_receivers,
               size_t nz, size_t nx,
#if NDIM == 3
               size_t ny,
#endif
               real dz, real dx,
#if NDIM == 3
               real dy,
#endif
               size_t saving_stride, real dt,
               size_t begin_timestep, size_t end_timestep,
               size_t space_order, size_t num_snapshots)
{
    struct timeval t_begin, t_end;
    gettimeofday(&t_begin, NULL);

    grid_t g;
    g.r = space_order / 2;
#if NDIM == 3
    g.n[0] = nz; g.n[1] = nx; g.n[2] = ny;
    g.s[0] = nx * ny; g.s[1] = ny; g.s[2] = 1;
    real hsq[3] = { dz * dz, dx * dx, dy * dy };
#else
    g.n[0] = nz; g.n[1] = nx; g.n[2] = 1;
    g.s[0] = nx; g.s[1] = 1; g.s[2] = 0;
    real hsq[3] = { dz * dz, dx * dx, 0 };
#endif
    g.cells = g.n[0] * g.n[1] * g.n[2];
    const real dtsq = dt * dt;

#if VARDEN
    const real *c2 = coeff_order2, *c1 = coeff_order1;
    const real *rho = density;
#else
    const real *c2 = coeff, *c1 = 0, *rho = 0;
#endif

    /* slot indices into u[num_snapshots][cells]; 3d/wave.c:47-50,113-124 */
    size_t prev_t = 0, cur_t = 1, next_t = 2;

    for (size_t n = begin_timestep; n <= end_timestep; n++) {
        if (saving_stride == 0) {
            prev_t = (n - 1) % 3;
            cur_t = n % 3;
            next_t = (n + 1) % 3;
        } else if (saving_stride == 1) {
            prev_t = n - 1;
            cur_t = n;
            next_t = n + 1;
        }

        real *prev = u + prev_t * g.cells;
        real *cur = u + cur_t * g.cells;
        real *next = u + next_t * g.cells;

        update_interior(&g, prev, cur, next, velocity, rho, damp, c2, c1,
                        hsq, dt, dtsq);

        inject_sources(&g, next, velocity, damp, wavelet, wavelet_count, n,
                       src_points_interval, src_points_values,
                       src_points_values_offset, num_sources, dt, dtsq);

        /* contiguous axis first, z last */
        for (int axis = NDIM - 1; axis >= 0; axis--)
            boundary_axis(&g, next, axis, boundary_conditions[2 * axis],
                          boundary_conditions[2 * axis + 1]);

        sample_receivers(&g, cur, receivers + (n - 1) * num_receivers,
                         rec_points_interval, rec_points_values,
                         rec_points_values_offset, num_receivers);

        /* slot bookkeeping for saving_stride > 1; 3d/wave.c:569-618 */
        if (saving_stride > 1) {
            if (n % saving_stride == 1) {
                prev_t = cur_t;
                cur_t += 1;
                next_t += 1;
                if (saving_stride % 2 == 0 && n < end_timestep) {
                    size_t t = cur_t;
                    cur_t = next_t;
                    next_t = t;
                    swap_slots(u + cur_t * g.cells, u + next_t * g.cells,
                               g.cells);
                }
            } else {
                prev_t = cur_t;
                cur_t = next_t;
                next_t = prev_t;
            }
        }
    }

    (void)wavelet_size; (void)src_points_interval_size;
    (void)src_points_values_size; (void)rec_points_interval_size;
    (void)rec_points_values_size; (void)num_snapshots;

    gettimeofday(&t_end, NULL);
    return (double)(t_end.tv_sec - t_begin.tv_sec)
         + (double)(t_end.tv_usec - t_begin.tv_usec) / 1000000.0;
}
