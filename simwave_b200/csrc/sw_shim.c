/*
 * Drop-in shim: one tiny shared library per kernel variant, each exporting the
 * single symbol `forward` that simwave's Middleware looks up
 * (simwave/kernel/backend/middleware.py:149).  Compiled eight times with
 *     -DSHIM_NDIM=2|3  -DSHIM_VARIABLE=0|1  -DSHIM_F32 | -DSHIM_F64
 * and linked against libsimwave_b200.so (rpath $ORIGIN).
 */
#include "../../include/simwave_cuda.h"

#if defined(SHIM_F32)
typedef float real;
#define SUFFIX f32
#elif defined(SHIM_F64)
typedef double real;
#define SUFFIX f64
#else
#error "define SHIM_F32 or SHIM_F64"
#endif

#if SHIM_VARIABLE
#define DENSITY_NAME variable
#else
#define DENSITY_NAME constant
#endif
#define JOIN5(a, b, c, d, e) a##b##c##d##e
#define MAKE(n, dens, suf) JOIN5(simwave_cuda_forward_, n, d_, dens, _##suf)
#define TARGET_NAME2(n, dens, suf) MAKE(n, dens, suf)
#define TARGET_NAME TARGET_NAME2(SHIM_NDIM, DENSITY_NAME, SUFFIX)

double forward(real *u, real *velocity,
#if SHIM_VARIABLE
               real *density,
#endif
               real *damp, real *wavelet, size_t wavelet_size, size_t wavelet_count,
#if SHIM_VARIABLE
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
#if SHIM_NDIM == 3
               size_t ny,
#endif
               real dz, real dx,
#if SHIM_NDIM == 3
               real dy,
#endif
               size_t saving_stride, real dt,
               size_t begin_timestep, size_t end_timestep,
               size_t space_order, size_t num_snapshots)
{
    return TARGET_NAME(u, velocity,
#if SHIM_VARIABLE
                       density,
#endif
                       damp, wavelet, wavelet_size, wavelet_count,
#if SHIM_VARIABLE
                       coeff_order2, coeff_order1,
#else
                       coeff,
#endif
                       boundary_conditions,
                       src_points_interval, src_points_interval_size,
                       src_points_values, src_points_values_size,
                       src_points_values_offset,
                       rec_points_interval, rec_points_interval_size,
                       rec_points_values, rec_points_values_size,
                       rec_points_values_offset,
                       receivers, num_sources, num_receivers,
                       nz, nx,
#if SHIM_NDIM == 3
                       ny,
#endif
                       dz, dx,
#if SHIM_NDIM == 3
                       dy,
#endif
                       saving_stride, dt, begin_timestep, end_timestep,
                       space_order, num_snapshots);
}
