/*
 * Drop-in shim: one tiny shared library per kernel variant, each exporting the
 * symbol `forward` that simwave's Middleware looks up
 * (simwave/kernel/backend/middleware.py:149) and, under the same argument
 * list, `adjoint` (include/simwave_cuda.h section 1b).  Compiled eight times with
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
#define JOIN6(a, b, c, d, e, f) a##b##c##d##e##f
#define MAKE(op, n, dens, suf) JOIN6(simwave_cuda_, op, _##n, d_, dens, _##suf)
#define TARGET_NAME2(op, n, dens, suf) MAKE(op, n, dens, suf)
#define TARGET_NAME(op) TARGET_NAME2(op, SHIM_NDIM, DENSITY_NAME, SUFFIX)

#define OPERATOR forward
#include "sw_shim_body.inc"
#undef OPERATOR
#define OPERATOR adjoint
#include "sw_shim_body.inc"
#undef OPERATOR
