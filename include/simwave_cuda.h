/*
 * simwave_cuda.h -- C ABI of the B200 (sm_100a) backend for simwave's acoustic
 * forward-modelling time loop.
 *
 * Two layers:
 *
 * 1. The drop-in boundary.  Eight shared libraries
 *        libsimwave_cuda_<2|3>d_<constant|variable>_<f32|f64>.so
 *    each export ONE symbol, `forward`, with exactly the argument list of the
 *    reference kernel it replaces, so simwave's Middleware can bind it with the
 *    argtypes it already builds (simwave/kernel/backend/middleware.py:107-158,
 *    argument order :166-202, restype c_double :150).  Which variant is used is
 *    decided by file, as in the reference (middleware.py:26-51).  The shims
 *    forward to the named entry points below, which live in the core library
 *    libsimwave_b200.so.
 *
 *    Every pointer is a caller-owned host buffer valid for the duration of the
 *    call; `u` and `receivers` are updated in place; nothing is retained.
 *    Integers are size_t (ctypes c_size_t), index arrays are size_t* (NumPy
 *    uint64).  Return value: elapsed wall-clock seconds (>= 0), like the
 *    reference (constant_density/3d/wave.c:57,658-662).  On failure the
 *    functions return -1.0 and keep a message retrievable with
 *    simwave_cuda_last_error(); they never call exit() (the reference CUDA
 *    path does: constant_density/3d/cuda/wave.cu:13-20) and never throw.
 *
 * 2. Side exports (device selection, timing of the last call, a plan API that
 *    keeps a problem resident on the device so the time loop can be timed
 *    without host<->device traffic, and the slab-decomposition hooks).
 *    The `forward` signature is frozen, so every new knob lives here or in
 *    environment variables:
 *        SIMWAVE_CUDA_DEVICE   device ordinal used by `forward` (default: current)
 *        SIMWAVE_CUDA_MATH     fast | strict  (default fast; strict is bit-identical
 *                              to the reference's sequential C kernel)
 *        SIMWAVE_CUDA_KERNEL   auto | simple  (auto picks the tiled kernels)
 *        SIMWAVE_CUDA_DEBUG    1 = synchronise and check after every launch
 */
#ifndef SIMWAVE_CUDA_H
#define SIMWAVE_CUDA_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------
 * 1. drop-in entry points (one per reference kernel file x precision)
 * ---------------------------------------------------------------------- */

/* replaces constant_density/2d/wave.c:23-36 built with -DFLOAT */
double simwave_cuda_forward_2d_constant_f32(
    float *u, float *velocity, float *damp,
    float *wavelet, size_t wavelet_size, size_t wavelet_count,
    float *coeff, size_t *boundary_conditions,
    size_t *src_points_interval, size_t src_points_interval_size,
    float *src_points_values, size_t src_points_values_size,
    size_t *src_points_values_offset,
    size_t *rec_points_interval, size_t rec_points_interval_size,
    float *rec_points_values, size_t rec_points_values_size,
    size_t *rec_points_values_offset,
    float *receivers, size_t num_sources, size_t num_receivers,
    size_t nz, size_t nx, float dz, float dx,
    size_t saving_stride, float dt,
    size_t begin_timestep, size_t end_timestep,
    size_t space_order, size_t num_snapshots);

/* replaces constant_density/2d/wave.c:23-36 built with -DDOUBLE */
double simwave_cuda_forward_2d_constant_f64(
    double *u, double *velocity, double *damp,
    double *wavelet, size_t wavelet_size, size_t wavelet_count,
    double *coeff, size_t *boundary_conditions,
    size_t *src_points_interval, size_t src_points_interval_size,
    double *src_points_values, size_t src_points_values_size,
    size_t *src_points_values_offset,
    size_t *rec_points_interval, size_t rec_points_interval_size,
    double *rec_points_values, size_t rec_points_values_size,
    size_t *rec_points_values_offset,
    double *receivers, size_t num_sources, size_t num_receivers,
    size_t nz, size_t nx, double dz, double dx,
    size_t saving_stride, double dt,
    size_t begin_timestep, size_t end_timestep,
    size_t space_order, size_t num_snapshots);

/* replaces constant_density/3d/wave.c:23-36 built with -DFLOAT */
double simwave_cuda_forward_3d_constant_f32(
    float *u, float *velocity, float *damp,
    float *wavelet, size_t wavelet_size, size_t wavelet_count,
    float *coeff, size_t *boundary_conditions,
    size_t *src_points_interval, size_t src_points_interval_size,
    float *src_points_values, size_t src_points_values_size,
    size_t *src_points_values_offset,
    size_t *rec_points_interval, size_t rec_points_interval_size,
    float *rec_points_values, size_t rec_points_values_size,
    size_t *rec_points_values_offset,
    float *receivers, size_t num_sources, size_t num_receivers,
    size_t nz, size_t nx, size_t ny, float dz, float dx, float dy,
    size_t saving_stride, float dt,
    size_t begin_timestep, size_t end_timestep,
    size_t space_order, size_t num_snapshots);

/* replaces constant_density/3d/wave.c:23-36 built with -DDOUBLE */
double simwave_cuda_forward_3d_constant_f64(
    double *u, double *velocity, double *damp,
    double *wavelet, size_t wavelet_size, size_t wavelet_count,
    double *coeff, size_t *boundary_conditions,
    size_t *src_points_interval, size_t src_points_interval_size,
    double *src_points_values, size_t src_points_values_size,
    size_t *src_points_values_offset,
    size_t *rec_points_interval, size_t rec_points_interval_size,
    double *rec_points_values, size_t rec_points_values_size,
    size_t *rec_points_values_offset,
    double *receivers, size_t num_sources, size_t num_receivers,
    size_t nz, size_t nx, size_t ny, double dz, double dx, double dy,
    size_t saving_stride, double dt,
    size_t begin_timestep, size_t end_timestep,
    size_t space_order, size_t num_snapshots);

/* replaces variable_density/2d/wave.c:23-36 built with -DFLOAT */
double simwave_cuda_forward_2d_variable_f32(
    float *u, float *velocity, float *density, float *damp,
    float *wavelet, size_t wavelet_size, size_t wavelet_count,
    float *coeff_order2, float *coeff_order1, size_t *boundary_conditions,
    size_t *src_points_interval, size_t src_points_interval_size,
    float *src_points_values, size_t src_points_values_size,
    size_t *src_points_values_offset,
    size_t *rec_points_interval, size_t rec_points_interval_size,
    float *rec_points_values, size_t rec_points_values_size,
    size_t *rec_points_values_offset,
    float *receivers, size_t num_sources, size_t num_receivers,
    size_t nz, size_t nx, float dz, float dx,
    size_t saving_stride, float dt,
    size_t begin_timestep, size_t end_timestep,
    size_t space_order, size_t num_snapshots);

/* replaces variable_density/2d/wave.c:23-36 built with -DDOUBLE */
double simwave_cuda_forward_2d_variable_f64(
    double *u, double *velocity, double *density, double *damp,
    double *wavelet, size_t wavelet_size, size_t wavelet_count,
    double *coeff_order2, double *coeff_order1, size_t *boundary_conditions,
    size_t *src_points_interval, size_t src_points_interval_size,
    double *src_points_values, size_t src_points_values_size,
    size_t *src_points_values_offset,
    size_t *rec_points_interval, size_t rec_points_interval_size,
    double *rec_points_values, size_t rec_points_values_size,
    size_t *rec_points_values_offset,
    double *receivers, size_t num_sources, size_t num_receivers,
    size_t nz, size_t nx, double dz, double dx,
    size_t saving_stride, double dt,
    size_t begin_timestep, size_t end_timestep,
    size_t space_order, size_t num_snapshots);

/* replaces variable_density/3d/wave.c:23-36 built with -DFLOAT */
double simwave_cuda_forward_3d_variable_f32(
    float *u, float *velocity, float *density, float *damp,
    float *wavelet, size_t wavelet_size, size_t wavelet_count,
    float *coeff_order2, float *coeff_order1, size_t *boundary_conditions,
    size_t *src_points_interval, size_t src_points_interval_size,
    float *src_points_values, size_t src_points_values_size,
    size_t *src_points_values_offset,
    size_t *rec_points_interval, size_t rec_points_interval_size,
    float *rec_points_values, size_t rec_points_values_size,
    size_t *rec_points_values_offset,
    float *receivers, size_t num_sources, size_t num_receivers,
    size_t nz, size_t nx, size_t ny, float dz, float dx, float dy,
    size_t saving_stride, float dt,
    size_t begin_timestep, size_t end_timestep,
    size_t space_order, size_t num_snapshots);

/* replaces variable_density/3d/wave.c:23-36 built with -DDOUBLE */
double simwave_cuda_forward_3d_variable_f64(
    double *u, double *velocity, double *density, double *damp,
    double *wavelet, size_t wavelet_size, size_t wavelet_count,
    double *coeff_order2, double *coeff_order1, size_t *boundary_conditions,
    size_t *src_points_interval, size_t src_points_interval_size,
    double *src_points_values, size_t src_points_values_size,
    size_t *src_points_values_offset,
    size_t *rec_points_interval, size_t rec_points_interval_size,
    double *rec_points_values, size_t rec_points_values_size,
    size_t *rec_points_values_offset,
    double *receivers, size_t num_sources, size_t num_receivers,
    size_t nz, size_t nx, size_t ny, double dz, double dx, double dy,
    size_t saving_stride, double dt,
    size_t begin_timestep, size_t end_timestep,
    size_t space_order, size_t num_snapshots);

/* ------------------------------------------------------------------------
 * 1b. adjoint operator (constant density)
 *
 * New: the reference ships only `forward`; its source tree and Middleware are
 * laid out per operator (simwave/kernel/backend/compiler.py:145-147
 * c_code/<operator>/..., middleware.py:103-104), and an `adjoint` is what an
 * FWI user needs next.  Same argument lists as the
 * forward entry points above (constant_density/{2,3}d/wave.c:23-36), with the
 * roles of two arrays exchanged:
 *     receivers [wavelet_size][num_receivers]  INPUT: the data d
 *     wavelet   [wavelet_size][wavelet_count]  OUTPUT: g = F^T d, where F is the
 *               linear map wavelet -> receivers of `forward` started from a
 *               zero wavefield over [begin_timestep, end_timestep] (rows
 *               outside that range are left untouched; wavelet_count == 1 gives
 *               the sum over sources, the transpose of one shared wavelet)
 *     u         in/out like `forward`: must be zero on entry; returns the last
 *               three adjoint wavefields (scaled by dt^2 v^2 / D)
 * Exact transpose (float64: <F w, d> = <w, F^T d> to rounding) for source and
 * receiver windows among the interior points, every boundary-condition mix and
 * damping layers; windows reaching into the halo are clipped to the interior.
 * The variable-density variants exist and return -1 ("constant density only").
 * The shims export them as `adjoint`, next to `forward`.
 * ---------------------------------------------------------------------- */
double simwave_cuda_adjoint_2d_constant_f32(
    float *u, float *velocity, float *damp,
    float *wavelet, size_t wavelet_size, size_t wavelet_count,
    float *coeff, size_t *boundary_conditions,
    size_t *src_points_interval, size_t src_points_interval_size,
    float *src_points_values, size_t src_points_values_size,
    size_t *src_points_values_offset,
    size_t *rec_points_interval, size_t rec_points_interval_size,
    float *rec_points_values, size_t rec_points_values_size,
    size_t *rec_points_values_offset,
    float *receivers, size_t num_sources, size_t num_receivers,
    size_t nz, size_t nx, float dz, float dx,
    size_t saving_stride, float dt,
    size_t begin_timestep, size_t end_timestep,
    size_t space_order, size_t num_snapshots);

double simwave_cuda_adjoint_2d_constant_f64(
    double *u, double *velocity, double *damp,
    double *wavelet, size_t wavelet_size, size_t wavelet_count,
    double *coeff, size_t *boundary_conditions,
    size_t *src_points_interval, size_t src_points_interval_size,
    double *src_points_values, size_t src_points_values_size,
    size_t *src_points_values_offset,
    size_t *rec_points_interval, size_t rec_points_interval_size,
    double *rec_points_values, size_t rec_points_values_size,
    size_t *rec_points_values_offset,
    double *receivers, size_t num_sources, size_t num_receivers,
    size_t nz, size_t nx, double dz, double dx,
    size_t saving_stride, double dt,
    size_t begin_timestep, size_t end_timestep,
    size_t space_order, size_t num_snapshots);

double simwave_cuda_adjoint_3d_constant_f32(
    float *u, float *velocity, float *damp,
    float *wavelet, size_t wavelet_size, size_t wavelet_count,
    float *coeff, size_t *boundary_conditions,
    size_t *src_points_interval, size_t src_points_interval_size,
    float *src_points_values, size_t src_points_values_size,
    size_t *src_points_values_offset,
    size_t *rec_points_interval, size_t rec_points_interval_size,
    float *rec_points_values, size_t rec_points_values_size,
    size_t *rec_points_values_offset,
    float *receivers, size_t num_sources, size_t num_receivers,
    size_t nz, size_t nx, size_t ny, float dz, float dx, float dy,
    size_t saving_stride, float dt,
    size_t begin_timestep, size_t end_timestep,
    size_t space_order, size_t num_snapshots);

double simwave_cuda_adjoint_3d_constant_f64(
    double *u, double *velocity, double *damp,
    double *wavelet, size_t wavelet_size, size_t wavelet_count,
    double *coeff, size_t *boundary_conditions,
    size_t *src_points_interval, size_t src_points_interval_size,
    double *src_points_values, size_t src_points_values_size,
    size_t *src_points_values_offset,
    size_t *rec_points_interval, size_t rec_points_interval_size,
    double *rec_points_values, size_t rec_points_values_size,
    size_t *rec_points_values_offset,
    double *receivers, size_t num_sources, size_t num_receivers,
    size_t nz, size_t nx, size_t ny, double dz, double dx, double dy,
    size_t saving_stride, double dt,
    size_t begin_timestep, size_t end_timestep,
    size_t space_order, size_t num_snapshots);

/* variable-density adjoint: declared for the shims, always refused (-1) */
double simwave_cuda_adjoint_2d_variable_f32(
    float *u, float *velocity, float *density, float *damp,
    float *wavelet, size_t wavelet_size, size_t wavelet_count,
    float *coeff_order2, float *coeff_order1, size_t *boundary_conditions,
    size_t *src_points_interval, size_t src_points_interval_size,
    float *src_points_values, size_t src_points_values_size,
    size_t *src_points_values_offset,
    size_t *rec_points_interval, size_t rec_points_interval_size,
    float *rec_points_values, size_t rec_points_values_size,
    size_t *rec_points_values_offset,
    float *receivers, size_t num_sources, size_t num_receivers,
    size_t nz, size_t nx, float dz, float dx,
    size_t saving_stride, float dt,
    size_t begin_timestep, size_t end_timestep,
    size_t space_order, size_t num_snapshots);

double simwave_cuda_adjoint_2d_variable_f64(
    double *u, double *velocity, double *density, double *damp,
    double *wavelet, size_t wavelet_size, size_t wavelet_count,
    double *coeff_order2, double *coeff_order1, size_t *boundary_conditions,
    size_t *src_points_interval, size_t src_points_interval_size,
    double *src_points_values, size_t src_points_values_size,
    size_t *src_points_values_offset,
    size_t *rec_points_interval, size_t rec_points_interval_size,
    double *rec_points_values, size_t rec_points_values_size,
    size_t *rec_points_values_offset,
    double *receivers, size_t num_sources, size_t num_receivers,
    size_t nz, size_t nx, double dz, double dx,
    size_t saving_stride, double dt,
    size_t begin_timestep, size_t end_timestep,
    size_t space_order, size_t num_snapshots);

double simwave_cuda_adjoint_3d_variable_f32(
    float *u, float *velocity, float *density, float *damp,
    float *wavelet, size_t wavelet_size, size_t wavelet_count,
    float *coeff_order2, float *coeff_order1, size_t *boundary_conditions,
    size_t *src_points_interval, size_t src_points_interval_size,
    float *src_points_values, size_t src_points_values_size,
    size_t *src_points_values_offset,
    size_t *rec_points_interval, size_t rec_points_interval_size,
    float *rec_points_values, size_t rec_points_values_size,
    size_t *rec_points_values_offset,
    float *receivers, size_t num_sources, size_t num_receivers,
    size_t nz, size_t nx, size_t ny, float dz, float dx, float dy,
    size_t saving_stride, float dt,
    size_t begin_timestep, size_t end_timestep,
    size_t space_order, size_t num_snapshots);

double simwave_cuda_adjoint_3d_variable_f64(
    double *u, double *velocity, double *density, double *damp,
    double *wavelet, size_t wavelet_size, size_t wavelet_count,
    double *coeff_order2, double *coeff_order1, size_t *boundary_conditions,
    size_t *src_points_interval, size_t src_points_interval_size,
    double *src_points_values, size_t src_points_values_size,
    size_t *src_points_values_offset,
    size_t *rec_points_interval, size_t rec_points_interval_size,
    double *rec_points_values, size_t rec_points_values_size,
    size_t *rec_points_values_offset,
    double *receivers, size_t num_sources, size_t num_receivers,
    size_t nz, size_t nx, size_t ny, double dz, double dx, double dy,
    size_t saving_stride, double dt,
    size_t begin_timestep, size_t end_timestep,
    size_t space_order, size_t num_snapshots);

/* ------------------------------------------------------------------------
 * 2. side exports
 * ---------------------------------------------------------------------- */

/* Message of the last failure on the calling thread ("" if none). */
const char *simwave_cuda_last_error(void);

/* Library version string, e.g. "simwave_b200 0.1 (sm_100a)". */
const char *simwave_cuda_version(void);

/* Number of CUDA devices visible (<= 0: none / driver error). */
int simwave_cuda_device_count(void);

/* Device used by subsequent calls from this thread; -1 restores the default
 * (SIMWAVE_CUDA_DEVICE, else the current device).  Returns 0 on success. */
int simwave_cuda_set_device(int device);

/* Timing breakdown of the last successful forward()/plan call on this thread,
 * in seconds; any pointer may be NULL.
 *   loop   device time of the time loop alone (CUDA events)
 *   h2d    uploading and preparing the model and the initial fields
 *   d2h    draining wavefield slots and receivers after the loop
 *   total  wall clock of the whole call (what forward() returns)           */
void simwave_cuda_last_timing(double *loop, double *h2d, double *d2h,
                              double *total);

/* Extended breakdown: fills out[0..n) with {loop, h2d, d2h, total, run_wall,
 * teardown} (seconds; run_wall = host wall clock of the time loop including
 * launch overhead, teardown = releasing device and pinned memory); returns the
 * number of values available. */
int simwave_cuda_last_timing_ex(double *out, int n);

/*
 * `forward` over several devices of this process (new functionality: the
 * reference is single-device).  A 3D problem with
 * saving_stride == 0 is cut into z-slabs, one per listed device, each with its
 * own host thread inside the call; ghost planes travel device to device over
 * NVLink (peer stores fused into the step kernel, device-side step flags);
 * every slab copies the planes it owns back into the caller's `u`, traces are
 * summed over slabs.  The wavefield is bit-identical to the single-device run.
 * Also selectable without code: SIMWAVE_CUDA_NGPUS=<n> (devices base .. base +
 * n - 1, base = SIMWAVE_CUDA_DEVICE or 0) or SIMWAVE_CUDA_DEVICES=<d0,d1,..>.
 * count == 0 returns this thread to the environment's setting.
 */
int simwave_cuda_set_slab_devices(const int *devices, int count);

/* Number of kernels launched by the last forward()/plan run on this thread. */
unsigned long long simwave_cuda_last_launch_count(void);

/* How the time loop of the last forward()/plan run on this thread was driven:
 * 0 = kernels launched per time step, 1 = the persistent 2D loop with a grid
 * barrier per step, 2 = the tile-resident 2D loop (wavefields and model kept in
 * shared memory, halo strips exchanged between neighbouring tiles). */
int simwave_cuda_last_loop_kind(void);

/* Device buffers and pinned staging buffers are kept between calls (a survey
 * calls forward() once per shot with the same shapes); this hands every cached
 * block back to the driver.  The cache holds the working set of the last
 * forward() only (see simwave_cuda_cached_bytes), is bounded by half of the
 * device memory and 512 MiB of pinned memory, is emptied automatically when an
 * allocation fails, and is off altogether with SIMWAVE_CUDA_CACHE=0. */
void simwave_cuda_release_cache(void);

/* Device memory the cache holds right now, in bytes, over all devices.  After
 * a forward() this is the working set of that call and nothing older: blocks
 * an earlier call left behind and this one did not take again are freed when
 * it returns (SIMWAVE_CUDA_CACHE=keep keeps everything, =0 nothing). */
unsigned long long simwave_cuda_cached_bytes(void);

/*
 * Hints: promises of the caller about the next forward() calls on this thread
 * (the signature of `forward` is frozen, SURVEY.md section 8b, so they travel
 * beside it).  Every hint is exact -- the arithmetic and the results that are
 * copied back do not change -- and off (0) by default; a hint stays in force
 * until it is set again.  They cover the host <-> device data path of a shot
 * (SURVEY.md section 8 f2), which the reference's GPU variants pay in full on
 * every call (constant_density/3d/wave.c:69-80, :623-624; cuda/wave.cu:498-543,
 * :700-704; solver.py:101-110 allocates a fresh zero `u` per call).
 *
 *  SIMWAVE_HINT_WAVEFIELD_IN_ZERO   value != 0: every slot of `u` is zero on
 *      entry (what simwave's Solver passes): nothing of `u` is read or uploaded.
 *  SIMWAVE_HINT_WAVEFIELD_OUT       with saving_stride == 0, which of the three
 *      rotating slots are copied back into `u`: 0 all (the ABI's contract),
 *      1 only slot end_timestep % 3, the wavefield simwave's Solver returns
 *      (model.py:639-641), 2 none (receiver traces only).  The other slots of
 *      `u` are left untouched.  Ignored when saving_stride > 0.
 *  SIMWAVE_HINT_MODEL_RESIDENT      value = a non-zero token chosen by the
 *      caller: velocity / damp / density, the grid, dt and space_order are the
 *      same for every call made under this token, so the device keeps the
 *      preprocessed model of the last call (per device) and the next one skips
 *      its upload.  0 switches the lookup off for the following calls; the
 *      device copy goes when a call under another token replaces it or with
 *      simwave_cuda_release_cache().
 *
 * Returns 0, or -1 for an unknown hint / value (see simwave_cuda_last_error).
 */
#define SIMWAVE_HINT_WAVEFIELD_IN_ZERO 1
#define SIMWAVE_HINT_WAVEFIELD_OUT 2
#define SIMWAVE_HINT_MODEL_RESIDENT 3
int simwave_cuda_set_hint(int hint, long long value);

/*
 * Plan API: a problem kept resident on one device.
 *
 * The descriptor carries the same information as the `forward` argument list
 * (field names follow it); `dtype_bytes` is 4 or 8, `ndim` 2 or 3, `density`
 * / `coeff_order1` are NULL for constant density.  simwave_plan_create()
 * uploads everything (model, tables, wavelet, the initial slots of `u`) and
 * returns NULL on failure.  simwave_plan_run() advances the time loop over
 * [begin_timestep, end_timestep] exactly as `forward` would and reports the
 * device time of the loop.  simwave_plan_download() writes the wavefield
 * slots and receiver rows produced so far into caller buffers laid out like
 * the `forward` arguments.  A plan belongs to the thread that created it.
 */
typedef struct simwave_plan simwave_plan;

typedef struct simwave_problem {
    int ndim;                 /* 2 or 3                                        */
    int dtype_bytes;          /* 4 = float, 8 = double                         */
    void *u;                  /* [num_snapshots][nz][nx]([ny]) in/out          */
    const void *velocity;     /* [nz][nx]([ny])                                */
    const void *density;      /* same shape or NULL                            */
    const void *damp;         /* same shape                                    */
    const void *wavelet;      /* [wavelet_size][wavelet_count]                 */
    size_t wavelet_size, wavelet_count;
    const void *coeff_order2; /* [space_order/2 + 1]                           */
    const void *coeff_order1; /* same length or NULL                           */
    const size_t *boundary_conditions;      /* [2*ndim]                       */
    const size_t *src_points_interval;      /* [num_sources][2*ndim]          */
    const void *src_points_values;
    size_t src_points_values_size;
    const size_t *src_points_values_offset; /* [num_sources] (only these are read) */
    const size_t *rec_points_interval;
    const void *rec_points_values;
    size_t rec_points_values_size;
    const size_t *rec_points_values_offset;
    void *receivers;          /* [wavelet_size][num_receivers] in/out          */
    size_t num_sources, num_receivers;
    size_t nz, nx, ny;        /* ny ignored in 2D                              */
    double dz, dx, dy;        /* already rounded to dtype by the caller        */
    size_t saving_stride;
    double dt;                /* already rounded to dtype by the caller        */
    size_t space_order;
    size_t num_snapshots;
    /* Slab decomposition along z (0 = none).  When set, the first / last
     * space_order/2 planes of every array are GHOST planes owned by the
     * neighbouring slab (lower z = "up", higher z = "down"); the time loop
     * refreshes them every step from the neighbour's device over NVLink.
     * Needs saving_stride == 0 and a connected plan (below). */
    int slab_up, slab_down;
    /* A slab whose host arrays are windows of larger ones (0 = dense): distance
     * in elements between consecutive slots of `u`, and the range of local
     * planes [out_plane_begin, out_plane_end) that simwave_plan_download()
     * writes back (0, 0 = all of them).  Used by the multi-device `forward`. */
    size_t u_slot_stride;
    size_t out_plane_begin, out_plane_end;
} simwave_problem;

simwave_plan *simwave_plan_create(const simwave_problem *problem);
int simwave_plan_run(simwave_plan *plan, size_t begin_timestep,
                     size_t end_timestep, double *loop_seconds);
int simwave_plan_download(simwave_plan *plan, void *u, void *receivers);
/* Reset the wavefield slots to the contents of `u` given at creation (zeros
 * if that was all zero) so the same plan can be timed repeatedly. */
int simwave_plan_reset(simwave_plan *plan);
void simwave_plan_destroy(simwave_plan *plan);

/*
 * Slab decomposition (one process per GPU).  Each process creates a plan for
 * its z-slab (ghost planes included, slab_up / slab_down set), exports an
 * opaque descriptor of its wavefield buffers (CUDA IPC handles), exchanges
 * descriptors with its neighbours by any host channel (torch.distributed,
 * MPI, files ...) and connects.  From then on simwave_plan_run() keeps the
 * ghost planes current: after every step the slab's outermost owned planes are
 * written into the neighbours' ghost planes through the peer mapping and a
 * per-step flag is published in the neighbour's memory; the next step waits on
 * the flags on the device.  No host synchronisation inside the time loop.
 * All slabs must call simwave_plan_run() with the same timestep range, and a
 * host barrier must separate simwave_plan_reset() / connect from the next run.
 */
#define SIMWAVE_SLAB_DESC_BYTES 512
int simwave_plan_slab_export(simwave_plan *plan, void *desc /* SIMWAVE_SLAB_DESC_BYTES */);
/* `up_desc` / `down_desc`: descriptors exported by the neighbouring processes
 * (NULL where the problem has no such neighbour). */
int simwave_plan_slab_connect(simwave_plan *plan, const void *up_desc,
                              const void *down_desc);

#ifdef __cplusplus
}
#endif
#endif /* SIMWAVE_CUDA_H */
