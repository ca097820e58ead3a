// extern "C" surface of libsimwave_b200.so (declared in include/simwave_cuda.h)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "sw_engine.h"

namespace sw {
static thread_local std::string g_lastError;
static thread_local int g_deviceOverride = -1;
static thread_local long long g_hint[4] = {0, 0, 0, 0};   // indexed by SIMWAVE_HINT_*
void set_last_error(const std::string &m) { g_lastError = m; }

static double wall()
{
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

static Options current_options()
{
    Options o = Options::from_env();
    if (g_deviceOverride >= 0)
        o.device = g_deviceOverride;
    o.zeroIn = g_hint[SIMWAVE_HINT_WAVEFIELD_IN_ZERO] != 0;
    o.outMode = (int)g_hint[SIMWAVE_HINT_WAVEFIELD_OUT];
    o.modelToken = g_hint[SIMWAVE_HINT_MODEL_RESIDENT];
    return o;
}

// The whole of `forward`: upload, time loop, drain.
static double run_forward(const simwave_problem &pb, size_t begin, size_t end)
{
    try {
        const double t0 = wall();
        std::unique_ptr<PlanBase> plan = make_plan(pb, current_options());
        if (begin <= end)
            plan->run(begin, end);
        plan->download(nullptr, nullptr);
        Timing t = plan->timing;
        const double t1 = wall();
        plan.reset();
        t.teardown = wall() - t1;
        t.total = wall() - t0;
        last_timing() = t;
        if (std::getenv("SIMWAVE_CUDA_VERBOSE"))
            std::fprintf(stderr,
                         "simwave_b200: forward %.3f s = upload %.3f + run %.3f (device loop %.3f) "
                         "+ download %.3f + teardown %.3f\n",
                         t.total, t.h2d, t.run_wall, t.loop, t.d2h, t.teardown);
        return t.total;
    } catch (const std::exception &e) {
        set_last_error(e.what());
        return -1.0;
    } catch (...) {
        set_last_error("unknown failure");
        return -1.0;
    }
}
}  // namespace sw

struct simwave_plan {
    std::unique_ptr<sw::PlanBase> impl;
};

extern "C" {

const char *simwave_cuda_last_error(void) { return sw::g_lastError.c_str(); }

const char *simwave_cuda_version(void) { return "simwave_b200 0.1 (sm_100a)"; }

int simwave_cuda_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
        return -1;
    return n;
}

int simwave_cuda_set_device(int device)
{
    if (device >= 0) {
        int n = simwave_cuda_device_count();
        if (device >= n) {
            sw::set_last_error("device ordinal out of range");
            return -1;
        }
    }
    sw::g_deviceOverride = device;
    return 0;
}

void simwave_cuda_last_timing(double *loop, double *h2d, double *d2h, double *total)
{
    const sw::Timing &t = sw::last_timing();
    if (loop) *loop = t.loop;
    if (h2d) *h2d = t.h2d;
    if (d2h) *d2h = t.d2h;
    if (total) *total = t.total;
}

int simwave_cuda_last_timing_ex(double *out, int n)
{
    const sw::Timing &t = sw::last_timing();
    const double v[6] = {t.loop, t.h2d, t.d2h, t.total, t.run_wall, t.teardown};
    for (int i = 0; i < n && i < 6; i++)
        out[i] = v[i];
    return 6;
}

void simwave_cuda_release_cache(void)
{
    sw::drop_resident_models();
    sw::release_caches();
}

int simwave_cuda_set_hint(int hint, long long value)
{
    if (hint < 1 || hint > 3 ||
        (hint == SIMWAVE_HINT_WAVEFIELD_OUT && (value < 0 || value > 2))) {
        sw::set_last_error("simwave_cuda_set_hint: unknown hint or value out of range");
        return -1;
    }
    sw::g_hint[hint] = value;
    return 0;
}

unsigned long long simwave_cuda_last_launch_count(void) { return sw::last_timing().launches; }

simwave_plan *simwave_plan_create(const simwave_problem *problem)
{
    try {
        if (!problem)
            throw sw::Error("null problem");
        std::unique_ptr<simwave_plan> p(new simwave_plan);
        p->impl = sw::make_plan(*problem, sw::current_options());
        sw::last_timing() = p->impl->timing;
        return p.release();
    } catch (const std::exception &e) {
        sw::set_last_error(e.what());
        return nullptr;
    }
}

int simwave_plan_run(simwave_plan *plan, size_t begin_timestep, size_t end_timestep,
                     double *loop_seconds)
{
    try {
        if (!plan)
            throw sw::Error("null plan");
        plan->impl->run(begin_timestep, end_timestep);
        sw::last_timing() = plan->impl->timing;
        if (loop_seconds)
            *loop_seconds = plan->impl->timing.loop;
        return 0;
    } catch (const std::exception &e) {
        sw::set_last_error(e.what());
        return -1;
    }
}

int simwave_plan_download(simwave_plan *plan, void *u, void *receivers)
{
    try {
        if (!plan)
            throw sw::Error("null plan");
        plan->impl->download(u, receivers);
        sw::last_timing() = plan->impl->timing;
        return 0;
    } catch (const std::exception &e) {
        sw::set_last_error(e.what());
        return -1;
    }
}

int simwave_plan_reset(simwave_plan *plan)
{
    try {
        if (!plan)
            throw sw::Error("null plan");
        plan->impl->reset();
        return 0;
    } catch (const std::exception &e) {
        sw::set_last_error(e.what());
        return -1;
    }
}

int simwave_plan_slab_export(simwave_plan *plan, void *desc)
{
    try {
        if (!plan || !desc)
            throw sw::Error("null plan or descriptor");
        plan->impl->slab_export(desc);
        return 0;
    } catch (const std::exception &e) {
        sw::set_last_error(e.what());
        return -1;
    }
}

int simwave_plan_slab_connect(simwave_plan *plan, const void *up_desc, const void *down_desc)
{
    try {
        if (!plan)
            throw sw::Error("null plan");
        plan->impl->slab_connect(up_desc, down_desc);
        return 0;
    } catch (const std::exception &e) {
        sw::set_last_error(e.what());
        return -1;
    }
}

void simwave_plan_destroy(simwave_plan *plan) { delete plan; }

// ---- the eight drop-in entry points -----------------------------------------
#define SW_FILL_COMMON(T)                                                          \
    simwave_problem pb;                                                            \
    std::memset(&pb, 0, sizeof(pb));                                               \
    pb.dtype_bytes = (int)sizeof(T);                                               \
    pb.u = u; pb.velocity = velocity; pb.damp = damp;                              \
    pb.wavelet = wavelet; pb.wavelet_size = wavelet_size;                          \
    pb.wavelet_count = wavelet_count;                                              \
    pb.boundary_conditions = boundary_conditions;                                  \
    pb.src_points_interval = src_points_interval;                                  \
    pb.src_points_values = src_points_values;                                      \
    pb.src_points_values_size = src_points_values_size;                            \
    pb.src_points_values_offset = src_points_values_offset;                        \
    pb.rec_points_interval = rec_points_interval;                                  \
    pb.rec_points_values = rec_points_values;                                      \
    pb.rec_points_values_size = rec_points_values_size;                            \
    pb.rec_points_values_offset = rec_points_values_offset;                        \
    pb.receivers = receivers; pb.num_sources = num_sources;                        \
    pb.num_receivers = num_receivers; pb.nz = nz; pb.nx = nx;                      \
    pb.dz = dz; pb.dx = dx; pb.saving_stride = saving_stride; pb.dt = dt;          \
    pb.space_order = space_order; pb.num_snapshots = num_snapshots;                \
    (void)src_points_interval_size; (void)rec_points_interval_size;

#define SW_TABLE_ARGS(T)                                                           \
    size_t *src_points_interval, size_t src_points_interval_size,                  \
    T *src_points_values, size_t src_points_values_size,                           \
    size_t *src_points_values_offset,                                              \
    size_t *rec_points_interval, size_t rec_points_interval_size,                  \
    T *rec_points_values, size_t rec_points_values_size,                           \
    size_t *rec_points_values_offset,                                              \
    T *receivers, size_t num_sources, size_t num_receivers

#define SW_TAIL_ARGS(T)                                                            \
    size_t saving_stride, T dt, size_t begin_timestep, size_t end_timestep,        \
    size_t space_order, size_t num_snapshots

#define SW_DEFINE_CONSTANT(NAME, T)                                                \
    double simwave_cuda_forward_2d_constant_##NAME(                                \
        T *u, T *velocity, T *damp, T *wavelet, size_t wavelet_size,               \
        size_t wavelet_count, T *coeff, size_t *boundary_conditions,               \
        SW_TABLE_ARGS(T), size_t nz, size_t nx, T dz, T dx, SW_TAIL_ARGS(T))       \
    {                                                                              \
        SW_FILL_COMMON(T)                                                          \
        pb.ndim = 2; pb.coeff_order2 = coeff;                                      \
        return sw::run_forward(pb, begin_timestep, end_timestep);                  \
    }                                                                              \
    double simwave_cuda_forward_3d_constant_##NAME(                                \
        T *u, T *velocity, T *damp, T *wavelet, size_t wavelet_size,               \
        size_t wavelet_count, T *coeff, size_t *boundary_conditions,               \
        SW_TABLE_ARGS(T), size_t nz, size_t nx, size_t ny, T dz, T dx, T dy,       \
        SW_TAIL_ARGS(T))                                                           \
    {                                                                              \
        SW_FILL_COMMON(T)                                                          \
        pb.ndim = 3; pb.ny = ny; pb.dy = dy; pb.coeff_order2 = coeff;              \
        return sw::run_forward(pb, begin_timestep, end_timestep);                  \
    }

#define SW_DEFINE_VARIABLE(NAME, T)                                                \
    double simwave_cuda_forward_2d_variable_##NAME(                                \
        T *u, T *velocity, T *density, T *damp, T *wavelet, size_t wavelet_size,   \
        size_t wavelet_count, T *coeff_order2, T *coeff_order1,                    \
        size_t *boundary_conditions, SW_TABLE_ARGS(T), size_t nz, size_t nx,       \
        T dz, T dx, SW_TAIL_ARGS(T))                                               \
    {                                                                              \
        SW_FILL_COMMON(T)                                                          \
        pb.ndim = 2; pb.density = density;                                         \
        pb.coeff_order2 = coeff_order2; pb.coeff_order1 = coeff_order1;            \
        return sw::run_forward(pb, begin_timestep, end_timestep);                  \
    }                                                                              \
    double simwave_cuda_forward_3d_variable_##NAME(                                \
        T *u, T *velocity, T *density, T *damp, T *wavelet, size_t wavelet_size,   \
        size_t wavelet_count, T *coeff_order2, T *coeff_order1,                    \
        size_t *boundary_conditions, SW_TABLE_ARGS(T), size_t nz, size_t nx,       \
        size_t ny, T dz, T dx, T dy, SW_TAIL_ARGS(T))                              \
    {                                                                              \
        SW_FILL_COMMON(T)                                                          \
        pb.ndim = 3; pb.ny = ny; pb.dy = dy; pb.density = density;                 \
        pb.coeff_order2 = coeff_order2; pb.coeff_order1 = coeff_order1;            \
        return sw::run_forward(pb, begin_timestep, end_timestep);                  \
    }

SW_DEFINE_CONSTANT(f32, float)
SW_DEFINE_CONSTANT(f64, double)
SW_DEFINE_VARIABLE(f32, float)
SW_DEFINE_VARIABLE(f64, double)

}  // extern "C"
