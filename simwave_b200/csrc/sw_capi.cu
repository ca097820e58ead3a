// extern "C" surface of libsimwave_b200.so (declared in include/simwave_cuda.h)
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

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

// ---------------------------------------------------------------------------
// `forward` on several devices of one process: slab decomposition along z.
//
// The caller's arrays are cut into contiguous z-ranges, one per device (plane
// = nx*ny contiguous elements, so a slab of the model is a pointer offset, no
// copy); every slab gets a plan with r ghost planes per inner face, its own
// host thread, and its neighbours' device pointers (peer access, no IPC).  The
// time loop is the one of the multi-process form (sw_engine.cu: fused peer
// stores + device-side step flags).  Each slab copies the planes it owns
// straight back into the caller's `u`; receiver traces are partial sums,
// added in slab order.  Selected with SIMWAVE_CUDA_NGPUS=<n> (devices base,
// base+1, ... where base = SIMWAVE_CUDA_DEVICE or 0), SIMWAVE_CUDA_DEVICES=
// <comma list>, or simwave_cuda_set_slab_devices(); 3D, saving_stride == 0.
// ---------------------------------------------------------------------------
static thread_local std::vector<int> g_slabDevices;

static std::vector<int> slab_devices()
{
    if (!g_slabDevices.empty())
        return g_slabDevices;
    std::vector<int> out;
    if (const char *list = std::getenv("SIMWAVE_CUDA_DEVICES")) {
        for (const char *p = list; *p;) {
            char *endp = nullptr;
            const long v = std::strtol(p, &endp, 10);
            if (endp == p)
                break;
            out.push_back((int)v);
            p = (*endp == ',') ? endp + 1 : endp;
        }
        return out;
    }
    if (const char *n = std::getenv("SIMWAVE_CUDA_NGPUS")) {
        const int count = std::atoi(n);
        int base = g_deviceOverride >= 0 ? g_deviceOverride : 0;
        if (g_deviceOverride < 0)
            if (const char *d = std::getenv("SIMWAVE_CUDA_DEVICE"))
                base = std::atoi(d);
        for (int k = 0; k < count; k++)
            out.push_back(base + k);
    }
    return out;
}

namespace {
// rendezvous of the slab threads; a thread that fails leaves for good
class Rendezvous {
public:
    explicit Rendezvous(int n) : expected_(n) {}
    // false: somebody failed, give up
    bool meet()
    {
        std::unique_lock<std::mutex> lk(mu_);
        if (failed_)
            return false;
        const unsigned long long gen = generation_;
        if (++waiting_ == expected_) {
            waiting_ = 0;
            generation_++;
            cv_.notify_all();
        } else {
            cv_.wait(lk, [&] { return generation_ != gen || failed_; });
        }
        return !failed_;
    }
    void fail(const std::string &why)
    {
        std::lock_guard<std::mutex> lk(mu_);
        if (!failed_)
            error_ = why;
        failed_ = true;
        cv_.notify_all();
    }
    bool failed() { std::lock_guard<std::mutex> lk(mu_); return failed_; }
    std::string error() { std::lock_guard<std::mutex> lk(mu_); return error_; }
private:
    std::mutex mu_;
    std::condition_variable cv_;
    int expected_, waiting_ = 0;
    unsigned long long generation_ = 0;
    bool failed_ = false;
    std::string error_;
};

// Clip every window's z-interval to the planes [zLo, zHi) a slab answers for
// and shift it to local plane indices; a window that lies elsewhere becomes
// one zero-weight point (same as simwave_b200/slab.py:_clip_tables).
struct ClippedTables {
    std::vector<size_t> intervals, offsets;
    std::vector<unsigned char> values;   // element size = dtype_bytes
};
ClippedTables clip_tables(const size_t *iv, const void *values, const size_t *offsets,
                          size_t count, int elem, size_t zLo, size_t zHi, size_t zShift)
{
    ClippedTables t;
    t.intervals.assign(iv, iv + count * 6);
    t.offsets.assign(1, 0);
    const unsigned char *v = (const unsigned char *)values;
    for (size_t i = 0; i < count; i++) {
        const size_t zb = iv[i * 6], ze = iv[i * 6 + 1];
        // a window's weights are its per-axis vectors back to back; the caller's
        // offset table has num_sources entries (wave.c reads offset[src] only)
        const size_t total = (ze - zb + 1) + (iv[i * 6 + 3] - iv[i * 6 + 2] + 1) +
                             (iv[i * 6 + 5] - iv[i * 6 + 4] + 1);
        const size_t nzw = ze - zb + 1;
        const unsigned char *wz = v + offsets[i] * elem;
        const unsigned char *rest = wz + nzw * elem;
        const size_t restCount = total - nzw;
        size_t cb = std::max(zb, zLo), ce = std::min(ze, zHi - 1);
        size_t kept = 0;
        if (cb > ce || ze < zLo) {
            cb = ce = std::min(std::max(zb, zLo), zHi - 1);
            t.values.insert(t.values.end(), (size_t)elem, 0);     // one zero weight
            kept = 1;
        } else {
            t.values.insert(t.values.end(), wz + (cb - zb) * elem, wz + (ce - zb + 1) * elem);
            kept = ce - cb + 1;
        }
        t.values.insert(t.values.end(), rest, rest + restCount * elem);
        t.intervals[i * 6] = cb - zShift;
        t.intervals[i * 6 + 1] = ce - zShift;
        t.offsets.push_back(t.offsets.back() + kept + restCount);
    }
    if (t.values.empty())
        t.values.resize((size_t)elem, 0);
    return t;
}
}  // namespace

static double run_forward_slabs(const simwave_problem &pb, size_t begin, size_t end,
                                const std::vector<int> &devices)
{
    const double t0 = wall();
    const int world = (int)devices.size();
    const size_t r = pb.space_order / 2;
    const size_t nz = pb.nz, plane = pb.nx * pb.ny;
    const size_t interior = nz - 2 * r;
    const int elem = pb.dtype_bytes;
    if (interior / world < 2 * r)
        throw Error("too many devices for " + std::to_string(interior) + " interior planes");
    int have = 0;
    SW_CUDA(cudaGetDeviceCount(&have));
    for (int d : devices)
        if (d < 0 || d >= have)
            throw Error("slab device ordinal " + std::to_string(d) + " out of range");

    // owned interior ranges, as even as possible (slab.py:split_planes)
    std::vector<size_t> lo(world), hi(world);
    {
        const size_t base = interior / world, extra = interior % world;
        size_t at = r;
        for (int k = 0; k < world; k++) {
            lo[k] = at;
            at += base + ((size_t)k < extra ? 1 : 0);
            hi[k] = at;
        }
    }
    const size_t rows = pb.wavelet_size;
    std::vector<std::vector<unsigned char>> traces(world);
    std::vector<SlabPeer> peers(world);
    std::vector<Timing> timings(world);
    Rendezvous meet(world);
    const Options base = current_options();

    auto slab_thread = [&](int k) {
        try {
            Options opt = base;
            opt.device = devices[k];
            const size_t a = lo[k] - r, b = hi[k] + r;
            const bool up = k > 0, down = k < world - 1;
            const size_t ownLo = up ? lo[k] : 0, ownHi = down ? hi[k] : nz;
            simwave_problem q = pb;
            auto shifted = [&](const void *p) {
                return p ? (const void *)((const char *)p + a * plane * elem) : nullptr;
            };
            q.u = (void *)shifted(pb.u);
            q.velocity = shifted(pb.velocity);
            q.density = shifted(pb.density);
            q.damp = shifted(pb.damp);
            q.nz = b - a;
            q.u_slot_stride = nz * plane;
            q.out_plane_begin = ownLo - a;
            q.out_plane_end = ownHi - a;
            q.slab_up = up;
            q.slab_down = down;
            size_t bc[6];
            for (int i = 0; i < 6; i++) bc[i] = pb.boundary_conditions[i];
            if (up) bc[0] = 0;
            if (down) bc[1] = 0;
            q.boundary_conditions = bc;
            const ClippedTables src =
                clip_tables(pb.src_points_interval, pb.src_points_values,
                            pb.src_points_values_offset, pb.num_sources, elem, ownLo, ownHi, a);
            const ClippedTables rec =
                clip_tables(pb.rec_points_interval, pb.rec_points_values,
                            pb.rec_points_values_offset, pb.num_receivers, elem, ownLo, ownHi, a);
            q.src_points_interval = src.intervals.data();
            q.src_points_values = src.values.data();
            q.src_points_values_size = src.values.size() / elem;
            q.src_points_values_offset = src.offsets.data();
            q.rec_points_interval = rec.intervals.data();
            q.rec_points_values = rec.values.data();
            q.rec_points_values_size = rec.values.size() / elem;
            q.rec_points_values_offset = rec.offsets.data();
            traces[k].assign(std::max<size_t>(1, rows * pb.num_receivers) * elem, 0);
            q.receivers = traces[k].data();

            std::unique_ptr<PlanBase> plan = make_plan(q, opt);
            plan->prefault_outputs(end);      // this slab's part of the caller's `u`
            plan->slab_peer(&peers[k]);
            if (!meet.meet())
                return;
            plan->slab_connect_direct(up ? &peers[k - 1] : nullptr,
                                      down ? &peers[k + 1] : nullptr);
            if (!meet.meet())
                return;
            if (begin <= end)
                plan->run(begin, end);
            plan->download(nullptr, nullptr);
            timings[k] = plan->timing;
            // the neighbours' kernels store into my ghost planes until THEIR
            // loops end: nobody frees anything before everybody is done
            meet.meet();
        } catch (const std::exception &e) {
            meet.fail("slab " + std::to_string(k) + " (device " + std::to_string(devices[k]) +
                      "): " + e.what());
        } catch (...) {
            meet.fail("slab " + std::to_string(k) + ": unknown failure");
        }
    };
    std::vector<std::thread> threads;
    for (int k = 0; k < world; k++)
        threads.emplace_back([&slab_thread, k] {
            sw::widen_helper_affinity();     // one launch thread per device, not one core for all
            slab_thread(k);
        });
    for (auto &t : threads)
        t.join();
    if (meet.failed())
        throw Error(meet.error());

    // traces: partial sums of the slabs, added in slab order
    if (pb.receivers && pb.num_receivers && begin <= end) {
        const size_t first = (begin - 1) * pb.num_receivers;
        const size_t count = (end - begin + 1) * pb.num_receivers;
        if (elem == 4) {
            float *out = (float *)pb.receivers + first;
            for (size_t i = 0; i < count; i++) {
                float s = ((const float *)traces[0].data())[first + i];
                for (int k = 1; k < world; k++)
                    s += ((const float *)traces[k].data())[first + i];
                out[i] = s;
            }
        } else {
            double *out = (double *)pb.receivers + first;
            for (size_t i = 0; i < count; i++) {
                double s = ((const double *)traces[0].data())[first + i];
                for (int k = 1; k < world; k++)
                    s += ((const double *)traces[k].data())[first + i];
                out[i] = s;
            }
        }
    }
    Timing t;
    for (const Timing &x : timings) {
        t.loop = std::max(t.loop, x.loop);
        t.h2d = std::max(t.h2d, x.h2d);
        t.d2h = std::max(t.d2h, x.d2h);
        t.run_wall = std::max(t.run_wall, x.run_wall);
        t.launches += x.launches;
    }
    t.total = wall() - t0;
    last_timing() = t;
    if (std::getenv("SIMWAVE_CUDA_VERBOSE"))
        std::fprintf(stderr,
                     "simwave_b200: forward on %d devices %.3f s = upload %.3f + run %.3f "
                     "(device loop %.3f) + download %.3f (max over slabs)\n",
                     world, t.total, t.h2d, t.run_wall, t.loop, t.d2h);
    return t.total;
}

// The whole of `forward`: upload, time loop, drain.  Throws.
static double forward_impl(const simwave_problem &pb, size_t begin, size_t end)
{
    // the allocation caches keep this call's working set, nothing older
    struct CacheScope {
        CacheScope() { sw::cache_begin_call(); }
        ~CacheScope() { sw::cache_end_call(); }
    } cacheScope;
    {
        const std::vector<int> devices = slab_devices();
        if (devices.size() > 1 && pb.ndim == 3 && pb.saving_stride == 0)
            return run_forward_slabs(pb, begin, end, devices);
        const double t0 = wall();
        std::unique_ptr<PlanBase> plan = make_plan(pb, current_options());
        plan->prefault_outputs(end);      // host pages of `u`, behind the time loop
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
    }
}

template <typename F>
static double guarded(F &&body)
{
    try {
        return body();
    } catch (const std::exception &e) {
        set_last_error(e.what());
        return -1.0;
    } catch (...) {
        set_last_error("unknown failure");
        return -1.0;
    }
}

static double run_forward(const simwave_problem &pb, size_t begin, size_t end)
{
    return guarded([&] { return forward_impl(pb, begin, end); });
}

// ---------------------------------------------------------------------------
// Adjoint operator (the reference ships only
// `forward`, its tree is laid out for more: compiler.py:145-147,
// middleware.py:103-104).
//
// F maps the source wavelets w[T][nsrc] to the receiver traces d[T][nrec] of
// `forward` started from a zero wavefield.  This computes g = F^T d.
//
// One time step is U^{n+1} = P (A U^n + B U^{n-1} + S_n) with A = diag(2/D) +
// diag(s) L, s = dt^2 v^2 / D, B = -diag(N/D), S_n = diag(s) K_src^T w[n-1], P
// the boundary conditions, and d[n-1] = K_rec U^n.  Restricted to the
// interior points (the halo holds zeros or Neumann mirrors), P A P equals
// diag(s / omega) times a SYMMETRIC matrix, where omega halves the weight of a
// point for every Neumann face it lies on (the mirrored stencil counts the
// face plane's neighbours twice).  Hence M^T = diag(omega / s) M diag(s /
// omega) for the whole two-level recursion, and
//
//     F^T d = reverse( F'( reverse(d) ) )
//
// where F' is the SAME forward operator with the roles of the tables
// exchanged: the traces are injected at the receiver windows (the kernel's own
// source scaling s, weights divided by omega) and the field is sampled at the
// source windows (weights times omega).  omega is separable, so it folds into
// the per-axis table weights; windows are clipped to interior points (a
// forward window that reaches into the halo samples zeros, or mirror copies
// under Neumann: F^T is exact for windows among the interior points, which is
// what the dot-product tests use).  Constant density only: the variable-
// density operator is not self-adjoint under any diagonal weight.  No new
// kernel: the time loop, its roofline and its multi-device form are those of
// `forward`.
// ---------------------------------------------------------------------------
namespace {
template <typename T>
struct SwappedTables {
    std::vector<size_t> intervals, offsets;
    std::vector<T> values;
};

template <typename T>
SwappedTables<T> adjoint_tables(const simwave_problem &pb, const size_t *iv, const T *values,
                                const size_t *offsets, size_t count, bool inject)
{
    const int ndim = pb.ndim;
    const size_t r = pb.space_order / 2;
    const size_t ext[3] = {pb.nz, pb.nx, pb.ny};
    SwappedTables<T> t;
    t.intervals.assign(iv, iv + count * 2 * ndim);
    t.offsets.assign(1, 0);
    for (size_t i = 0; i < count; i++) {
        const T *w = values + offsets[i];
        size_t kept = 0;
        for (int ax = 0; ax < ndim; ax++) {
            const size_t b = iv[(i * ndim + ax) * 2], e = iv[(i * ndim + ax) * 2 + 1];
            const size_t lo = r, hi = ext[ax] - r - 1;
            size_t cb = std::max(b, lo), ce = std::min(e, hi);
            if (cb > ce || e < lo) {
                cb = ce = std::min(std::max(b, lo), hi);
                t.values.push_back(T(0));
                kept += 1;
            } else {
                for (size_t k = cb; k <= ce; k++) {
                    T v = w[k - b];
                    const bool onBefore = pb.boundary_conditions[2 * ax] == 2 && k == lo;
                    const bool onAfter = pb.boundary_conditions[2 * ax + 1] == 2 && k == hi;
                    for (int f = 0; f < (int)onBefore + (int)onAfter; f++)
                        v = inject ? v * T(2) : v * T(0.5);
                    t.values.push_back(v);
                }
                kept += ce - cb + 1;
            }
            t.intervals[(i * ndim + ax) * 2] = cb;
            t.intervals[(i * ndim + ax) * 2 + 1] = ce;
            w += e - b + 1;
        }
        t.offsets.push_back(t.offsets.back() + kept);
    }
    if (t.values.empty())
        t.values.push_back(T(0));
    return t;
}

template <typename T>
double adjoint_impl(const simwave_problem &pb, size_t begin, size_t end)
{
    if (pb.density)
        throw Error("adjoint operator: constant density only (the variable-density "
                    "operator is not self-adjoint)");
    if (pb.saving_stride != 0)
        throw Error("adjoint operator: saving_stride must be 0");
    if (begin < 1 || end > pb.wavelet_size || begin > end)
        throw Error("adjoint operator: timestep range outside [1, wavelet_size]");
    const size_t nsrc = pb.num_sources, nrec = pb.num_receivers;
    if (!nsrc || !nrec)
        throw Error("adjoint operator needs sources and receivers");
    const size_t steps = end - begin + 1;
    const SwappedTables<T> inj = adjoint_tables<T>(
        pb, pb.rec_points_interval, (const T *)pb.rec_points_values,
        pb.rec_points_values_offset, nrec, true);
    const SwappedTables<T> smp = adjoint_tables<T>(
        pb, pb.src_points_interval, (const T *)pb.src_points_values,
        pb.src_points_values_offset, nsrc, false);
    // the traces, last row first, as one wavelet per receiver
    const T *d = (const T *)pb.receivers;
    std::vector<T> reversed(steps * nrec);
    for (size_t i = 0; i < steps; i++)
        std::memcpy(&reversed[i * nrec], d + (end - 1 - i) * nrec, nrec * sizeof(T));
    std::vector<T> sampled(steps * nsrc, T(0));

    simwave_problem q = pb;
    q.wavelet = reversed.data();
    q.wavelet_size = steps;
    q.wavelet_count = nrec;
    q.src_points_interval = inj.intervals.data();
    q.src_points_values = inj.values.data();
    q.src_points_values_size = inj.values.size();
    q.src_points_values_offset = inj.offsets.data();
    q.rec_points_interval = smp.intervals.data();
    q.rec_points_values = smp.values.data();
    q.rec_points_values_size = smp.values.size();
    q.rec_points_values_offset = smp.offsets.data();
    q.num_sources = nrec;
    q.num_receivers = nsrc;
    q.receivers = sampled.data();
    const double seconds = forward_impl(q, 1, steps);

    // g[begin-1+i] = sampled[steps-1-i]; one shared wavelet: the sum over sources
    T *g = (T *)pb.wavelet;
    const size_t wc = pb.wavelet_count > 1 ? pb.wavelet_count : 1;
    if (wc > 1 && wc != nsrc)
        throw Error("adjoint operator: wavelet_count must be 1 or num_sources");
    for (size_t i = 0; i < steps; i++) {
        const T *row = &sampled[(steps - 1 - i) * nsrc];
        T *out = g + (begin - 1 + i) * wc;
        if (wc > 1) {
            std::memcpy(out, row, nsrc * sizeof(T));
        } else {
            T sum = T(0);
            for (size_t k = 0; k < nsrc; k++)
                sum += row[k];
            out[0] = sum;
        }
    }
    return seconds;
}
}  // namespace

static double run_adjoint(const simwave_problem &pb, size_t begin, size_t end)
{
    return guarded([&] {
        return pb.dtype_bytes == 4 ? adjoint_impl<float>(pb, begin, end)
                                   : adjoint_impl<double>(pb, begin, end);
    });
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

int simwave_cuda_set_slab_devices(const int *devices, int count)
{
    if (count < 0 || (count > 0 && !devices)) {
        sw::set_last_error("simwave_cuda_set_slab_devices: bad arguments");
        return -1;
    }
    sw::g_slabDevices.assign(devices, devices + count);
    return 0;
}

unsigned long long simwave_cuda_cached_bytes(void) { return sw::cached_device_bytes(); }

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

int simwave_cuda_last_loop_kind(void) { return sw::last_timing().loopKind; }

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

#define SW_DEFINE_CONSTANT(OP, NAME, T)                                              \
    double simwave_cuda_##OP##_2d_constant_##NAME(                                \
        T *u, T *velocity, T *damp, T *wavelet, size_t wavelet_size,               \
        size_t wavelet_count, T *coeff, size_t *boundary_conditions,               \
        SW_TABLE_ARGS(T), size_t nz, size_t nx, T dz, T dx, SW_TAIL_ARGS(T))       \
    {                                                                              \
        SW_FILL_COMMON(T)                                                          \
        pb.ndim = 2; pb.coeff_order2 = coeff;                                      \
        return sw::run_##OP(pb, begin_timestep, end_timestep);                  \
    }                                                                              \
    double simwave_cuda_##OP##_3d_constant_##NAME(                                \
        T *u, T *velocity, T *damp, T *wavelet, size_t wavelet_size,               \
        size_t wavelet_count, T *coeff, size_t *boundary_conditions,               \
        SW_TABLE_ARGS(T), size_t nz, size_t nx, size_t ny, T dz, T dx, T dy,       \
        SW_TAIL_ARGS(T))                                                           \
    {                                                                              \
        SW_FILL_COMMON(T)                                                          \
        pb.ndim = 3; pb.ny = ny; pb.dy = dy; pb.coeff_order2 = coeff;              \
        return sw::run_##OP(pb, begin_timestep, end_timestep);                  \
    }

#define SW_DEFINE_VARIABLE(OP, NAME, T)                                              \
    double simwave_cuda_##OP##_2d_variable_##NAME(                                \
        T *u, T *velocity, T *density, T *damp, T *wavelet, size_t wavelet_size,   \
        size_t wavelet_count, T *coeff_order2, T *coeff_order1,                    \
        size_t *boundary_conditions, SW_TABLE_ARGS(T), size_t nz, size_t nx,       \
        T dz, T dx, SW_TAIL_ARGS(T))                                               \
    {                                                                              \
        SW_FILL_COMMON(T)                                                          \
        pb.ndim = 2; pb.density = density;                                         \
        pb.coeff_order2 = coeff_order2; pb.coeff_order1 = coeff_order1;            \
        return sw::run_##OP(pb, begin_timestep, end_timestep);                  \
    }                                                                              \
    double simwave_cuda_##OP##_3d_variable_##NAME(                                \
        T *u, T *velocity, T *density, T *damp, T *wavelet, size_t wavelet_size,   \
        size_t wavelet_count, T *coeff_order2, T *coeff_order1,                    \
        size_t *boundary_conditions, SW_TABLE_ARGS(T), size_t nz, size_t nx,       \
        size_t ny, T dz, T dx, T dy, SW_TAIL_ARGS(T))                              \
    {                                                                              \
        SW_FILL_COMMON(T)                                                          \
        pb.ndim = 3; pb.ny = ny; pb.dy = dy; pb.density = density;                 \
        pb.coeff_order2 = coeff_order2; pb.coeff_order1 = coeff_order1;            \
        return sw::run_##OP(pb, begin_timestep, end_timestep);                  \
    }

SW_DEFINE_CONSTANT(forward, f32, float)
SW_DEFINE_CONSTANT(forward, f64, double)
SW_DEFINE_VARIABLE(forward, f32, float)
SW_DEFINE_VARIABLE(forward, f64, double)
// the adjoint operator: same argument lists; `receivers` is the input,
// `wavelet` the output (variable density is refused with an error message)
SW_DEFINE_CONSTANT(adjoint, f32, float)
SW_DEFINE_CONSTANT(adjoint, f64, double)
SW_DEFINE_VARIABLE(adjoint, f32, float)
SW_DEFINE_VARIABLE(adjoint, f64, double)

}  // extern "C"
