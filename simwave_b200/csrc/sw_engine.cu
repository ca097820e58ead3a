// Time-loop engine: uploads a problem, runs the reference's loop structure on
// the device and drains the results.  See DESIGN.md ("Engine") for the slot
// model; the short version: the reference keeps all fields in one array
// u[num_snapshots][cells] and moves three slot indices around it
// (constant_density/3d/wave.c:47-50,113-124,569-618).  We execute exactly that
// index logic, but a slot only owns a device buffer while the loop can still
// touch it; once the indices have moved past a slot its field is drained to
// the caller's array on a side stream and the buffer is recycled.  Slot
// identity is preserved, so halo cells (which the reference never clears) see
// the same history as in the reference.
#include "sw_engine.h"

#include <sched.h>
#include <sys/mman.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <future>
#include <set>

#include <type_traits>

#include <nvtx3/nvToolsExt.h>

#include "sw_launch.h"
#include "sw_points.cuh"

namespace sw {

static double wall()
{
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

// phase log of the upload (printed under SIMWAVE_CUDA_VERBOSE)
// NVTX range for the lifetime of the object: the phases of a call show up in
// Nsight Systems / ncu timelines (costs nothing without a tool attached)
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange &) = delete;
    NvtxRange &operator=(const NvtxRange &) = delete;
};

struct PhaseLog {
    bool on;
    double t;
    std::string out;
    PhaseLog() : on(std::getenv("SIMWAVE_CUDA_VERBOSE") != nullptr), t(wall()) {}
    void mark(const char *what)
    {
        nvtxMarkA(what);
        if (!on) return;
        const double now = wall();
        char buf[96];
        std::snprintf(buf, sizeof(buf), " %s %.1f ms;", what, 1e3 * (now - t));
        out += buf;
        t = now;
    }
    ~PhaseLog()
    {
        if (on && !out.empty())
            std::fprintf(stderr, "simwave_b200: upload phases:%s\n", out.c_str());
    }
};

static thread_local Timing g_lastTiming;
Timing &last_timing() { return g_lastTiming; }

static bool env_is(const char *name, const char *value)
{
    const char *v = std::getenv(name);
    return v && std::strcmp(v, value) == 0;
}

Options Options::from_env()
{
    Options o;
    o.math = env_is("SIMWAVE_CUDA_MATH", "strict") ? MATH_STRICT : MATH_FAST;
    o.simple = env_is("SIMWAVE_CUDA_KERNEL", "simple");
    o.debug = env_is("SIMWAVE_CUDA_DEBUG", "1");
    o.separateBc = env_is("SIMWAVE_CUDA_BC", "separate");
    o.perStep = env_is("SIMWAVE_CUDA_LOOP", "launch");
    o.prefetch = -1;
    if (const char *d = std::getenv("SIMWAVE_CUDA_PREFETCH"))
        o.prefetch = std::atoi(d);
    o.device = -1;
    if (const char *d = std::getenv("SIMWAVE_CUDA_DEVICE"))
        o.device = std::atoi(d);
    return o;
}

// ---------------------------------------------------------------------------
// allocation caches
// ---------------------------------------------------------------------------
namespace {
struct CachedBlock {
    void *p;
    unsigned long long gen;     // forward() call during which the block came back
};
struct Caches {
    std::mutex mu;
    std::multimap<std::pair<int, size_t>, CachedBlock> device;   // (ordinal, bytes) -> block
    std::map<int, size_t> deviceBytes, deviceLimit;
    std::multimap<size_t, CachedBlock> pinned;
    size_t pinnedBytes = 0;
    unsigned long long gen = 0;
    // allocations a peer process may still have mapped through CUDA IPC: freed
    // outright, never recycled
    std::set<void *> exported;
    bool enabled = !env_is("SIMWAVE_CUDA_CACHE", "0");
    // SIMWAVE_CUDA_CACHE=keep: blocks stay until simwave_cuda_release_cache()
    bool keepAll = env_is("SIMWAVE_CUDA_CACHE", "keep");
};
// never destroyed: no CUDA calls during static destruction
Caches &caches()
{
    static Caches *c = new Caches;
    return *c;
}

void *device_take(size_t bytes, int *deviceOut)
{
    int dev = 0;
    SW_CUDA(cudaGetDevice(&dev));
    *deviceOut = dev;
    Caches &c = caches();
    {
        std::lock_guard<std::mutex> lk(c.mu);
        auto it = c.device.find(std::make_pair(dev, bytes));
        if (it != c.device.end()) {
            void *p = it->second.p;
            c.device.erase(it);
            c.deviceBytes[dev] -= bytes;
            return p;
        }
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        release_caches();
        e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess)
        throw Error(std::string("cudaMalloc of ") + std::to_string(bytes) +
                    " bytes failed: " + cudaGetErrorString(e));
    return p;
}

void device_give(void *p, size_t bytes, int dev)
{
    Caches &c = caches();
    {
        std::lock_guard<std::mutex> lk(c.mu);
        if (c.exported.erase(p)) {
            cudaFree(p);
            return;
        }
    }
    if (c.enabled) {
        std::lock_guard<std::mutex> lk(c.mu);
        if (!c.deviceLimit.count(dev)) {
            size_t freeB = 0, totalB = 0;
            int cur = 0;
            cudaGetDevice(&cur);
            if (cur == dev && cudaMemGetInfo(&freeB, &totalB) == cudaSuccess)
                c.deviceLimit[dev] = totalB / 2;
            else
                cudaGetLastError();
        }
        const size_t limit = c.deviceLimit.count(dev) ? c.deviceLimit[dev] : 0;
        if (c.deviceBytes[dev] + bytes <= limit) {
            c.device.emplace(std::make_pair(dev, bytes), CachedBlock{p, c.gen});
            c.deviceBytes[dev] += bytes;
            return;
        }
    }
    cudaFree(p);
}
}  // namespace

void *pinned_take(size_t bytes)
{
    Caches &c = caches();
    {
        std::lock_guard<std::mutex> lk(c.mu);
        auto it = c.pinned.find(bytes);
        if (it != c.pinned.end()) {
            void *p = it->second.p;
            c.pinned.erase(it);
            c.pinnedBytes -= bytes;
            return p;
        }
    }
    void *p = nullptr;
    SW_CUDA(cudaMallocHost(&p, bytes));
    return p;
}

void pinned_give(void *p, size_t bytes)
{
    Caches &c = caches();
    if (c.enabled) {
        std::lock_guard<std::mutex> lk(c.mu);
        if (c.pinnedBytes + bytes <= (512u << 20)) {
            c.pinned.emplace(bytes, CachedBlock{p, c.gen});
            c.pinnedBytes += bytes;
            return;
        }
    }
    cudaFreeHost(p);
}

void release_caches()
{
    Caches &c = caches();
    std::lock_guard<std::mutex> lk(c.mu);
    for (auto &kv : c.device)
        cudaFree(kv.second.p);
    c.device.clear();
    c.deviceBytes.clear();
    for (auto &kv : c.pinned)
        cudaFreeHost(kv.second.p);
    c.pinned.clear();
    c.pinnedBytes = 0;
    cudaGetLastError();
}

// A drop-in forward() brackets itself with these two: what the caches hold
// after the call is the working set of THIS call only -- blocks that came back
// during an earlier call and were not taken again by this one are freed, so a
// survey over changing shapes does not pile up dead blocks next to another
// framework in the same process, while a survey over one shape still pays
// cudaMalloc / cudaMallocHost once.
void cache_begin_call()
{
    Caches &c = caches();
    std::lock_guard<std::mutex> lk(c.mu);
    c.gen++;
}

void cache_end_call()
{
    Caches &c = caches();
    std::lock_guard<std::mutex> lk(c.mu);
    if (c.keepAll)
        return;
    for (auto it = c.device.begin(); it != c.device.end();) {
        if (it->second.gen < c.gen) {
            cudaFree(it->second.p);
            c.deviceBytes[it->first.first] -= it->first.second;
            it = c.device.erase(it);
        } else {
            ++it;
        }
    }
    for (auto it = c.pinned.begin(); it != c.pinned.end();) {
        if (it->second.gen < c.gen) {
            cudaFreeHost(it->second.p);
            c.pinnedBytes -= it->first;
            it = c.pinned.erase(it);
        } else {
            ++it;
        }
    }
    cudaGetLastError();
}

size_t cached_device_bytes()
{
    Caches &c = caches();
    std::lock_guard<std::mutex> lk(c.mu);
    size_t total = 0;
    for (auto &kv : c.deviceBytes)
        total += kv.second;
    return total;
}

void cache_mark_exported(void *base)
{
    Caches &c = caches();
    std::lock_guard<std::mutex> lk(c.mu);
    c.exported.insert(base);
}

bool is_pinned_host(const void *p)
{
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}

// Helper threads inherit the CPU mask of the thread that creates them.  An
// OpenMP runtime started with OMP_PROC_BIND pins the application's main thread
// to ONE core when it initialises (libgomp does so at `import torch`), and
// helpers that are meant to run side by side would then share that core
// (measured: staged uploads at 5 GB/s instead of 27).  A helper that finds
// itself with a one-CPU mask on a larger machine therefore asks for every CPU
// the process may use (the kernel intersects the request with the cgroup's
// cpuset).  SIMWAVE_CUDA_HELPER_AFFINITY=inherit keeps the inherited mask,
// =all widens unconditionally.
void widen_helper_affinity()
{
    const char *mode = std::getenv("SIMWAVE_CUDA_HELPER_AFFINITY");
    if (mode && std::strcmp(mode, "inherit") == 0)
        return;
    cpu_set_t cur;
    CPU_ZERO(&cur);
    if (sched_getaffinity(0, sizeof(cur), &cur) != 0)
        return;
    const bool all = mode && std::strcmp(mode, "all") == 0;
    if (!all && !(CPU_COUNT(&cur) == 1 && std::thread::hardware_concurrency() > 1))
        return;
    cpu_set_t want;
    CPU_ZERO(&want);
    for (int i = 0; i < CPU_SETSIZE; i++)
        CPU_SET(i, &want);
    sched_setaffinity(0, sizeof(want), &want);   // best effort
}

// Host-side copies between pageable arrays and the pinned staging buffers.
// One core of these hosts moves ~2-3 GB/s, the PCIe link 50: the copy is cut
// into pieces for a pool of helper threads that lives as long as the process
// (SIMWAVE_CUDA_COPY_THREADS, default half of the cores, 4..16); callers --
// the upload of a plan, its drain thread, the per-device threads of a slab
// run -- share the pool and take part in their own copy.
namespace {
class CopyPool {
public:
    struct Piece {
        char *dst;
        const char *src;
        size_t bytes;
        std::atomic<int> *left;
    };
    static CopyPool &get()
    {
        static CopyPool *p = new CopyPool;    // never destroyed: no joins at exit
        return *p;
    }
    void run(char *dst, const char *src, size_t bytes)
    {
        const size_t kMinPiece = 1u << 20;
        const size_t parts = std::min<size_t>(threads_.size() + 1, std::max<size_t>(1, bytes / kMinPiece));
        if (parts <= 1) {
            std::memcpy(dst, src, bytes);
            return;
        }
        const size_t part = (bytes / parts + 4095) & ~size_t(4095);
        std::atomic<int> left{0};
        size_t mine = std::min(part, bytes);
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (size_t b = mine; b < bytes; b += part) {
                left.fetch_add(1, std::memory_order_relaxed);
                queue_.push_back(Piece{dst + b, src + b, std::min(part, bytes - b), &left});
            }
        }
        cv_.notify_all();
        std::memcpy(dst, src, mine);
        // help with whatever is queued (mine or another caller's), then wait
        for (;;) {
            Piece pc;
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (queue_.empty())
                    break;
                pc = queue_.front();
                queue_.pop_front();
            }
            std::memcpy(pc.dst, pc.src, pc.bytes);
            pc.left->fetch_sub(1, std::memory_order_release);
        }
        while (left.load(std::memory_order_acquire) != 0)
            std::this_thread::yield();
    }

private:
    CopyPool()
    {
        unsigned hw = std::thread::hardware_concurrency();
        int n = (int)std::min(16u, std::max(4u, (hw ? hw : 8u) / 2));
        if (const char *e = std::getenv("SIMWAVE_CUDA_COPY_THREADS"))
            n = std::max(1, std::min(64, std::atoi(e)));
        for (int i = 0; i < n - 1; i++)
            threads_.emplace_back([this] { worker(); });
        for (auto &t : threads_)
            t.detach();
    }
    void worker()
    {
        widen_helper_affinity();
        for (;;) {
            Piece pc;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return !queue_.empty(); });
                pc = queue_.front();
                queue_.pop_front();
            }
            std::memcpy(pc.dst, pc.src, pc.bytes);
            pc.left->fetch_sub(1, std::memory_order_release);
        }
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Piece> queue_;
    std::vector<std::thread> threads_;
};
}  // namespace

void parallel_memcpy(void *dst, const void *src, size_t bytes)
{
    CopyPool::get().run((char *)dst, (const char *)src, bytes);
}

void DeviceBuffer::alloc(size_t bytes)
{
    release();
    ptr_ = device_take(bytes ? bytes : 1, &device_);
    bytes_ = bytes;
}

void DeviceBuffer::release()
{
    if (ptr_)
        device_give(ptr_, bytes_ ? bytes_ : 1, device_);
    ptr_ = nullptr;
    bytes_ = 0;
}

// ---------------------------------------------------------------------------
// HostDrain
// ---------------------------------------------------------------------------
HostDrain::HostDrain(int device, size_t chunkBytes) : device_(device), chunkBytes_(chunkBytes)
{
    SW_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++)
        SW_CUDA(cudaEventCreateWithFlags(&copied_[i], cudaEventDisableTiming));
    thread_ = std::thread([this] { worker(); });
}

HostDrain::~HostDrain()
{
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_.notify_all();
    if (thread_.joinable())
        thread_.join();
    for (int i = 0; i < 2; i++) {
        if (pinned_[i]) pinned_give(pinned_[i], chunkBytes_);
        if (copied_[i]) cudaEventDestroy(copied_[i]);
    }
    if (stream_) cudaStreamDestroy(stream_);
}

void HostDrain::submit(Job job)
{
    {
        std::lock_guard<std::mutex> lk(mu_);
        queue_.push_back(std::move(job));
    }
    cv_.notify_all();
}

void HostDrain::wait_idle()
{
    std::unique_lock<std::mutex> lk(mu_);
    idle_.wait(lk, [this] { return queue_.empty() && !busy_; });
    if (!error_.empty()) {
        std::string e = error_;
        error_.clear();
        throw Error("snapshot drain failed: " + e);
    }
}

void HostDrain::ensure_staging()
{
    for (int i = 0; i < 2; i++)
        if (!pinned_[i])
            pinned_[i] = pinned_take(chunkBytes_);
}

void HostDrain::worker()
{
    widen_helper_affinity();
    cudaSetDevice(device_);
    for (;;) {
        Job job;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [this] { return stop_ || !queue_.empty(); });
            if (queue_.empty())
                return;  // stop requested and nothing left
            job = std::move(queue_.front());
            queue_.pop_front();
            busy_ = true;
        }
        try {
            SW_CUDA(cudaEventSynchronize(job.ready));
            if (is_pinned_host(job.dst)) {
                // page-locked destination: one strided DMA, no staging
                SW_CUDA(cudaMemcpy2DAsync(job.dst, job.rowBytes, job.src, job.srcPitchBytes,
                                          job.rowBytes, job.rows, cudaMemcpyDeviceToHost,
                                          stream_));
                SW_CUDA(cudaStreamSynchronize(stream_));
                job.rows = 0;
            } else {
                ensure_staging();
            }
            const size_t rowsPerChunk = std::max<size_t>(1, chunkBytes_ / job.rowBytes);
            const size_t chunks = (job.rows + rowsPerChunk - 1) / rowsPerChunk;
            auto issue = [&](size_t c) {
                const size_t r0 = c * rowsPerChunk;
                const size_t nr = std::min(rowsPerChunk, job.rows - r0);
                SW_CUDA(cudaMemcpy2DAsync(pinned_[c & 1], job.rowBytes,
                                          (const char *)job.src + r0 * job.srcPitchBytes,
                                          job.srcPitchBytes, job.rowBytes, nr,
                                          cudaMemcpyDeviceToHost, stream_));
                SW_CUDA(cudaEventRecord(copied_[c & 1], stream_));
            };
            if (chunks)
                issue(0);
            double tWait = 0, tCopy = 0;
            for (size_t c = 0; c < chunks; c++) {
                const double a0 = wall();
                SW_CUDA(cudaEventSynchronize(copied_[c & 1]));
                if (c + 1 < chunks)
                    issue(c + 1);  // overlaps with the memcpy below
                const double a1 = wall();
                const size_t r0 = c * rowsPerChunk;
                const size_t nr = std::min(rowsPerChunk, job.rows - r0);
                parallel_memcpy((char *)job.dst + r0 * job.rowBytes, pinned_[c & 1],
                                nr * job.rowBytes);
                tWait += a1 - a0;
                tCopy += wall() - a1;
            }
            if (chunks && std::getenv("SIMWAVE_CUDA_VERBOSE"))
                std::fprintf(stderr,
                             "simwave_b200: drain of %.0f MB through staging: %.1f ms waiting for "
                             "the DMA, %.1f ms copying to the caller's pages\n",
                             job.rows * job.rowBytes / 1e6, 1e3 * tWait, 1e3 * tCopy);
        } catch (const std::exception &e) {
            std::lock_guard<std::mutex> lk(mu_);
            if (error_.empty())
                error_ = e.what();
        }
        if (job.ready)
            cudaEventDestroy(job.ready);
        if (job.done)
            job.done();
        {
            std::lock_guard<std::mutex> lk(mu_);
            busy_ = false;
        }
        idle_.notify_all();
    }
}

// ---------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------

// true if [p, p+bytes) is all zero bits; scanned by a few threads, stops early
static bool all_zero(const void *p, size_t bytes)
{
    const size_t words = bytes / 8;
    const uint64_t *w = (const uint64_t *)p;
    unsigned hw = std::thread::hardware_concurrency();
    const int nthreads = (bytes > (64u << 20)) ? (int)std::min(16u, std::max(3u, hw) - 1) : 1;
    std::vector<char> nz(nthreads, 0);
    auto scan = [&](int t) {
        if (nthreads > 1)
            widen_helper_affinity();
        const size_t b = words * t / nthreads, e = words * (t + 1) / nthreads;
        const size_t blk = 4096;
        for (size_t i = b; i < e; i += blk) {
            uint64_t acc = 0;
            const size_t m = std::min(e, i + blk);
            for (size_t j = i; j < m; j++)
                acc |= w[j];
            if (acc) { nz[t] = 1; return; }
        }
    };
    if (nthreads == 1) {
        scan(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) th.emplace_back(scan, t);
        for (auto &x : th) x.join();
    }
    for (char c : nz)
        if (c) return false;
    const unsigned char *tail = (const unsigned char *)p + words * 8;
    for (size_t i = 0; i < bytes - words * 8; i++)
        if (tail[i]) return false;
    return true;
}

static dim3 row_grid(const Grid &g)
{
    const long long rows = (long long)g.nS * g.nM;
    return dim3((g.nF + 255) / 256, (unsigned)std::min<long long>(rows, 65535), 1);
}

// ---------------------------------------------------------------------------
// tiled 3D kernel plumbing
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        SW_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess || !p)
            throw Error("driver does not provide cuTensorMapEncodeTiled");
        fn = (EncodeTiledFn)p;
    }
    return fn;
}

typedef bool (*TiledQueryFn)(int, bool, int, TiledInfo *);
typedef bool (*TiledLaunchFn)(int, bool, int, const StepArgs<float> &, const StepMaps &,
                              const unsigned char *, int, cudaStream_t);
static const TiledQueryFn kTiledQuery[kMaxRadius + 1] = {
    nullptr, tiled3d_query_r1, tiled3d_query_r2, tiled3d_query_r3, tiled3d_query_r4,
    tiled3d_query_r5, tiled3d_query_r6, tiled3d_query_r7, tiled3d_query_r8, tiled3d_query_r9,
    tiled3d_query_r10};
static const TiledLaunchFn kTiledLaunch[kMaxRadius + 1] = {
    nullptr, tiled3d_launch_r1, tiled3d_launch_r2, tiled3d_launch_r3, tiled3d_launch_r4,
    tiled3d_launch_r5, tiled3d_launch_r6, tiled3d_launch_r7, tiled3d_launch_r8,
    tiled3d_launch_r9, tiled3d_launch_r10};

typedef bool (*Tiled64QueryFn)(bool, int, TiledInfo *);
typedef bool (*Tiled64LaunchFn)(bool, int, const StepArgs<double> &, const StepMaps &,
                                const unsigned char *, int, cudaStream_t);
static const Tiled64QueryFn kTiled64Query[kMaxRadius + 1] = {
    nullptr, tiled3d64_query_r1, tiled3d64_query_r2, tiled3d64_query_r3, tiled3d64_query_r4,
    tiled3d64_query_r5, tiled3d64_query_r6, tiled3d64_query_r7, tiled3d64_query_r8,
    tiled3d64_query_r9, tiled3d64_query_r10};
static const Tiled64LaunchFn kTiled64Launch[kMaxRadius + 1] = {
    nullptr, tiled3d64_launch_r1, tiled3d64_launch_r2, tiled3d64_launch_r3, tiled3d64_launch_r4,
    tiled3d64_launch_r5, tiled3d64_launch_r6, tiled3d64_launch_r7, tiled3d64_launch_r8,
    tiled3d64_launch_r9, tiled3d64_launch_r10};

// ---------------------------------------------------------------------------
// slab decomposition plumbing
// ---------------------------------------------------------------------------
struct SlabDesc {
    unsigned magic;
    int dtypeBytes, nS, nM, nF, r, lpad, device;
    long long pitch, planeStride;
    unsigned long long slotOffset[3];   // field base (element (0,0,0)) from the mapped base
    unsigned long long flagsOffset;
    cudaIpcMemHandle_t slot[3];
    cudaIpcMemHandle_t flags;
};
static_assert(sizeof(SlabDesc) <= SIMWAVE_SLAB_DESC_BYTES, "descriptor too large");
constexpr unsigned kSlabMagic = 0x534c4142u;   // "SLAB"

// flags[0]: last step whose halo the UP neighbour has delivered, [1]: same for
// DOWN, [2]: error word (1 = timed out)
__global__ void slab_wait_kernel(int *flags, int needUp, int needDown, int value,
                                 long long timeoutCycles)
{
    const long long start = clock64();
    for (int side = 0; side < 2; side++) {
        if (!(side == 0 ? needUp : needDown))
            continue;
        for (;;) {
            int v;
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + side)
                         : "memory");
            if (v >= value)
                break;
            if (clock64() - start > timeoutCycles) {   // neighbour is gone
                flags[2] = 1;
                return;
            }
            __nanosleep(200);
        }
    }
}

__global__ void slab_publish_kernel(int *peerFlag, int value)
{
    __threadfence_system();
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(peerFlag), "r"(value) : "memory");
}

// handle + offset of a device pointer inside its cudaMalloc allocation
static void ipc_export(const void *ptr, cudaIpcMemHandle_t *handle, unsigned long long *offset)
{
    typedef CUresult (*RangeFn)(CUdeviceptr *, size_t *, CUdeviceptr);
    static RangeFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        SW_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q));
        if (q != cudaDriverEntryPointSuccess || !p)
            throw Error("driver does not provide cuMemGetAddressRange");
        fn = (RangeFn)p;
    }
    CUdeviceptr base = 0;
    size_t size = 0;
    if (fn(&base, &size, (CUdeviceptr)ptr) != CUDA_SUCCESS)
        throw Error("cuMemGetAddressRange failed");
    SW_CUDA(cudaIpcGetMemHandle(handle, (void *)base));
    cache_mark_exported((void *)base);
    *offset = (unsigned long long)((CUdeviceptr)ptr - base);
}

// ---------------------------------------------------------------------------
// Model fields of a problem on the device: everything that depends only on
// velocity / damping / density, the grid and dt.  A plan normally owns its
// own; under SIMWAVE_HINT_MODEL_RESIDENT (the caller vouches that the model
// arrays behind a token do not change) the set of the last forward() on a
// device is kept and handed to the next plan with the same key, so a survey
// uploads and preprocesses the model once instead of once per shot.
// ---------------------------------------------------------------------------
struct ModelKey {
    long long token = 0;
    int device = -1, ndim = 0, dtypeBytes = 0, order = 0, math = 0, varden = 0;
    size_t n[3] = {0, 0, 0};
    double h[3] = {0, 0, 0}, dt = 0;
    bool operator==(const ModelKey &o) const
    {
        return token == o.token && device == o.device && ndim == o.ndim &&
               dtypeBytes == o.dtypeBytes && order == o.order && math == o.math &&
               varden == o.varden && n[0] == o.n[0] && n[1] == o.n[1] && n[2] == o.n[2] &&
               h[0] == o.h[0] && h[1] == o.h[1] && h[2] == o.h[2] && dt == o.dt;
    }
};
struct ModelFields {
    ModelKey key;
    DeviceBuffer c0, q, rho;
    bool built = false;
    // tiled 3D kernel extras
    DeviceBuffer qflags;            // [nS][tilesM][tilesF]: damping profile non-zero in the tile?
    DeviceBuffer frF, frM, frS;     // first derivatives of the density
    int tileM = 0, tileF = 0;       // tile the flags were built for (0: not built)
};
namespace {
std::mutex g_residentMu;
std::map<int, std::shared_ptr<ModelFields>> g_resident;   // per device ordinal
}
void drop_resident_models()
{
    std::lock_guard<std::mutex> lk(g_residentMu);
    g_resident.clear();
}

// ---------------------------------------------------------------------------
// Plan
// ---------------------------------------------------------------------------
template <typename T>
class Plan : public PlanBase {
public:
    Plan(const simwave_problem &pb, const Options &opt);
    ~Plan() override;
    void run(size_t begin, size_t end) override;
    void download(void *u, void *receivers) override;
    void prefault_outputs(size_t end) override;
    void reset() override;
    void slab_export(void *desc) override;
    void slab_connect(const void *up, const void *down) override;
    void slab_peer(SlabPeer *out) override;
    void slab_connect_direct(const SlabPeer *up, const SlabPeer *down) override;

private:
    using StepFn = void (*)(int, const StepArgs<T> &, cudaStream_t);

    void check_launch(const char *what);
    T *field_base(const DeviceBuffer &b) const { return b.as<T>() + guard_ + g_.lpad; }
    // F halo of a TMA box in elements: the radius rounded up to 16 bytes
    static int halo_f(int r) { return sizeof(T) == 4 ? (r + 3) / 4 * 4 : (r + 1) / 2 * 2; }
    void new_field(DeviceBuffer &b);
    T *acquire();
    void give_back(T *buf);
    void ensure_live(size_t slot);
    void retire(size_t slot, bool keep);
    void retire_below(size_t bound);
    void upload_dense(const T *host, T *pitchedBase);
    void h2d(void *dst, const void *src, size_t bytes);
    void launch_step(const StepArgs<T> &a);
    void launch_sources(const StepArgs<T> &a, size_t n);
    void launch_receivers(const T *cur, size_t n);
    void launch_boundaries(T *next);
    bool run_persistent(size_t begin, size_t end);

    Options opt_;
    int device_ = 0;
    int ndim_;
    bool varden_;
    Grid g_;
    size_t guard_;          // elements before/after each field allocation
    size_t denseCells_;
    size_t hostSlotStride_;            // elements between slots of the caller's u
    size_t outPlaneLo_ = 0, outPlaneHi_ = 0;   // local planes download() writes back
    size_t fieldBytes_;

    T *hostU_;
    T *hostRec_;
    size_t numSlots_, stride_, waveletSize_, waveletCount_, nsrc_, nrec_;
    std::vector<char> slotZero_;

    std::shared_ptr<ModelFields> model_;
    DeviceBuffer stage_;
    DeviceBuffer wavelet_, srcIv_, srcVal_, srcOff_, recIv_, recVal_, recOff_, recOut_;
    StepArgs<T> args_;
    PointTables<T> srcTab_, recTab_;
    int srcMode_ = SRC_DISJOINT;
    int srcMaxPoints_ = 1;
    bool srcInterior_ = false;          // every source window lies among the interior points
    int srcBox_[3][2] = {{0, 0}, {0, 0}, {0, 0}};   // bounding box of the windows, (S,M,F)
    StepFn stepSimple_ = nullptr;

    // tiled 3D kernel (float32, constant density)
    bool useTiled_ = false;
    bool srcFusedTiled_ = false;   // sources added inside the tiled step kernel
    size_t stepNow_ = 0;           // time step being launched
    int tiledCfg_ = 0;
    int zChunk_ = 0;
    TiledInfo tiledInfo_{};
    std::map<std::pair<const void *, bool>, CUtensorMap> maps_;
    const CUtensorMap &field_map(const T *base, bool halo);
    void choose_tiling();
    DeviceBuffer loopBarrier_;   // grid barrier word of the persistent 2D loop
    // tile-resident 2D loop: origin of every receiver window (host copy), the
    // tiling once chosen, receivers grouped by owner tile, per-tile step flags
    std::vector<int> recLoM_, recLoF_;
    int recMaxM_ = 1, recMaxF_ = 1;
    int residentState_ = 0;      // 0 = not tried yet, 1 = available, -1 = not available
    Loop2dTiling residentTiling_{};
    DeviceBuffer residentRecStart_, residentRecIndex_, residentFlags_;
    bool run_resident(LoopArgs<T> &L);
    std::thread prefault_;       // see PlanBase::prefault_outputs
    std::atomic<bool> prefaultStop_{false};
    void join_prefault()
    {
        if (prefault_.joinable()) {
            prefault_.join();
        }
    }

    cudaStream_t stream_ = nullptr;
    cudaEvent_t evBegin_ = nullptr, evEnd_ = nullptr;
    cudaStream_t recStream_ = nullptr;     // receiver kernels (nullptr: on stream_)
    cudaEvent_t recReady_ = nullptr, recDone_ = nullptr;
    bool recPending_ = false;

    // slots
    std::map<size_t, T *> live_;
    std::map<size_t, bool> dirty_;
    std::vector<std::unique_ptr<DeviceBuffer>> buffers_;
    std::vector<T *> free_;
    std::mutex poolMu_;
    std::condition_variable poolCv_;
    size_t maxBuffers_ = 6;
    size_t prevT_ = 0, curT_ = 1, nextT_ = 2;
    size_t recBegin_ = 0, recEnd_ = 0;   // receiver rows produced so far [begin,end)

    std::unique_ptr<HostDrain> drain_;

    // pinned staging for uploads from pageable memory
    void *upPinned_[2] = {nullptr, nullptr};
    cudaEvent_t upDone_[2] = {nullptr, nullptr};
    static constexpr size_t kUpChunk = 32u << 20;

    // slab decomposition: neighbours' slot buffers and flag words, mapped
    // through CUDA IPC; flags_ = {from up, from down, error}
    bool slabUp_ = false, slabDown_ = false, slabConnected_ = false;
    DeviceBuffer slabFlags_;
    T *peerSlot_[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    int *peerFlags_[2] = {nullptr, nullptr};
    int peerNS_[2] = {0, 0};
    std::vector<void *> ipcMapped_;
    bool slabFused_ = false;   // ghost copies written by the step / source kernels themselves
    size_t ranLo_ = 0, ranHi_ = 0;   // contiguous range of steps run since the last reset (0: none)
    void slab_wait(size_t n);
    void slab_push(size_t n, size_t slot);
};

template <typename T>
void Plan<T>::check_launch(const char *what)
{
    timing.launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && opt_.debug)
        e = cudaStreamSynchronize(stream_);
    if (e != cudaSuccess)
        throw Error(std::string(what) + ": " + cudaGetErrorString(e));
}

template <typename T>
Plan<T>::Plan(const simwave_problem &pb, const Options &opt) : opt_(opt)
{
    const double t0 = wall();
    NvtxRange nvtx("simwave_b200: upload");
    PhaseLog phases;
    ndim_ = pb.ndim;
    varden_ = pb.density != nullptr;
    if (ndim_ != 2 && ndim_ != 3)
        throw Error("ndim must be 2 or 3");
    if (pb.space_order < 2 || pb.space_order > 2 * kMaxRadius || pb.space_order % 2)
        throw Error("space_order must be even and between 2 and 20");
    if (varden_ && !pb.coeff_order1)
        throw Error("variable density needs coeff_order1");

    if (opt_.device >= 0)
        SW_CUDA(cudaSetDevice(opt_.device));
    SW_CUDA(cudaGetDevice(&device_));

    const int r = (int)(pb.space_order / 2);
    Grid &g = g_;
    g.ndim = ndim_;
    if (ndim_ == 3) {
        g.nS = (int)pb.nz; g.nM = (int)pb.nx; g.nF = (int)pb.ny;
    } else {
        g.nS = 1; g.nM = (int)pb.nz; g.nF = (int)pb.nx;
    }
    g.r = r;
    if (g.nF < 2 * r + 1 || g.nM < 2 * r + 1 || (ndim_ == 3 && g.nS < 2 * r + 1))
        throw Error("grid smaller than the stencil halo");
    const int align = 128 / (int)sizeof(T);
    g.lpad = (align - r % align) % align;
    g.pitch = ((long long)g.lpad + g.nF + align - 1) / align * align;
    g.planeStride = (long long)g.nM * g.pitch;
    g.cells = (long long)g.nS * g.planeStride;
    guard_ = 2 * align;
    denseCells_ = (size_t)g.nS * g.nM * g.nF;
    hostSlotStride_ = pb.u_slot_stride ? pb.u_slot_stride : denseCells_;
    outPlaneLo_ = 0;
    outPlaneHi_ = (size_t)g.nS;
    if (pb.out_plane_end > pb.out_plane_begin) {
        if (pb.out_plane_end > (size_t)g.nS)
            throw Error("out_plane range outside the slab");
        outPlaneLo_ = pb.out_plane_begin;
        outPlaneHi_ = pb.out_plane_end;
    }
    fieldBytes_ = (g.cells + 2 * guard_) * sizeof(T);

    hostU_ = (T *)pb.u;
    hostRec_ = (T *)pb.receivers;
    numSlots_ = pb.num_snapshots;
    stride_ = pb.saving_stride;
    waveletSize_ = pb.wavelet_size;
    waveletCount_ = pb.wavelet_count ? pb.wavelet_count : 1;
    nsrc_ = pb.num_sources;
    nrec_ = pb.num_receivers;
    if (numSlots_ < 3)
        throw Error("u must hold at least 3 slots");
    // Which of the caller's slots start as zeros (those are never uploaded):
    // scanned on helper threads while the model goes up.  A page-locked `u`
    // with only the three rotating slots is cheaper to upload outright (one
    // DMA at link speed) than to read once with the CPU.
    slotZero_.assign(numSlots_, opt_.zeroIn ? 1 : 0);
    std::future<void> zeroScan;
    if (!opt_.zeroIn && !(numSlots_ == 3 && is_pinned_host(hostU_)))
        zeroScan = std::async(std::launch::async, [this] {
            if (hostSlotStride_ == denseCells_ &&
                all_zero(hostU_, numSlots_ * denseCells_ * sizeof(T))) {
                std::fill(slotZero_.begin(), slotZero_.end(), 1);
            } else {
                for (size_t s = 0; s < numSlots_; s++)
                    slotZero_[s] = all_zero(hostU_ + s * hostSlotStride_, denseCells_ * sizeof(T));
            }
        });
    slabUp_ = pb.slab_up != 0;
    slabDown_ = pb.slab_down != 0;
    if ((slabUp_ || slabDown_) && (ndim_ != 3 || stride_ != 0))
        throw Error("slab decomposition needs a 3D problem with saving_stride == 0");
    if ((slabUp_ || slabDown_) && pb.nz < 4 * (pb.space_order / 2))
        throw Error("a slab must own at least space_order planes");

    SW_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    SW_CUDA(cudaEventCreate(&evBegin_));
    SW_CUDA(cudaEventCreate(&evEnd_));
    if (pb.num_receivers && !opt_.debug && !env_is("SIMWAVE_CUDA_RECEIVERS", "inline")) {
        SW_CUDA(cudaStreamCreateWithFlags(&recStream_, cudaStreamNonBlocking));
        SW_CUDA(cudaEventCreateWithFlags(&recReady_, cudaEventDisableTiming));
        SW_CUDA(cudaEventCreateWithFlags(&recDone_, cudaEventDisableTiming));
    }
    phases.mark("context+stream");

    // ---- static part of the step arguments ------------------------------
    StepArgs<T> &a = args_;
    std::memset(&a, 0, sizeof(a));
    a.g = g;
    const T *c2 = (const T *)pb.coeff_order2;
    const T *c1 = (const T *)pb.coeff_order1;
    for (int i = 0; i <= r; i++) {
        a.c2[i] = c2[i];
        a.c1[i] = c1 ? c1[i] : T(0);
    }
    for (int k = -kSplitMid; k <= kSplitMid; k++) {
        auto coef = [&](const T *c, int off, bool odd_symmetry) {
            const int m = off < 0 ? -off : off;
            if (m > r)
                return T(0);
            return (odd_symmetry && off < 0) ? -c[m] : c[m];
        };
        a.c2odd[kSplitMid + k][0] = coef(a.c2, 2 * k - 1, false);
        a.c2odd[kSplitMid + k][1] = coef(a.c2, 2 * k + 1, false);
        a.c1odd[kSplitMid + k][0] = coef(a.c1, 2 * k - 1, true);
        a.c1odd[kSplitMid + k][1] = coef(a.c1, 2 * k + 1, true);
    }
    // spacing per axis in (S,M,F) order; squares rounded in T like the
    // reference's `f_type dzSquared = dz * dz`
    T h[3];
    if (ndim_ == 3) { h[0] = (T)pb.dz; h[1] = (T)pb.dx; h[2] = (T)pb.dy; }
    else            { h[0] = T(1);     h[1] = (T)pb.dz; h[2] = (T)pb.dx; }
    for (int i = 0; i < 3; i++) {
        volatile T sq = h[i] * h[i];
        a.h2[i] = sq;
        a.inv_h2[i] = T(1) / a.h2[i];
        a.inv_h2_lo[i] = std::fma(-a.inv_h2[i], a.h2[i], T(1)) / a.h2[i];
        volatile T f4 = T(4) * a.h2[i];
        a.four_h2[i] = f4;
        a.inv_four_h2[i] = T(1) / a.four_h2[i];
    }
    const size_t *bc = pb.boundary_conditions;
    if (ndim_ == 3) {
        for (int i = 0; i < 6; i++) a.bc[i] = (int)bc[i];
    } else {
        a.bc[0] = a.bc[1] = 0;
        for (int i = 0; i < 4; i++) a.bc[2 + i] = (int)bc[i];
    }
    for (int i = 0; i < 6; i++)
        if (a.bc[i] < 0 || a.bc[i] > 2)
            throw Error("boundary condition codes must be 0, 1 or 2");
    a.quirk = (ndim_ == 3 && varden_ && pb.nx != pb.ny) ? 1 : 0;
    a.denseNx = (int)pb.nx;
    a.denseNy = (int)(ndim_ == 3 ? pb.ny : 1);
    const int need = 3 * r + 2;
    const bool roomy = g.nF >= need && g.nM >= need && (ndim_ == 2 || g.nS >= need);
    a.fuse_bc = (roomy && !opt_.separateBc) ? 1 : 0;

    // ---- model: c0, q (and density) in pitched layout ----------------------
    const T dt = (T)pb.dt;
    volatile T dtsqv = dt * dt;   // rounded in T, like `f_type dtSquared = dt * dt`
    const T dtsq = dtsqv;

    // resident model of an earlier forward() with the same key, or a fresh one
    {
        ModelKey key;
        key.token = opt_.modelToken;
        key.device = device_; key.ndim = ndim_; key.dtypeBytes = (int)sizeof(T);
        key.order = (int)pb.space_order; key.math = opt_.math; key.varden = varden_ ? 1 : 0;
        key.n[0] = pb.nz; key.n[1] = pb.nx; key.n[2] = ndim_ == 3 ? pb.ny : 1;
        key.h[0] = pb.dz; key.h[1] = pb.dx; key.h[2] = ndim_ == 3 ? pb.dy : 0;
        key.dt = pb.dt;
        if (opt_.modelToken != 0) {
            std::lock_guard<std::mutex> lk(g_residentMu);
            auto it = g_resident.find(device_);
            if (it != g_resident.end() && it->second->key == key && it->second.use_count() == 1)
                model_ = it->second;
            else {
                model_ = std::make_shared<ModelFields>();
                model_->key = key;
                g_resident[device_] = model_;   // replaces (and frees) any other model
            }
        } else {
            model_ = std::make_shared<ModelFields>();
            model_->key = key;
        }
    }
    if (!model_->built) {
        new_field(model_->c0);
        new_field(model_->q);
        stage_.alloc(2 * denseCells_ * sizeof(T));
        T *stageA = stage_.as<T>(), *stageB = stage_.as<T>() + denseCells_;
        h2d(stageA, pb.velocity, denseCells_ * sizeof(T));
        h2d(stageB, pb.damp, denseCells_ * sizeof(T));
        model_kernel<T><<<row_grid(g), 256, 0, stream_>>>(g, stageA, stageB, dt, dtsq,
                                                          field_base(model_->c0),
                                                          field_base(model_->q));
        check_launch("model_kernel");
        if (varden_) {
            new_field(model_->rho);
            upload_dense((const T *)pb.density, field_base(model_->rho));
        }
        model_->built = true;
    }
    a.c0 = field_base(model_->c0);
    a.q = field_base(model_->q);
    a.rho = varden_ ? field_base(model_->rho) : nullptr;
    phases.mark("model alloc+enqueue");

    // ---- wavelet and tables -------------------------------------------------
    auto to_device = [&](DeviceBuffer &b, const void *src, size_t bytes) {
        b.alloc(bytes);
        if (bytes)
            SW_CUDA(cudaMemcpyAsync(b.get(), src, bytes, cudaMemcpyHostToDevice, stream_));
    };
    to_device(wavelet_, pb.wavelet, waveletSize_ * waveletCount_ * sizeof(T));
    to_device(srcIv_, pb.src_points_interval, nsrc_ * 2 * ndim_ * sizeof(size_t));
    to_device(srcVal_, pb.src_points_values, pb.src_points_values_size * sizeof(T));
    // num_sources / num_receivers entries: all the reference ABI ever reads
    to_device(srcOff_, pb.src_points_values_offset, nsrc_ * sizeof(size_t));
    to_device(recIv_, pb.rec_points_interval, nrec_ * 2 * ndim_ * sizeof(size_t));
    to_device(recVal_, pb.rec_points_values, pb.rec_points_values_size * sizeof(T));
    to_device(recOff_, pb.rec_points_values_offset, nrec_ * sizeof(size_t));
    recOut_.alloc(std::max<size_t>(1, waveletSize_ * nrec_) * sizeof(T));
    SW_CUDA(cudaMemsetAsync(recOut_.get(), 0, recOut_.bytes(), stream_));
    srcTab_ = {srcIv_.as<unsigned long long>(), srcVal_.as<T>(),
               srcOff_.as<unsigned long long>(), (int)nsrc_};
    recTab_ = {recIv_.as<unsigned long long>(), recVal_.as<T>(),
               recOff_.as<unsigned long long>(), (int)nrec_};

    // validate windows and classify source overlap
    auto check_windows = [&](const size_t *iv, size_t count, const char *what) {
        const int ext[3] = {g.nS, g.nM, g.nF};
        for (size_t i = 0; i < count; i++)
            for (int ax = 0; ax < ndim_; ax++) {
                const size_t lo = iv[(i * ndim_ + ax) * 2], hi = iv[(i * ndim_ + ax) * 2 + 1];
                if (lo > hi || hi >= (size_t)ext[ax + 3 - ndim_])
                    throw Error(std::string(what) + " window outside the grid");
                if (ax == ndim_ - 1 && hi - lo + 1 > 32)
                    throw Error(std::string(what) + " window wider than 32 points");
            }
    };
    check_windows(pb.src_points_interval, nsrc_, "source");
    check_windows(pb.rec_points_interval, nrec_, "receiver");
    if (ndim_ == 2) {
        const size_t *iv = pb.rec_points_interval;
        recLoM_.resize(nrec_);
        recLoF_.resize(nrec_);
        for (size_t i = 0; i < nrec_; i++) {
            recLoM_[i] = (int)iv[i * 4];
            recLoF_[i] = (int)iv[i * 4 + 2];
            recMaxM_ = std::max(recMaxM_, (int)(iv[i * 4 + 1] - iv[i * 4] + 1));
            recMaxF_ = std::max(recMaxF_, (int)(iv[i * 4 + 3] - iv[i * 4 + 2] + 1));
        }
    }
    {
        const size_t *iv = pb.src_points_interval;
        bool overlap = false;
        srcMaxPoints_ = 1;
        for (size_t i = 0; i < nsrc_; i++) {
            int pts = 1;
            for (int ax = 0; ax < ndim_; ax++)
                pts *= (int)(iv[(i * ndim_ + ax) * 2 + 1] - iv[(i * ndim_ + ax) * 2] + 1);
            srcMaxPoints_ = std::max(srcMaxPoints_, pts);
        }
        if (nsrc_ > 4096) {
            overlap = true;
        } else {
            for (size_t i = 0; i < nsrc_ && !overlap; i++)
                for (size_t j = i + 1; j < nsrc_ && !overlap; j++) {
                    bool hit = true;
                    for (int ax = 0; ax < ndim_; ax++) {
                        const size_t li = iv[(i * ndim_ + ax) * 2], hi_ = iv[(i * ndim_ + ax) * 2 + 1];
                        const size_t lj = iv[(j * ndim_ + ax) * 2], hj = iv[(j * ndim_ + ax) * 2 + 1];
                        hit = hit && li <= hj && lj <= hi_;
                    }
                    overlap = hit;
                }
        }
        srcMode_ = !overlap ? SRC_DISJOINT : (nsrc_ <= 64 ? SRC_ORDERED : SRC_ATOMIC);
        const int ext[3] = {g.nS, g.nM, g.nF};
        srcInterior_ = nsrc_ > 0;
        for (int ax = 0; ax < 3; ax++) { srcBox_[ax][0] = 1 << 30; srcBox_[ax][1] = -1; }
        for (size_t i = 0; i < nsrc_; i++)
            for (int ax = 0; ax < ndim_; ax++) {
                const int lo = (int)iv[(i * ndim_ + ax) * 2], hi = (int)iv[(i * ndim_ + ax) * 2 + 1];
                const int axis = ax + 3 - ndim_;
                srcBox_[axis][0] = std::min(srcBox_[axis][0], lo);
                srcBox_[axis][1] = std::max(srcBox_[axis][1], hi);
                if (lo < r || hi > ext[axis] - r - 1)
                    srcInterior_ = false;
            }
    }

    // ---- kernels ---------------------------------------------------------------
    if (ndim_ == 3)
        stepSimple_ = varden_ ? &launch_step_simple<T, 3, true> : &launch_step_simple<T, 3, false>;
    else
        stepSimple_ = varden_ ? &launch_step_simple<T, 2, true> : &launch_step_simple<T, 2, false>;

    phases.mark("tables");
    choose_tiling();
    phases.mark("tiling");

    if (zeroScan.valid())
        zeroScan.get();
    phases.mark("zero scan of u");
    drain_.reset(new HostDrain(device_, 32u << 20));
    phases.mark("drain thread+pinned");
    if (slabUp_ || slabDown_) {
        // the three slots stay resident so that the neighbours can map them
        for (size_t s = 0; s < 3; s++)
            ensure_live(s);
        slabFlags_.alloc(4 * sizeof(int));
        SW_CUDA(cudaMemsetAsync(slabFlags_.get(), 0, slabFlags_.bytes(), stream_));
    }
    SW_CUDA(cudaStreamSynchronize(stream_));
    phases.mark("sync");
    timing.h2d = wall() - t0;
}

template <typename T>
Plan<T>::~Plan()
{
    prefaultStop_.store(true);
    join_prefault();
    drain_.reset();   // joins the worker before buffers go away
    if (recStream_) {
        cudaStreamSynchronize(recStream_);
        cudaStreamDestroy(recStream_);
        cudaEventDestroy(recReady_);
        cudaEventDestroy(recDone_);
    }
    if (stream_) {
        cudaStreamSynchronize(stream_);
        cudaStreamDestroy(stream_);
    }
    for (int i = 0; i < 2; i++) {
        if (upPinned_[i]) pinned_give(upPinned_[i], kUpChunk);
        if (upDone_[i]) cudaEventDestroy(upDone_[i]);
    }
    for (void *p : ipcMapped_)
        cudaIpcCloseMemHandle(p);
    if (evBegin_) cudaEventDestroy(evBegin_);
    if (evEnd_) cudaEventDestroy(evEnd_);
}

template <typename T>
void Plan<T>::new_field(DeviceBuffer &b)
{
    b.alloc(fieldBytes_);
    SW_CUDA(cudaMemsetAsync(b.get(), 0, fieldBytes_, stream_));
}

// Host -> device copy on the compute stream.  Page-locked sources go as one
// DMA; pageable ones are staged through two pinned buffers filled by a few
// threads, so the copy runs at link speed instead of the driver's
// single-threaded pageable path.
template <typename T>
void Plan<T>::h2d(void *dst, const void *src, size_t bytes)
{
    if (bytes < (8u << 20) || is_pinned_host(src)) {
        SW_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream_));
        return;
    }
    for (int i = 0; i < 2; i++) {
        if (!upPinned_[i]) {
            upPinned_[i] = pinned_take(kUpChunk);
            SW_CUDA(cudaEventCreateWithFlags(&upDone_[i], cudaEventDisableTiming));
        }
    }
    size_t c = 0;
    for (size_t off = 0; off < bytes; off += kUpChunk, c++) {
        const size_t n = std::min(kUpChunk, bytes - off);
        SW_CUDA(cudaEventSynchronize(upDone_[c & 1]));   // staging buffer free again
        parallel_memcpy(upPinned_[c & 1], (const char *)src + off, n);
        SW_CUDA(cudaMemcpyAsync((char *)dst + off, upPinned_[c & 1], n, cudaMemcpyHostToDevice,
                                stream_));
        SW_CUDA(cudaEventRecord(upDone_[c & 1], stream_));
    }
}

template <typename T>
void Plan<T>::upload_dense(const T *host, T *pitchedBase)
{
    if (!stage_.get())
        stage_.alloc(2 * denseCells_ * sizeof(T));
    h2d(stage_.get(), host, denseCells_ * sizeof(T));
    pack_kernel<T><<<row_grid(g_), 256, 0, stream_>>>(g_, stage_.as<T>(), pitchedBase);
    check_launch("pack_kernel");
}

template <typename T>
T *Plan<T>::acquire()
{
    std::unique_lock<std::mutex> lk(poolMu_);
    if (free_.empty() && buffers_.size() < maxBuffers_) {
        lk.unlock();
        std::unique_ptr<DeviceBuffer> b(new DeviceBuffer());
        new_field(*b);
        T *base = field_base(*b);
        buffers_.push_back(std::move(b));
        return base;
    }
    poolCv_.wait(lk, [this] { return !free_.empty(); });
    T *base = free_.back();
    free_.pop_back();
    return base;
}

template <typename T>
void Plan<T>::give_back(T *buf)
{
    {
        std::lock_guard<std::mutex> lk(poolMu_);
        free_.push_back(buf);
    }
    poolCv_.notify_all();
}

template <typename T>
void Plan<T>::ensure_live(size_t slot)
{
    if (live_.count(slot))
        return;
    if (slot >= numSlots_)
        throw Error("time loop would touch slot " + std::to_string(slot) + " but u has only " +
                    std::to_string(numSlots_) + " slots");
    T *buf = acquire();
    // whole allocation (guards and row padding included) starts from zero
    SW_CUDA(cudaMemsetAsync((char *)(buf - g_.lpad - guard_), 0, fieldBytes_, stream_));
    if (!slotZero_[slot])
        upload_dense(hostU_ + slot * hostSlotStride_, buf);
    live_[slot] = buf;
    dirty_[slot] = false;
}

// Slot leaves the device: drained to the caller's array if it was written.
// keep == true leaves the slot live (plan download).
template <typename T>
void Plan<T>::retire(size_t slot, bool keep)
{
    T *buf = live_.at(slot);
    const bool dirty = dirty_[slot];
    if (!keep) {
        live_.erase(slot);
        dirty_.erase(slot);
    }
    if (!dirty) {
        if (!keep)
            give_back(buf);
        return;
    }
    HostDrain::Job job;
    job.src = buf + (long long)outPlaneLo_ * g_.planeStride;
    job.dst = hostU_ + slot * hostSlotStride_ + outPlaneLo_ * (size_t)g_.nM * g_.nF;
    job.rowBytes = (size_t)g_.nF * sizeof(T);
    job.srcPitchBytes = (size_t)g_.pitch * sizeof(T);
    job.rows = (outPlaneHi_ - outPlaneLo_) * (size_t)g_.nM;
    SW_CUDA(cudaEventCreateWithFlags(&job.ready, cudaEventDisableTiming));
    SW_CUDA(cudaEventRecord(job.ready, stream_));
    if (!keep)
        job.done = [this, buf] { give_back(buf); };
    drain_->submit(std::move(job));
}

template <typename T>
void Plan<T>::retire_below(size_t bound)
{
    while (!live_.empty() && live_.begin()->first < bound)
        retire(live_.begin()->first, false);
}

template <typename T>
void Plan<T>::choose_tiling()
{
    useTiled_ = false;
    constexpr bool kF32 = std::is_same<T, float>::value;
    {
        // every 3D variant but the stride-quirk case (float64:
        // sw_step_tiled3d64.cuh)
        if (opt_.simple || ndim_ != 3 || (varden_ && args_.quirk))
            return;
        const int r = g_.r;
        // default configuration, overridable as SIMWAVE_CUDA_TILE=<cfg>[:<zchunk>]
        int maxSmem = 0;
        SW_CUDA(cudaDeviceGetAttribute(&maxSmem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device_));
        // default: configuration 0; large-radius variable density prefers the
        // deeper stream ring of configuration 5 (then 7) where it fits (FAST layout)
        int cfg = 0;
        if (kF32 && varden_ && r > 5) {
            for (int c : {5, 7})
                if (kTiledQuery[r](c, varden_, opt_.math, &tiledInfo_) &&
                    tiledInfo_.smemBytes <= maxSmem) {
                    cfg = c;
                    break;
                }
        }
        int zchunk = 0;
        const char *tileEnv = std::getenv("SIMWAVE_CUDA_TILE");
        if (tileEnv) {
            cfg = std::atoi(tileEnv);
            if (const char *c = std::strchr(tileEnv, ':'))
                zchunk = std::atoi(c + 1);
        }

        // How a configuration would split S into chunks on this grid: the CTAs
        // should fill whole waves (a wave = SMs x resident CTAs), against the
        // cost of a chunk's 2r priming planes of u_cur (re-read by the
        // neighbouring chunk).  Returns the number of chunks; *score rates the
        // configuration: wave fill x priming overhead / bytes a point moves
        // between L2 and the SMs (the halo of the u_cur tile included).
        int sms = 148, maxSmemSm = 0;
        SW_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device_));
        SW_CUDA(cudaDeviceGetAttribute(&maxSmemSm, cudaDevAttrMaxSharedMemoryPerMultiprocessor,
                                       device_));
        const int interior = g_.nS - 2 * r;
        auto chunking = [&](const TiledInfo &info, double *score) {
            const long long tF = (g_.nF - 2 * r + info.tileF() - 1) / info.tileF();
            const long long tM = (g_.nM - 2 * r + info.tileM() - 1) / info.tileM();
            const int threads = info.tx * info.ty + 32;
            int resident = std::max(1, std::min(maxSmemSm / (info.smemBytes + 1024),
                                                2048 / threads));
            resident = std::min(resident, info.minBlocks);
            const double slots = (double)sms * resident;
            const double haloBytes = (double)sizeof(T) * (info.tileM() + 2 * r) *
                                     (info.tileF() + 2 * halo_f(r)) /
                                     ((double)info.tileM() * info.tileF());
            const double alg = (varden_ ? 36.0 : 20.0) * sizeof(T) / 4;
            double best = -1;
            int bestChunks = 1;
            for (int chunks = 1; chunks <= 64 && interior / chunks >= 2 * r; chunks++) {
                const int len = (interior + chunks - 1) / chunks;
                const int real = (interior + len - 1) / len;
                const double waves = tF * tM * real / slots;
                const double fill = waves / std::ceil(waves);
                const double traffic = alg / (alg + 2.0 * r * haloBytes / len);
                const double sc = fill * traffic;
                if (sc > best + 1e-9) { best = sc; bestChunks = chunks; }
            }
            if (score)
                *score = best / (alg - sizeof(T) + haloBytes);
            return bestChunks;
        };

        // float32, constant density, r <= 5: the 14 x 64 tile (two CTAs per SM)
        // or the 30 x 64 tile (one CTA of 16 warps: a fifth less halo traffic,
        // which is what counts once a long loop runs under the power cap --
        // C3 over 2651 steps: 286 against 273 Gpts/s; 512^3 shots: 329 against
        // 301 -- while short bursts on C3 prefer the small tile, 300 against
        // 286), whichever rates higher on this grid
        // Variable density, r <= 5: likewise the 22 x 64 tile (one CTA of 12
        // warps, three stream stages) against the 16 x 64 tile (two CTAs); the
        // large tile measured 195 against 178 Gpts/s on a 512^3 so-8 model and
        // wins a near tie (profiles/r02_sweep_vd_so8_sustained.txt).
        if (kF32 && r <= 5 && !tileEnv) {
            const int large = varden_ ? 3 : 7;
            double score[2] = {-1, -1};
            const int cand[2] = {0, large};
            for (int i = 0; i < 2; i++) {
                TiledInfo info{};
                if (kTiledQuery[r](cand[i], varden_, opt_.math, &info) && info.smemBytes <= maxSmem)
                    chunking(info, &score[i]);
            }
            cfg = (score[1] >= (varden_ ? 0.97 : 1.0) * score[0]) ? large : 0;
        }
        if (kF32) {
            if (!kTiledQuery[r](cfg, varden_, opt_.math, &tiledInfo_))
                throw Error("SIMWAVE_CUDA_TILE: no such tile configuration");
        } else {
            cfg = 0;
            kTiled64Query[r](varden_, opt_.math, &tiledInfo_);
        }
        if (tiledInfo_.smemBytes > maxSmem)
            return;   // plain kernel
        tiledCfg_ = cfg;
        const long long tilesF = (g_.nF - 2 * r + tiledInfo_.tileF() - 1) / tiledInfo_.tileF();
        const long long tilesM = (g_.nM - 2 * r + tiledInfo_.tileM() - 1) / tiledInfo_.tileM();
        if (zchunk <= 0) {
            const int chunks = chunking(tiledInfo_, nullptr);
            zchunk = (interior + chunks - 1) / chunks;
        }
        zChunk_ = std::max(1, std::min(zchunk, interior));
        useTiled_ = true;
        // few sources whose windows lie among the interior points: the step
        // kernel adds them itself (SIMWAVE_CUDA_SOURCES=kernel keeps the launch)
        srcFusedTiled_ = kF32 && srcInterior_ && nsrc_ <= 8 && srcMode_ != SRC_ATOMIC &&
                         !env_is("SIMWAVE_CUDA_SOURCES", "kernel");

        // per (plane, tile) flag: does the damping profile act inside the tile?
        if (model_->tileM != tiledInfo_.tileM() || model_->tileF != tiledInfo_.tileF()) {
            model_->qflags.alloc((size_t)g_.nS * tilesM * tilesF);
            SW_CUDA(cudaMemsetAsync(model_->qflags.get(), 0, model_->qflags.bytes(), stream_));
            dim3 grid((unsigned)tilesF, (unsigned)tilesM, (unsigned)interior);
            qflag_kernel<T><<<grid, 128, 0, stream_>>>(g_, field_base(model_->q),
                                                       tiledInfo_.tileM(), tiledInfo_.tileF(),
                                                       model_->qflags.as<unsigned char>());
            check_launch("qflag_kernel");
            if (varden_ && !model_->frF.get()) {
                new_field(model_->frF);
                new_field(model_->frM);
                new_field(model_->frS);
                dim3 gg((g_.nF - 2 * r + 127) / 128, g_.nM - 2 * r, interior);
                if (opt_.math == MATH_STRICT)
                    rho_gradient_kernel<T, MATH_STRICT><<<gg, 128, 0, stream_>>>(
                        args_, field_base(model_->frF), field_base(model_->frM),
                        field_base(model_->frS));
                else
                    rho_gradient_kernel<T, MATH_FAST><<<gg, 128, 0, stream_>>>(
                        args_, field_base(model_->frF), field_base(model_->frM),
                        field_base(model_->frS));
                check_launch("rho_gradient_kernel");
            }
            model_->tileM = tiledInfo_.tileM();
            model_->tileF = tiledInfo_.tileF();
        }
        if (std::getenv("SIMWAVE_CUDA_VERBOSE"))
            std::fprintf(stderr,
                         "simwave_b200: tiled 3D kernel cfg %d, tile %dx%d, z chunk %d (%d chunks), "
                         "grid %lldx%lldx%d, smem %d B\n",
                         tiledCfg_, tiledInfo_.tileM(), tiledInfo_.tileF(), zChunk_,
                         (interior + zChunk_ - 1) / zChunk_, tilesF, tilesM,
                         (interior + zChunk_ - 1) / zChunk_, tiledInfo_.smemBytes);
    }
}

template <typename T>
const CUtensorMap &Plan<T>::field_map(const T *base, bool halo)
{
    const auto key = std::make_pair((const void *)base, halo);
    auto it = maps_.find(key);
    if (it != maps_.end())
        return it->second;
    CUtensorMap m;
    // the tensor starts at the beginning of the padded row, so that its base
    // is 16-byte aligned for any radius; coordinates are shifted by lpad
    void *rowStart = (void *)(base - g_.lpad);
    const cuuint64_t dims[3] = {(cuuint64_t)g_.pitch, (cuuint64_t)g_.nM, (cuuint64_t)g_.nS};
    const cuuint64_t strides[2] = {(cuuint64_t)g_.pitch * sizeof(T),
                                   (cuuint64_t)g_.planeStride * sizeof(T)};
    const int rp = halo_f(g_.r);
    const cuuint32_t box[3] = {(cuuint32_t)(tiledInfo_.tileF() + (halo ? 2 * rp : 0)),
                               (cuuint32_t)(tiledInfo_.tileM() + (halo ? 2 * g_.r : 0)), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapDataType dtype = sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                                     : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    CUresult rc = encode_tiled_fn()(&m, dtype, 3, rowStart, dims, strides,
                                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS)
        throw Error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)rc));
    return maps_.emplace(key, m).first->second;
}

// One time step of the stencil.
template <typename T>
void Plan<T>::launch_step(const StepArgs<T> &a)
{
    if constexpr (std::is_same<T, double>::value) {
        if (useTiled_) {
            StepMaps maps;
            std::memset(&maps, 0, sizeof(maps));
            maps.prefetch = opt_.prefetch >= 0 ? opt_.prefetch : 2;
            maps.cur = field_map(a.cur, true);
            maps.prev = field_map(a.prev, false);
            maps.c0 = field_map(a.c0, false);
            maps.q = field_map(a.q, false);
            if (varden_) {
                maps.rho = field_map(a.rho, false);
                maps.frF = field_map(field_base(model_->frF), false);
                maps.frM = field_map(field_base(model_->frM), false);
                maps.frS = field_map(field_base(model_->frS), false);
            }
            if (!kTiled64Launch[g_.r](varden_, opt_.math, a, maps,
                                      model_->qflags.as<unsigned char>(), zChunk_, stream_))
                throw Error("tiled float64 kernel vanished");
            check_launch("tiled float64 step kernel");
            return;
        }
    }
    if constexpr (std::is_same<T, float>::value) {
        if (useTiled_) {
            StepMaps maps;
            // two planes ahead: +3 % on C3, +6 % on the so-16 variable-density slab;
            // beyond ~4 the prefetched tiles start evicting each other from L2
            maps.prefetch = opt_.prefetch >= 0 ? opt_.prefetch : 2;
            maps.cur = field_map(a.cur, true);
            maps.prev = field_map(a.prev, false);
            maps.c0 = field_map(a.c0, false);
            maps.q = field_map(a.q, false);
            if (varden_) {
                maps.rho = field_map(a.rho, false);
                maps.frF = field_map(field_base(model_->frF), false);
                maps.frM = field_map(field_base(model_->frM), false);
                maps.frS = field_map(field_base(model_->frS), false);
            }
            maps.srcFused = srcFusedTiled_ ? 1 : 0;
            maps.src = srcTab_;
            maps.wavelet = wavelet_.as<float>();
            maps.waveletCount = (int)waveletCount_;
            maps.step = (long long)stepNow_;
            for (int ax = 0; ax < 3; ax++) {
                maps.srcLo[ax] = srcBox_[ax][0];
                maps.srcHi[ax] = srcBox_[ax][1];
            }
            auto launch = [&](const StepArgs<T> &args) {
                if (!kTiledLaunch[g_.r](tiledCfg_, varden_, opt_.math, args, maps,
                                        model_->qflags.as<unsigned char>(), zChunk_, stream_))
                    throw Error("tiled kernel configuration vanished");
                check_launch("tiled step kernel");
            };
            launch(a);
            return;
        }
    }
    stepSimple_(opt_.math, a, stream_);
    check_launch("step kernel");
}

template <typename T>
void Plan<T>::launch_sources(const StepArgs<T> &a, size_t n)
{
    if (!nsrc_ || srcFusedTiled_)
        return;
    dim3 grid((srcMaxPoints_ + 127) / 128, (unsigned)std::min<size_t>(nsrc_, 65535), 1);
    if (ndim_ == 3)
        source_kernel<T, 3><<<grid, 128, 0, stream_>>>(a, srcTab_, wavelet_.as<T>(),
                                                       (int)waveletCount_, (long long)n, srcMode_);
    else
        source_kernel<T, 2><<<grid, 128, 0, stream_>>>(a, srcTab_, wavelet_.as<T>(),
                                                       (int)waveletCount_, (long long)n, srcMode_);
    check_launch("source_kernel");
}

// Receivers of step n sample U^n (= `cur`), which step n only reads: the
// kernel runs on a side stream beside the step kernel.  It starts once step
// n-1 has been stored (event on the compute stream); the compute stream in
// turn makes step n+1 wait for it (event on the side stream), long before
// any kernel overwrites the slot.
template <typename T>
void Plan<T>::launch_receivers(const T *cur, size_t n)
{
    if (!nrec_)
        return;
    T *row = recOut_.as<T>() + (n - 1) * nrec_;
    const unsigned blocks = (unsigned)((nrec_ * 32 + 127) / 128);
    cudaStream_t st = stream_;
    if (recStream_) {
        // step n's kernels (queued after this call) wait for every earlier
        // receiver launch: whatever slot they overwrite has been sampled
        if (recPending_)
            SW_CUDA(cudaStreamWaitEvent(stream_, recDone_, 0));
        SW_CUDA(cudaEventRecord(recReady_, stream_));
        SW_CUDA(cudaStreamWaitEvent(recStream_, recReady_, 0));
        st = recStream_;
    }
    const bool strict = opt_.math == MATH_STRICT;
    if (ndim_ == 3) {
        if (strict)
            receiver_kernel<T, 3, MATH_STRICT><<<blocks, 128, 0, st>>>(g_, cur, recTab_, row);
        else
            receiver_kernel<T, 3, MATH_FAST><<<blocks, 128, 0, st>>>(g_, cur, recTab_, row);
    } else {
        if (strict)
            receiver_kernel<T, 2, MATH_STRICT><<<blocks, 128, 0, st>>>(g_, cur, recTab_, row);
        else
            receiver_kernel<T, 2, MATH_FAST><<<blocks, 128, 0, st>>>(g_, cur, recTab_, row);
    }
    check_launch("receiver_kernel");
    if (recStream_) {
        SW_CUDA(cudaEventRecord(recDone_, recStream_));
        recPending_ = true;
    }
}

template <typename T>
void Plan<T>::launch_boundaries(T *next)
{
    // F, then M, then S (3d/wave.c:311-480; 2d/wave.c:285-393)
    const int n[3] = {g_.nS, g_.nM, g_.nF};
    for (int axis = AX_F; axis >= (ndim_ == 3 ? AX_S : AX_M); axis--) {
        const int before = args_.bc[2 * axis], after = args_.bc[2 * axis + 1];
        if (!before && !after)
            continue;
        long long lines = 1;
        for (int ax = (ndim_ == 3 ? AX_S : AX_M); ax <= AX_F; ax++)
            if (ax != axis)
                lines *= n[ax] - 2 * g_.r;
        boundary_axis_kernel<T><<<(unsigned)((lines + 127) / 128), 128, 0, stream_>>>(
            g_, next, axis, before, after);
        check_launch("boundary_axis_kernel");
    }
}

template <typename T>
void Plan<T>::run(size_t begin, size_t end)
{
    if (begin < 1 || end > waveletSize_)
        throw Error("timestep range outside [1, wavelet_size]");
    if ((slabUp_ || slabDown_) && !slabConnected_)
        throw Error("slab plan is not connected to its neighbours");
    NvtxRange nvtx("simwave_b200: time loop");
    const double t0 = wall();
    SW_CUDA(cudaEventRecord(evBegin_, stream_));
    if (recBegin_ == recEnd_) { recBegin_ = begin - 1; recEnd_ = begin - 1; }

    const bool persistent = run_persistent(begin, end);
    for (size_t n = begin; n <= end && !persistent; n++) {
        // slot indices, as constant_density/3d/wave.c:113-124
        if (stride_ == 0) {
            prevT_ = (n - 1) % 3; curT_ = n % 3; nextT_ = (n + 1) % 3;
        } else if (stride_ == 1) {
            prevT_ = n - 1; curT_ = n; nextT_ = n + 1;
        }
        if (stride_ != 0)
            retire_below(std::min(prevT_, std::min(curT_, nextT_)));
        ensure_live(prevT_);
        ensure_live(curT_);
        ensure_live(nextT_);

        StepArgs<T> a = args_;
        a.prev = live_[prevT_];
        a.cur = live_[curT_];
        a.next = live_[nextT_];
        if (slabFused_) {
            // shifted so that a local index in my outermost owned planes
            // addresses the ghost copy on the neighbour
            if (slabUp_)
                a.peer[0] = peerSlot_[0][nextT_] +
                            (long long)(peerNS_[0] - 2 * g_.r) * g_.planeStride;
            if (slabDown_)
                a.peer[1] = peerSlot_[1][nextT_] -
                            (long long)(g_.nS - 2 * g_.r) * g_.planeStride;
        }

        // the halo of step n-1 arrives only if that step ran since the last
        // reset (all slabs run the same ranges); the first step of a fresh
        // range reads the ghost planes that were uploaded
        if ((slabUp_ || slabDown_) && ranHi_ != 0 && n - 1 >= ranLo_ && n - 1 <= ranHi_)
            slab_wait(n);
        if (ranHi_ == 0 || n != ranHi_ + 1)
            ranLo_ = n;
        ranHi_ = n;
        stepNow_ = n;
        launch_receivers(a.cur, n);
        launch_step(a);
        launch_sources(a, n);
        if (!a.fuse_bc)
            launch_boundaries(a.next);
        dirty_[nextT_] = true;
        if (slabUp_ || slabDown_)
            slab_push(n, nextT_);

        // slot bookkeeping for saving_stride > 1, as 3d/wave.c:569-618
        if (stride_ > 1) {
            if (n % stride_ == 1) {
                prevT_ = curT_;
                curT_ += 1;
                nextT_ += 1;
                if (stride_ % 2 == 0 && n < end) {
                    std::swap(curT_, nextT_);
                    // the reference exchanges the two slots' contents; here
                    // the slots exchange their buffers
                    ensure_live(curT_);
                    ensure_live(nextT_);
                    std::swap(live_[curT_], live_[nextT_]);
                    dirty_[curT_] = dirty_[nextT_] = true;
                }
            } else {
                prevT_ = curT_;
                curT_ = nextT_;
                nextT_ = prevT_;
            }
        }
    }
    recEnd_ = std::max(recEnd_, end);
    if (recPending_) {
        SW_CUDA(cudaStreamWaitEvent(stream_, recDone_, 0));   // the loop ends with its receivers
        recPending_ = false;
    }
    SW_CUDA(cudaEventRecord(evEnd_, stream_));
    SW_CUDA(cudaEventSynchronize(evEnd_));
    if (slabUp_ || slabDown_) {
        int flags[4] = {0, 0, 0, 0};
        SW_CUDA(cudaMemcpy(flags, slabFlags_.get(), sizeof(flags), cudaMemcpyDeviceToHost));
        if (flags[2])
            throw Error("slab halo exchange timed out waiting for a neighbour");
    }
    float ms = 0;
    SW_CUDA(cudaEventElapsedTime(&ms, evBegin_, evEnd_));
    timing.loop = ms * 1e-3;
    timing.run_wall = wall() - t0;
}

// 2D, three rotating slots: the whole range in one cooperative launch
// (sw_loop2d.cuh).  Returns false when the per-step path has to run.
template <typename T>
bool Plan<T>::run_persistent(size_t begin, size_t end)
{
    timing.loopKind = 0;
    if (ndim_ != 2 || stride_ != 0 || opt_.perStep || opt_.debug || begin > end)
        return false;
    for (size_t s = 0; s < 3; s++)
        ensure_live(s);
    LoopArgs<T> L;
    std::memset(&L, 0, sizeof(L));
    L.a = args_;
    for (size_t s = 0; s < 3; s++)
        L.slot[s] = live_[s];
    L.src = srcTab_;
    L.rec = recTab_;
    L.wavelet = wavelet_.as<T>();
    L.waveletCount = (int)waveletCount_;
    L.srcMode = srcMode_;
    L.srcMaxPoints = srcMaxPoints_;
    L.fuseSources = (srcInterior_ && nsrc_ <= 8 && srcMode_ != SRC_ATOMIC) ? 1 : 0;
    L.srcLoM = srcBox_[AX_M][0]; L.srcHiM = srcBox_[AX_M][1];
    L.srcLoF = srcBox_[AX_F][0]; L.srcHiF = srcBox_[AX_F][1];
    L.recOut = recOut_.as<T>();
    L.begin = (long long)begin;
    L.end = (long long)end;
    if (!loopBarrier_.get())
        loopBarrier_.alloc(128);
    SW_CUDA(cudaMemsetAsync(loopBarrier_.get(), 0, 128, stream_));
    L.barrier = loopBarrier_.as<unsigned>();
    const bool trace = std::getenv("SIMWAVE_CUDA_LOOP2D_TRACE") != nullptr;
    DeviceBuffer traceBuf;
    if (trace) {
        traceBuf.alloc(256 * sizeof(unsigned long long));
        SW_CUDA(cudaMemsetAsync(traceBuf.get(), 0, traceBuf.bytes(), stream_));
        L.trace = traceBuf.as<unsigned long long>();
    }
    // SIMWAVE_CUDA_LOOP2D=resident: the tile-resident loop
    // (sw_loop2d_resident.cuh) where the problem fits one co-resident wave of
    // tiles; otherwise, and by default, the grid-barrier loop
    bool ok = run_resident(L);
    timing.loopKind = ok ? 2 : 1;
    if (!ok)
        ok = varden_ ? launch_loop2d<T, true>(opt_.math, L, stream_)
                     : launch_loop2d<T, false>(opt_.math, L, stream_);
    if (!ok) {
        timing.loopKind = 0;
        return false;
    }
    check_launch("loop2d kernel");
    if (trace) {
        unsigned long long t[256];
        SW_CUDA(cudaMemcpyAsync(t, traceBuf.get(), sizeof(t), cudaMemcpyDeviceToHost, stream_));
        SW_CUDA(cudaStreamSynchronize(stream_));
        std::fprintf(stderr, "simwave_b200: loop2d phases of CTA 0 (ns: receivers, stencil, "
                             "sources+barriers | per step):");
        for (int i = 4 * 20; i + 4 < 4 * 26 && t[i + 4]; i += 4)
            std::fprintf(stderr, " [%llu %llu %llu]", t[i + 1] - t[i], t[i + 2] - t[i + 1],
                         t[i + 3] - t[i + 2]);
        std::fprintf(stderr, "\n");
    }
    for (size_t n = begin; n <= end; n++)
        dirty_[(n + 1) % 3] = true;
    prevT_ = (end - 1) % 3; curT_ = end % 3; nextT_ = (end + 1) % 3;
    return true;
}

// The tile-resident 2D loop: needs fused boundary conditions, sources added by
// the owning thread (or none), and a tiling that fits one wave of CTAs.
template <typename T>
bool Plan<T>::run_resident(LoopArgs<T> &L)
{
    // opt-in: measured slower than the grid-barrier loop on B200 (see the
    // header of sw_loop2d_resident.cuh)
    if (residentState_ < 0 || !env_is("SIMWAVE_CUDA_LOOP2D", "resident"))
        return false;
    if (!args_.fuse_bc || (nsrc_ > 0 && !L.fuseSources))
        return false;
    if (residentState_ == 0) {
        residentState_ = -1;
        Loop2dTiling tl{};
        const bool fits = varden_ ? loop2d_resident_tiling<T, true>(opt_.math, g_, recMaxM_,
                                                                    recMaxF_, &tl)
                                  : loop2d_resident_tiling<T, false>(opt_.math, g_, recMaxM_,
                                                                     recMaxF_, &tl);
        if (!fits)
            return false;
        residentTiling_ = tl;
        // receivers grouped by the tile that owns the first cell of their window
        // (a window that starts in the halo belongs to the first tile of the axis)
        const int tiles = tl.tilesM * tl.tilesF, r = g_.r;
        std::vector<int> start(tiles + 1, 0), index(std::max<size_t>(1, nrec_));
        auto owner = [&](size_t i) {
            const int tmi = std::min(std::max(recLoM_[i] - r, 0) / tl.tm, tl.tilesM - 1);
            const int tfi = std::min(std::max(recLoF_[i] - r, 0) / tl.tf, tl.tilesF - 1);
            return tmi * tl.tilesF + tfi;
        };
        for (size_t i = 0; i < nrec_; i++)
            start[owner(i) + 1]++;
        for (int t = 0; t < tiles; t++)
            start[t + 1] += start[t];
        std::vector<int> fill(start.begin(), start.end() - 1);
        for (size_t i = 0; i < nrec_; i++)
            index[fill[owner(i)]++] = (int)i;
        residentRecStart_.alloc(start.size() * sizeof(int));
        residentRecIndex_.alloc(index.size() * sizeof(int));
        residentFlags_.alloc(tiles * sizeof(unsigned));
        SW_CUDA(cudaMemcpyAsync(residentRecStart_.get(), start.data(), start.size() * sizeof(int),
                                cudaMemcpyHostToDevice, stream_));
        SW_CUDA(cudaMemcpyAsync(residentRecIndex_.get(), index.data(), index.size() * sizeof(int),
                                cudaMemcpyHostToDevice, stream_));
        SW_CUDA(cudaStreamSynchronize(stream_));     // the vectors die here
        residentState_ = 1;
        if (std::getenv("SIMWAVE_CUDA_VERBOSE"))
            std::fprintf(stderr,
                         "simwave_b200: tile-resident 2D loop, %d x %d tiles of %d x %d points, "
                         "%zu B of shared memory per CTA\n",
                         tl.tilesM, tl.tilesF, tl.tm, tl.tf, tl.smemBytes);
    }
    const Loop2dTiling &tl = residentTiling_;
    L.tm = tl.tm; L.tf = tl.tf; L.tilesM = tl.tilesM; L.tilesF = tl.tilesF;
    L.flags = residentFlags_.as<unsigned>();
    L.recStart = residentRecStart_.as<int>();
    L.recIndex = residentRecIndex_.as<int>();
    SW_CUDA(cudaMemsetAsync(residentFlags_.get(), 0, residentFlags_.bytes(), stream_));
    const bool ok = varden_ ? launch_loop2d_resident<T, true>(opt_.math, L, tl, stream_)
                            : launch_loop2d_resident<T, false>(opt_.math, L, tl, stream_);
    if (!ok)
        residentState_ = -1;
    return ok;
}

// Writable page-table entries for [p, p + bytes) without changing a byte:
// MADV_POPULATE_WRITE where the kernel has it (5.14+), else an atomic OR of
// zero into one byte per page.
static void populate_write(char *p, size_t bytes, const std::atomic<bool> &stop)
{
    const size_t page = 4096, chunk = 32u << 20;
    char *b = (char *)((uintptr_t)p & ~(uintptr_t)(page - 1));
    char *e = p + bytes;
    bool useMadvise = true;
    for (char *c = b; c < e && !stop.load(std::memory_order_relaxed); c += chunk) {
        const size_t n = std::min<size_t>(chunk, (size_t)(e - c));
#ifdef MADV_POPULATE_WRITE
        if (useMadvise && madvise(c, n, MADV_POPULATE_WRITE) == 0)
            continue;
#endif
        useMadvise = false;
        for (char *q = std::max(c, p); q < c + n; q = (char *)(((uintptr_t)q & ~(uintptr_t)(page - 1)) + page))
            __atomic_fetch_or(q, 0, __ATOMIC_RELAXED);
    }
}

template <typename T>
void Plan<T>::prefault_outputs(size_t end)
{
    if (!hostU_ || (opt_.outMode == 2 && stride_ == 0))
        return;
    const size_t slotBytes = denseCells_ * sizeof(T);
    if (slotBytes < (16u << 20) || is_pinned_host(hostU_))
        return;
    // what download() will write: every slot, or (returned-slot hint, three
    // rotating slots) only slot end % 3
    size_t first = 0, count = numSlots_;
    if (stride_ == 0 && opt_.outMode == 1) {
        first = end % 3;
        count = 1;
    }
    join_prefault();
    prefaultStop_.store(false);
    // slot by slot: a slab plan's slots are windows of the caller's larger array
    prefault_ = std::thread([this, first, count, slotBytes] {
        widen_helper_affinity();
        for (size_t s = first; s < first + count; s++)
            populate_write((char *)(hostU_ + s * hostSlotStride_), slotBytes, prefaultStop_);
    });
}

template <typename T>
void Plan<T>::download(void *u, void *receivers)
{
    NvtxRange nvtx("simwave_b200: drain");
    join_prefault();
    const double t0 = wall();
    T *saveU = hostU_;
    if (u)
        hostU_ = (T *)u;
    // which slots the caller wants back: all of them (the ABI's contract),
    // or -- SIMWAVE_HINT_WAVEFIELD_OUT, three rotating slots only -- just slot
    // end_timestep % 3 (the one simwave's Solver returns, model.py:639-641),
    // or none
    std::vector<size_t> slots;
    for (auto &kv : live_)
        if (stride_ != 0 || opt_.outMode == 0 || (opt_.outMode == 1 && kv.first == curT_))
            slots.push_back(kv.first);
    for (size_t s : slots)
        retire(s, true);
    drain_->wait_idle();
    hostU_ = saveU;
    T *rec = receivers ? (T *)receivers : hostRec_;
    if (rec && nrec_ && recEnd_ > recBegin_) {
        SW_CUDA(cudaMemcpyAsync(rec + recBegin_ * nrec_, recOut_.as<T>() + recBegin_ * nrec_,
                                (recEnd_ - recBegin_) * nrec_ * sizeof(T),
                                cudaMemcpyDeviceToHost, stream_));
        SW_CUDA(cudaStreamSynchronize(stream_));
    }
    timing.d2h = wall() - t0;
}

template <typename T>
void Plan<T>::reset()
{
    drain_->wait_idle();
    SW_CUDA(cudaStreamSynchronize(stream_));
    const bool slab = slabUp_ || slabDown_;
    if (slab) {
        // the neighbours hold mappings of these buffers: refill them in place
        for (size_t s = 0; s < 3; s++) {
            T *buf = live_.at(s);
            SW_CUDA(cudaMemsetAsync((char *)(buf - g_.lpad - guard_), 0, fieldBytes_, stream_));
            if (!slotZero_[s])
                upload_dense(hostU_ + s * hostSlotStride_, buf);
            dirty_[s] = false;
        }
        SW_CUDA(cudaMemsetAsync(slabFlags_.get(), 0, slabFlags_.bytes(), stream_));
    } else {
        for (auto &kv : live_)
            give_back(kv.second);
        live_.clear();
        dirty_.clear();
    }
    prevT_ = 0; curT_ = 1; nextT_ = 2;
    recBegin_ = recEnd_ = 0;
    ranLo_ = ranHi_ = 0;
    SW_CUDA(cudaMemsetAsync(recOut_.get(), 0, recOut_.bytes(), stream_));
    timing.launches = 0;
    SW_CUDA(cudaStreamSynchronize(stream_));
}

template <typename T>
void Plan<T>::slab_export(void *out)
{
    if (!(slabUp_ || slabDown_))
        throw Error("not a slab plan");
    SlabDesc d;
    std::memset(&d, 0, sizeof(d));
    d.magic = kSlabMagic;
    d.dtypeBytes = (int)sizeof(T);
    d.nS = g_.nS; d.nM = g_.nM; d.nF = g_.nF; d.r = g_.r; d.lpad = g_.lpad;
    d.device = device_;
    d.pitch = g_.pitch; d.planeStride = g_.planeStride;
    for (size_t s = 0; s < 3; s++)
        ipc_export(live_.at(s), &d.slot[s], &d.slotOffset[s]);
    ipc_export(slabFlags_.get(), &d.flags, &d.flagsOffset);
    std::memset(out, 0, SIMWAVE_SLAB_DESC_BYTES);
    std::memcpy(out, &d, sizeof(d));
}

template <typename T>
void Plan<T>::slab_connect(const void *up, const void *down)
{
    if ((slabUp_ && !up) || (slabDown_ && !down))
        throw Error("slab_connect: missing neighbour descriptor");
    const void *descs[2] = {slabUp_ ? up : nullptr, slabDown_ ? down : nullptr};
    for (int side = 0; side < 2; side++) {
        if (!descs[side])
            continue;
        SlabDesc d;
        std::memcpy(&d, descs[side], sizeof(d));
        if (d.magic != kSlabMagic || d.dtypeBytes != (int)sizeof(T) || d.nM != g_.nM ||
            d.nF != g_.nF || d.r != g_.r || d.pitch != g_.pitch || d.lpad != g_.lpad)
            throw Error("slab_connect: neighbour slab has a different plane layout");
        auto open = [&](const cudaIpcMemHandle_t &h) {
            void *p = nullptr;
            SW_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
            ipcMapped_.push_back(p);
            return (char *)p;
        };
        for (int s = 0; s < 3; s++)
            peerSlot_[side][s] = (T *)(open(d.slot[s]) + d.slotOffset[s]);
        // my up neighbour records my deliveries in ITS "from down" word and vice versa
        peerFlags_[side] = (int *)(open(d.flags) + d.flagsOffset) + (side == 0 ? 1 : 0);
        peerNS_[side] = d.nS;
    }
    slabConnected_ = true;
    // ghost copies are written by the kernels themselves (overlapped with the
    // rest of the step) unless something would bypass them: separate boundary
    // kernels, atomically accumulated sources, or SIMWAVE_CUDA_SLAB_PUSH=copy
    slabFused_ = args_.fuse_bc && srcMode_ != SRC_ATOMIC &&
                 !env_is("SIMWAVE_CUDA_SLAB_PUSH", "copy");
}

template <typename T>
void Plan<T>::slab_peer(SlabPeer *out)
{
    if (!(slabUp_ || slabDown_))
        throw Error("not a slab plan");
    for (size_t s = 0; s < 3; s++)
        out->slot[s] = live_.at(s);
    out->flags = slabFlags_.as<int>();
    out->nS = g_.nS; out->nM = g_.nM; out->nF = g_.nF; out->r = g_.r; out->lpad = g_.lpad;
    out->pitch = g_.pitch;
    out->device = device_;
    out->dtypeBytes = (int)sizeof(T);
}

// Neighbours living in the same process (one plan per device, one host thread
// each): peer access instead of IPC mappings, otherwise the same protocol.
template <typename T>
void Plan<T>::slab_connect_direct(const SlabPeer *up, const SlabPeer *down)
{
    if ((slabUp_ && !up) || (slabDown_ && !down))
        throw Error("slab_connect_direct: missing neighbour");
    const SlabPeer *peers[2] = {slabUp_ ? up : nullptr, slabDown_ ? down : nullptr};
    for (int side = 0; side < 2; side++) {
        const SlabPeer *d = peers[side];
        if (!d)
            continue;
        if (d->dtypeBytes != (int)sizeof(T) || d->nM != g_.nM || d->nF != g_.nF ||
            d->r != g_.r || d->pitch != g_.pitch || d->lpad != g_.lpad)
            throw Error("slab_connect_direct: neighbour slab has a different plane layout");
        if (d->device != device_) {
            int can = 0;
            SW_CUDA(cudaDeviceCanAccessPeer(&can, device_, d->device));
            if (!can)
                throw Error("device " + std::to_string(device_) + " cannot access device " +
                            std::to_string(d->device) + " (no peer path)");
            cudaError_t e = cudaDeviceEnablePeerAccess(d->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                throw Error(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
            cudaGetLastError();
        }
        for (int s = 0; s < 3; s++)
            peerSlot_[side][s] = (T *)d->slot[s];
        peerFlags_[side] = d->flags + (side == 0 ? 1 : 0);
        peerNS_[side] = d->nS;
    }
    slabConnected_ = true;
    slabFused_ = args_.fuse_bc && srcMode_ != SRC_ATOMIC &&
                 !env_is("SIMWAVE_CUDA_SLAB_PUSH", "copy");
}

// before step n reads the ghost planes of u_cur: the neighbours must have
// delivered the halo they produced in step n-1
template <typename T>
void Plan<T>::slab_wait(size_t n)
{
    // SIMWAVE_CUDA_SLAB_TIMEOUT: seconds a step waits for its neighbours'
    // halos before the run is declared broken (default 20; ranks that start
    // with more skew than that need a larger value)
    static const double seconds = [] {
        const char *e = std::getenv("SIMWAVE_CUDA_SLAB_TIMEOUT");
        const double v = e ? std::atof(e) : 0.0;
        return v > 0 ? v : 20.0;
    }();
    slab_wait_kernel<<<1, 1, 0, stream_>>>(slabFlags_.as<int>(), slabUp_ ? 1 : 0,
                                           slabDown_ ? 1 : 0, (int)n - 1,
                                           (long long)(seconds * 2.0e9));
    check_launch("slab_wait_kernel");
}

// after step n: my outermost owned planes of u_next become the neighbours'
// ghost planes (whole padded planes, so F/M halo cells travel too), then the
// per-step flag is published in the neighbour's memory
template <typename T>
void Plan<T>::slab_push(size_t n, size_t slot)
{
    const int r = g_.r;
    const size_t planeBytes = (size_t)g_.planeStride * sizeof(T);
    const T *mine = live_.at(slot);
    for (int side = 0; side < 2; side++) {
        if (!(side == 0 ? slabUp_ : slabDown_))
            continue;
        const long long srcPlane = (side == 0) ? r : g_.nS - 2 * r;
        const long long dstPlane = (side == 0) ? peerNS_[side] - r : 0;
        if (!slabFused_)
            SW_CUDA(cudaMemcpyAsync(peerSlot_[side][slot] + dstPlane * g_.planeStride - g_.lpad,
                                    mine + srcPlane * g_.planeStride - g_.lpad, r * planeBytes,
                                    cudaMemcpyDefault, stream_));
        slab_publish_kernel<<<1, 1, 0, stream_>>>(peerFlags_[side], (int)n);
        check_launch("slab_publish_kernel");
    }
}

std::unique_ptr<PlanBase> make_plan(const simwave_problem &pb, const Options &opt)
{
    if (pb.dtype_bytes == 4)
        return std::unique_ptr<PlanBase>(new Plan<float>(pb, opt));
    if (pb.dtype_bytes == 8)
        return std::unique_ptr<PlanBase>(new Plan<double>(pb, opt));
    throw Error("dtype_bytes must be 4 or 8");
}

}  // namespace sw
