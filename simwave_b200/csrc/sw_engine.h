// Time-loop engine: a problem resident on one device.
#pragma once

#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/simwave_cuda.h"
#include "sw_common.h"

namespace sw {

struct Options {
    int math;          // MATH_STRICT / MATH_FAST
    bool simple;       // force the plain kernels
    bool debug;        // synchronise + check after every launch
    bool separateBc;   // stand-alone boundary kernels instead of the fused form
    int prefetch;      // planes of L2 prefetch ahead of the tiled kernel's ring loads (-1: default)
    bool perStep;      // 2D: per-step launches instead of the persistent loop kernel
    int device;        // -1: current
    // promises of the caller (simwave_cuda_set_hint), all off by default
    bool zeroIn = false;        // every slot of u is zero on entry
    int outMode = 0;            // slots copied back: 0 all, 1 slot end_timestep % 3 only, 2 none
    long long modelToken = 0;   // != 0: model arrays unchanged while the token is
    static Options from_env();
};

struct Timing {
    double loop = 0, h2d = 0, d2h = 0, total = 0;
    double run_wall = 0, teardown = 0;   // host wall of run(); plan destruction
    unsigned long long launches = 0;
    int loopKind = 0;   // time loop of the last run: 0 per-step launches, 1 grid-barrier 2D loop, 2 tile-resident 2D loop
};

// RAII device allocation
class DeviceBuffer {
public:
    DeviceBuffer() = default;
    explicit DeviceBuffer(size_t bytes) { alloc(bytes); }
    ~DeviceBuffer() { release(); }
    DeviceBuffer(const DeviceBuffer &) = delete;
    DeviceBuffer &operator=(const DeviceBuffer &) = delete;
    DeviceBuffer(DeviceBuffer &&o) noexcept : ptr_(o.ptr_), bytes_(o.bytes_), device_(o.device_)
    {
        o.ptr_ = nullptr; o.bytes_ = 0;
    }
    DeviceBuffer &operator=(DeviceBuffer &&o) noexcept
    {
        if (this != &o) {
            release();
            ptr_ = o.ptr_; bytes_ = o.bytes_; device_ = o.device_;
            o.ptr_ = nullptr; o.bytes_ = 0;
        }
        return *this;
    }
    void alloc(size_t bytes);
    void release();
    void *get() const { return ptr_; }
    template <typename U> U *as() const { return static_cast<U *>(ptr_); }
    size_t bytes() const { return bytes_; }

private:
    void *ptr_ = nullptr;
    size_t bytes_ = 0;
    int device_ = 0;
};

// Process-wide caches behind DeviceBuffer and the pinned staging buffers: a
// survey calls `forward` once per shot with the same shapes, and cudaMalloc /
// cudaFree / cudaMallocHost of GB-sized blocks cost more than the copies they
// serve.  Blocks are handed back to the driver when an allocation fails, when
// the cache would exceed half of the device memory, through
// simwave_cuda_release_cache(), or never (SIMWAVE_CUDA_CACHE=0 disables it).
void *pinned_take(size_t bytes);
void pinned_give(void *p, size_t bytes);
void release_caches();
// bracket of one drop-in forward(): afterwards the caches hold what this call
// used and nothing older (SIMWAVE_CUDA_CACHE=keep turns the trimming off)
void cache_begin_call();
void cache_end_call();
size_t cached_device_bytes();       // device memory held by the cache, all devices

// true if [p, p+1) is page-locked host memory known to the CUDA driver
bool is_pinned_host(const void *p);
// helper threads: escape a one-core CPU mask inherited from a bound main thread
void widen_helper_affinity();
// memcpy split over a few threads (pageable <-> pinned staging copies)
void parallel_memcpy(void *dst, const void *src, size_t bytes);

// Drains pitched device fields into dense caller memory on a side stream:
// device -> pinned staging (cudaMemcpy2DAsync, double buffered) -> memcpy into
// the (pageable) destination, all on a worker thread so the launch thread
// keeps queueing time steps.  This is how saving_stride snapshots and the
// final slots leave the device.
class HostDrain {
public:
    struct Job {
        const void *src;        // device, pitched, element (0,0,0)
        void *dst;              // host, dense
        size_t rowBytes;        // nF * sizeof(T)
        size_t srcPitchBytes;
        size_t rows;            // nS * nM
        cudaEvent_t ready;      // recorded on the compute stream (owned by the job)
        std::function<void()> done;
    };
    HostDrain(int device, size_t chunkBytes);
    ~HostDrain();
    void submit(Job job);
    void wait_idle();           // rethrows a worker failure
private:
    void worker();
    void ensure_staging();
    int device_;
    size_t chunkBytes_;
    void *pinned_[2] = {nullptr, nullptr};
    cudaStream_t stream_ = nullptr;
    cudaEvent_t copied_[2] = {nullptr, nullptr};
    std::thread thread_;
    std::mutex mu_;
    std::condition_variable cv_, idle_;
    std::deque<Job> queue_;
    bool busy_ = false, stop_ = false;
    std::string error_;
};

// What a slab plan shows its neighbours inside one process (the single-process
// multi-device `forward`): raw device pointers instead of CUDA IPC handles.
struct SlabPeer {
    void *slot[3] = {nullptr, nullptr, nullptr};   // element (0,0,0) of each wavefield slot
    int *flags = nullptr;                          // {from up, from down, error}
    int nS = 0, nM = 0, nF = 0, r = 0, lpad = 0, device = -1, dtypeBytes = 0;
    long long pitch = 0;
};

class PlanBase {
public:
    virtual ~PlanBase() {}
    virtual void run(size_t begin, size_t end) = 0;
    // Drop-in forward() only: fault in the pages of the caller's (pageable) `u`
    // that the drain of [.., end] will write, on a helper thread, while the
    // time loop runs.  simwave's Solver hands over a fresh np.zeros array: its
    // pages do not exist yet, and taking the first-touch page faults inside
    // the drain made it crawl at 5-8 GB/s.
    virtual void prefault_outputs(size_t end) = 0;
    virtual void download(void *u, void *receivers) = 0;
    virtual void reset() = 0;
    virtual void slab_export(void *desc) = 0;
    virtual void slab_connect(const void *up, const void *down) = 0;
    virtual void slab_peer(SlabPeer *out) = 0;
    virtual void slab_connect_direct(const SlabPeer *up, const SlabPeer *down) = 0;
    Timing timing;
};

void drop_resident_models();

std::unique_ptr<PlanBase> make_plan(const simwave_problem &pb, const Options &opt);

Timing &last_timing();

}  // namespace sw
