// wah_device.cu -- per-device launch state shared by the C ABI entry points.
//
// wah_compress_kernel and wah_decode_kernel are persistent grids (SMs x occupancy CTAs) whose CTAs wait for results
// of other CTAs of the same grid.  Two such grids running on one device at the same time (launched on different
// streams, or from different threads) could each end up partly resident and wait for ever, so the library orders its
// own launches per device: a launch on a stream other than the previous launch's stream first waits for an event
// recorded behind that launch.  Launches that follow each other on ONE stream -- the normal pipeline -- pay nothing
// and keep their programmatic-dependent-launch overlap (no event is recorded between them).
//
// The decode kernel's zero-initialised counters also live here: an array of slots per device, one slot per launch in
// launch order; every launch zeroes its own slot when its last CTA leaves AND the slot of the launch after it when its
// first CTA starts, so a slot is clean even if the launch that last used it was killed half way.
#include "wah_kernels.h"

#include <mutex>

namespace wahb200 {

namespace {

constexpr int MAX_DEV = 64;
constexpr uint32_t SLOTS = 4096;

struct DeviceState {
    std::mutex mu;
    bool has_last = false;
    cudaStream_t last_stream = nullptr;
    cudaEvent_t ev = nullptr;
    DecodeCounters *slots = nullptr;
    uint32_t next_slot = 0;
};
DeviceState g_state[MAX_DEV];

}  // namespace

LaunchOrder::LaunchOrder(cudaStream_t stream) : stream_(stream)
{
    err_ = cudaGetDevice(&dev_);
    if (err_ == cudaSuccess && (dev_ < 0 || dev_ >= MAX_DEV)) err_ = cudaErrorInvalidDevice;
    if (err_ != cudaSuccess) {
        dev_ = -1;
        return;
    }
    DeviceState &s = g_state[dev_];
    s.mu.lock();
    if (s.has_last && s.last_stream != stream) {
        if (!s.ev) err_ = cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming);
        if (err_ == cudaSuccess) {
            // everything submitted to the previous launch's stream so far, that launch included
            cudaError_t e = cudaEventRecord(s.ev, s.last_stream);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(stream, s.ev, 0);
            if (e != cudaSuccess) {
                // (the stream may be gone: its work is not, so wait for the device instead)
                cudaGetLastError();
                err_ = cudaDeviceSynchronize();
            }
        }
    }
}

LaunchOrder::~LaunchOrder()
{
    if (dev_ < 0) return;
    DeviceState &s = g_state[dev_];
    s.has_last = true;
    s.last_stream = stream_;
    s.mu.unlock();
}

cudaError_t LaunchOrder::counter_slots(DecodeCounters **mine, DecodeCounters **next)
{
    if (dev_ < 0) return err_;
    DeviceState &s = g_state[dev_];   // (locked by the constructor)
    if (!s.slots) {
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, sizeof(DecodeCounters) * SLOTS);
        if (e != cudaSuccess) return e;
        e = cudaMemset(p, 0, sizeof(DecodeCounters) * SLOTS);
        if (e != cudaSuccess) {
            cudaFree(p);
            return e;
        }
        s.slots = static_cast<DecodeCounters *>(p);
    }
    *mine = s.slots + s.next_slot;
    s.next_slot = (s.next_slot + 1u) % SLOTS;
    *next = s.slots + s.next_slot;
    return cudaSuccess;
}

// a stream the library created itself is about to be destroyed (wah_host_release): do not record events on it later
void forget_stream(cudaStream_t stream)
{
    for (DeviceState &s : g_state) {
        std::lock_guard<std::mutex> g(s.mu);
        if (s.has_last && s.last_stream == stream) s.has_last = false;
    }
}

// test hook (wah_test_poison_counter_slots): what a launch killed half way leaves behind, in every slot of the device
cudaError_t poison_counter_slots()
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= MAX_DEV) return cudaErrorInvalidDevice;
    DeviceState &s = g_state[dev];
    std::lock_guard<std::mutex> g(s.mu);
    if (!s.slots) return cudaSuccess;
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return e;
    return cudaMemset(s.slots, 0x5A, sizeof(DecodeCounters) * SLOTS);
}

}  // namespace wahb200
