// csrc/common.cu -- host-side state of libxyz_b200.so: launch counter, scratch arenas, version.
#include "common.cuh"

#include <map>
#include <mutex>
#include <utility>
#include <vector>

namespace xyzb {

std::atomic<uint64_t> g_launch_count{0};

namespace {
constexpr int kMaxDevices = 64;
struct Arena {
    void* ptr = nullptr;
    size_t bytes = 0;
    uint64_t generation = 0;  // bumped whenever ptr changes: holders of old pointers compare before dereferencing
    bool captured = false;    // handed out while its stream was capturing: a CUDA graph may hold raw pointers into it
};
struct StreamArenas {
    Arena slot[SCRATCH_SLOTS];
};
// Library-owned scratch is keyed by (device, stream): launches on different streams of one device never share tickets,
// partial rows or tile lists, so they may be in flight together.  (Two host threads driving ONE stream concurrently is
// a misuse of that stream's ordering, here as everywhere in CUDA.)
std::map<std::pair<int, cudaStream_t>, StreamArenas> g_arenas;
std::vector<std::pair<int, void*>> g_retired;  // buffers a captured graph may still reference: freed at shutdown only
uint64_t g_generation = 0;
int g_sm_count[kMaxDevices] = {0};
std::mutex g_mu;
}  // namespace

int scratch_get(ScratchSlot slot, size_t bytes, void** ptr, cudaStream_t stream, uint64_t* generation) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return static_cast<int>(e);
    if (dev < 0 || dev >= kMaxDevices) return XYZ_ERR_INVALID_ARGUMENT;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    e = cudaStreamIsCapturing(stream, &cap);
    if (e != cudaSuccess) return static_cast<int>(e);
    const bool capturing = cap != cudaStreamCaptureStatusNone;
    std::lock_guard<std::mutex> lock(g_mu);
    Arena& a = g_arenas[{dev, stream}].slot[slot];
    if (a.bytes < bytes) {
        // growing means cudaMalloc / cudaMemset / a synchronisation, none of which a capturing stream allows: the
        // caller warms the path up outside the capture (or brings its own workspace, xyz_*_ws)
        if (capturing) return XYZ_ERR_WORKSPACE;
        if (a.ptr) {
            if (a.captured) {
                g_retired.emplace_back(dev, a.ptr);  // a graph may replay on it: never freed before shutdown
            } else {
                e = cudaStreamSynchronize(stream);  // only this stream's launches use the old buffer
                if (e != cudaSuccess) return static_cast<int>(e);
                cudaFree(a.ptr);
            }
            a.ptr = nullptr;
            a.bytes = 0;
            a.captured = false;
        }
        size_t want = bytes + bytes / 4 + 256;  // 25 % head-room: lists grow slowly during training
        e = cudaMalloc(&a.ptr, want);
        if (e != cudaSuccess) {
            a.ptr = nullptr;
            return static_cast<int>(e);
        }
        e = cudaMemset(a.ptr, 0, want);  // tickets / counters start at zero; kernels leave them at zero
        if (e != cudaSuccess) return static_cast<int>(e);
        a.bytes = want;
        a.generation = ++g_generation;
    }
    if (capturing) a.captured = true;
    *ptr = a.ptr;
    if (generation) *generation = a.generation;
    return 0;
}

bool scratch_is_current(ScratchSlot slot, cudaStream_t stream, uint64_t generation) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_arenas.find({dev, stream});
    return it != g_arenas.end() && it->second.slot[slot].ptr != nullptr && it->second.slot[slot].generation == generation;
}

void scratch_free_all() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return;
    std::lock_guard<std::mutex> lock(g_mu);
    cudaDeviceSynchronize();
    for (auto it = g_arenas.begin(); it != g_arenas.end();) {
        if (it->first.first == dev) {
            for (int s = 0; s < SCRATCH_SLOTS; ++s)
                if (it->second.slot[s].ptr) cudaFree(it->second.slot[s].ptr);
            it = g_arenas.erase(it);
        } else {
            ++it;
        }
    }
    for (auto it = g_retired.begin(); it != g_retired.end();) {
        if (it->first == dev) {
            cudaFree(it->second);
            it = g_retired.erase(it);
        } else {
            ++it;
        }
    }
}

int sm_count() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return kNumSM;
    if (g_sm_count[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSM;
        g_sm_count[dev] = n;
    }
    return g_sm_count[dev];
}

}  // namespace xyzb

extern "C" {
const char* xyz_b200_version(void) { return "xyz_b200 0.1 (sm_100a)"; }
int xyz_b200_shutdown(void) {
    xyzb::scratch_free_all();
    return 0;
}
uint64_t xyz_b200_launch_count(void) { return xyzb::g_launch_count.load(); }
void xyz_b200_reset_launch_count(void) { xyzb::g_launch_count.store(0); }
}
