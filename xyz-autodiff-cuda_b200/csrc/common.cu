// csrc/common.cu -- host-side state of libxyz_b200.so: launch counter, scratch arenas, version.
#include "common.cuh"

#include <mutex>

namespace xyzb {

std::atomic<uint64_t> g_launch_count{0};

namespace {
constexpr int kMaxDevices = 64;
struct Arena {
    void* ptr = nullptr;
    size_t bytes = 0;
};
Arena g_arena[kMaxDevices][SCRATCH_SLOTS];
int g_sm_count[kMaxDevices] = {0};
std::mutex g_mu;
}  // namespace

int scratch_get(ScratchSlot slot, size_t bytes, void** ptr) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return static_cast<int>(e);
    if (dev < 0 || dev >= kMaxDevices) return XYZ_ERR_INVALID_ARGUMENT;
    std::lock_guard<std::mutex> lock(g_mu);
    Arena& a = g_arena[dev][slot];
    if (a.bytes < bytes) {
        if (a.ptr) {
            // the old buffer may still be in use by launches queued on any stream
            e = cudaDeviceSynchronize();
            if (e != cudaSuccess) return static_cast<int>(e);
            cudaFree(a.ptr);
            a.ptr = nullptr;
            a.bytes = 0;
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
    }
    *ptr = a.ptr;
    return 0;
}

void scratch_free_all() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return;
    std::lock_guard<std::mutex> lock(g_mu);
    cudaDeviceSynchronize();
    for (int s = 0; s < SCRATCH_SLOTS; ++s) {
        if (g_arena[dev][s].ptr) cudaFree(g_arena[dev][s].ptr);
        g_arena[dev][s] = Arena{};
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
