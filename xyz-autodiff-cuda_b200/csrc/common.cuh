// csrc/common.cuh -- shared device/host helpers of libxyz_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstring>

#include "../../include/xyz_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libxyz_b200 is written for sm_100a (B200) only"
#endif

namespace xyzb {

constexpr int kNumSM = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// ---- host side ---------------------------------------------------------------------------------
extern std::atomic<uint64_t> g_launch_count;
inline void count_launch(int n = 1) { g_launch_count.fetch_add(static_cast<uint64_t>(n), std::memory_order_relaxed); }

inline int last_error() { return static_cast<int>(cudaGetLastError()); }

// Library-owned scratch, one growable arena per (device, STREAM, slot): calls in flight on different streams never share
// tickets, partial rows or tile lists.  Grown with cudaMalloc on first use or when a launch needs more (never while
// the stream is capturing: XYZ_ERR_WORKSPACE); never shrinks; a buffer handed out during a capture is never freed
// before xyz_b200_shutdown() (a graph may replay on it).  `generation` changes whenever the arena's pointer does.
enum ScratchSlot : int { SCRATCH_REDUCE = 0, SCRATCH_SPLAT = 1, SCRATCH_SPLAT_SORT = 2, SCRATCH_SLOTS = 3 };
int scratch_get(ScratchSlot slot, size_t bytes, void** ptr, cudaStream_t stream, uint64_t* generation = nullptr);
bool scratch_is_current(ScratchSlot slot, cudaStream_t stream, uint64_t generation);
void scratch_free_all();
int sm_count();  // SMs of the current device (148 on B200)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- peer mailboxes (csrc/peer.cu) ---------------------------------------------------------------
struct PeerMailbox {
    double data[2][XYZ_PEER_MAX_WORLD][XYZ_PEER_SLOT_DOUBLES];   // small rows (least squares: 5 doubles)
    unsigned long long flag[XYZ_PEER_MAX_WORLD];                 // shared by both areas: one sequence per group
    float vec[2][XYZ_PEER_MAX_WORLD][XYZ_PEER_VEC_FLOATS];       // K-bin rows (accumulation), K <= XYZ_PEER_VEC_FLOATS
    // the splat optimiser exchange (xyz_adam_step_individual_peer) has its own sequence space, kept ON THE DEVICE so that
    // a captured CUDA graph can be replayed: every rank's kernel reads splat_seq / splat_iter of its OWN mailbox and
    // advances them when it is done (all ranks make the same calls, so the counters agree)
    unsigned long long flag_splat[XYZ_PEER_MAX_WORLD];
    double loss_splat[2][XYZ_PEER_MAX_WORLD];
    unsigned long long splat_seq;   // last sequence number used (two per call)
    unsigned long long splat_iter;  // optimiser steps taken with iteration == 0 ("count them yourself")
    unsigned int splat_ticket;      // last-CTA election of the owner's kernel (left at 0 by every launch)
};
struct PeerArgs {  // passed to kernels by value; world <= 1 means "no exchange"
    PeerMailbox* box[XYZ_PEER_MAX_WORLD];
    int rank, world;
    unsigned long long seq;
};

// ---- device side -------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_relaxed_sys_f32(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// All-reduce (sum, rank order) of `len` <= XYZ_PEER_SLOT_DOUBLES doubles held by threads 0..len-1 of ONE CTA per
// rank.  Must be called by every thread of the CTA (it synchronises); returns the sum in threads 0..len-1.
// If a peer does not show up within ~4 s the result is NaN (loud, not a hang).
__device__ __forceinline__ double peer_allreduce_cta(const PeerArgs& pa, double mine, int len, int tid) {
    const int par = static_cast<int>(pa.seq & 1ull);
    if (tid < len) {
        for (int p = 0; p < pa.world; ++p) pa.box[p]->data[par][pa.rank][tid] = mine;  // NVLink stores (and one local)
        __threadfence_system();
    }
    __syncthreads();
    __shared__ int s_ok;
    if (tid == 0) s_ok = 1;
    __syncthreads();
    if (tid < pa.world) {
        st_release_sys(&pa.box[tid]->flag[pa.rank], pa.seq);
        const unsigned long long* f = &pa.box[pa.rank]->flag[tid];
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(f) < pa.seq) {
            if (global_timer_ns() - t0 > 4000000000ull) {
                s_ok = 0;
                break;
            }
        }
    }
    __syncthreads();
    double s = 0.0;
    if (tid < len) {
        for (int q = 0; q < pa.world; ++q) s += ld_relaxed_sys_f64(&pa.box[pa.rank]->data[par][q][tid]);
        if (!s_ok) s = __longlong_as_double(0x7ff8000000000000ll);
    }
    return s;
}

// Publish + wait half of the same protocol for rows the caller has already stored into every rank's
// vec[seq & 1][rank][...] (and fenced with __threadfence_system()).  Every thread of the CTA must call it.
// Returns false on timeout.
__device__ __forceinline__ bool peer_publish_and_wait_cta(const PeerArgs& pa, int tid) {
    __shared__ int s_ok2;
    if (tid == 0) s_ok2 = 1;
    __syncthreads();  // also orders the callers' fenced row stores before the flags
    if (tid < pa.world) {
        st_release_sys(&pa.box[tid]->flag[pa.rank], pa.seq);
        const unsigned long long* f = &pa.box[pa.rank]->flag[tid];
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(f) < pa.seq) {
            if (global_timer_ns() - t0 > 4000000000ull) {
                s_ok2 = 0;
                break;
            }
        }
    }
    __syncthreads();
    return s_ok2 != 0;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// mbarrier (shared::cta) -- used as the "full" barrier of TMA bulk-copy pipelines
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    // make the initialised barrier visible to the async proxy (TMA unit)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// TMA 1-D bulk copies (SASS: UBLKCP).  bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(__cvta_generic_to_global(gmem_src)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(gmem_dst)),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// generic-proxy writes to shared memory -> visible to the async proxy (before a bulk store)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// vector reduction into global memory, one instruction for 4 floats (sm_90+): REDG.E.ADD.F32x4
__device__ __forceinline__ void red_add_v4(float* addr16, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(__cvta_generic_to_global(addr16)), "f"(a), "f"(b),
                 "f"(c), "f"(d)
                 : "memory");
}
__device__ __forceinline__ void red_add_v2(float* addr8, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(__cvta_generic_to_global(addr8)), "f"(a), "f"(b) : "memory");
}

template <class T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#endif  // __CUDACC__

}  // namespace xyzb
