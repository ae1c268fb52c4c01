// dev/accum_lab.cu -- development harness (not shipped): variants of the C2 accumulation kernel timed
// side by side on one GPU, each checked against an fp64 host sum.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo dev/accum_lab.cu -o dev/_build/accum_lab
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#define CKV(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr unsigned kFull = 0xffffffffu;

template <class T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// fold the values of lanes in `peers` onto the lowest lane of each group; fixed pairing order
__device__ __forceinline__ float reduce_peers(unsigned peers, float x, int lane) {
    int rel = __popc(peers & ((1u << lane) - 1u));
    unsigned above = peers & ~((2u << lane) - 1u);
    while (__any_sync(kFull, above != 0u)) {
        const int next = __ffs(above);
        const float t = __shfl_sync(kFull, x, (next - 1) & 31);
        if (next) x += t;
        const unsigned alive = __ballot_sync(kFull, (rel & 1) == 0);
        above &= alive;
        rel >>= 1;
    }
    return x;
}

// ---- variant T: per-warp table + byte tag table, arbitration by write-then-read-back ----------------------
template <int MODE>  // 0: tags + serial/ballot-match slow path ; 1: tags + match_any slow path
__device__ __forceinline__ void tag_batch(float* table, uint8_t* tag, int id, float v, int k, int kbits, int lane) {
    const bool valid = static_cast<unsigned>(id) < static_cast<unsigned>(k);
    const int id0 = __shfl_sync(kFull, id, 0);
    if (__all_sync(kFull, id == id0)) {
        const float s = warp_sum(v);
        if (lane == 0 && valid) table[id] += s;
        __syncwarp();
        return;
    }
    if (valid) tag[id] = static_cast<uint8_t>(lane);
    __syncwarp();
    const bool won = valid && tag[id] == lane;
    if (won) table[id] += v;
    const bool mine = valid && !won;
    unsigned lost = __ballot_sync(kFull, mine);
    __syncwarp();
    if (lost == 0u) return;
    if (__popc(lost) <= 2) {
        while (lost) {
            const int l = __ffs(lost) - 1;
            if (lane == l) table[id] += v;
            __syncwarp();
            lost &= lost - 1u;
        }
        return;
    }
    unsigned peers;
    if (MODE == 1) {
        peers = __match_any_sync(kFull, mine ? id : -1);
        if (!mine) peers = 1u << lane;
    } else {
        peers = lost;
        for (int b = 0; b < kbits; ++b) {
            const bool bit = (id >> b) & 1;
            const unsigned m = __ballot_sync(kFull, mine && bit);
            peers &= bit ? m : ~m;
        }
        if (!mine) peers = 1u << lane;
    }
    v = reduce_peers(peers, v, lane);
    if (mine && lane == __ffs(peers) - 1) table[id] += v;
    __syncwarp();
}

template <int WARPS, int DEPTH, int MODE>
__global__ void __launch_bounds__(WARPS * 32, 1)
    accum_tag_kernel(const int32_t* __restrict__ idx, const float* __restrict__ val, long long n, float* grad, int k,
                     int kbits) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tables = reinterpret_cast<float*>(smem_raw);
    uint8_t* tags = reinterpret_cast<uint8_t*>(tables + static_cast<size_t>(WARPS) * k);
    constexpr int kThreads = WARPS * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < WARPS * k; i += kThreads) tables[i] = 0.f;
    __syncthreads();
    float* table = tables + static_cast<size_t>(warp) * k;
    uint8_t* tag = tags + static_cast<size_t>(warp) * k;

    const long long n_chunks = (n + 127) / 128;
    const long long gw = static_cast<long long>(blockIdx.x) * WARPS + warp;
    const long long wstride = static_cast<long long>(gridDim.x) * WARPS;

    int4 idb[DEPTH];
    float4 vb[DEPTH];
    auto load_chunk = [&](long long c, int4& q, float4& f) {
        const long long e0 = c * 128 + lane * 4;
        if (c < n_chunks && e0 + 4 <= n) {
            q = __ldcs(reinterpret_cast<const int4*>(idx + e0));
            f = __ldcs(reinterpret_cast<const float4*>(val + e0));
        } else {
            int t[4]; float u[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long e = e0 + j;
                const bool in = (c < n_chunks) && (e < n);
                t[j] = in ? __ldg(idx + e) : -1;
                u[j] = in ? __ldg(val + e) : 0.f;
            }
            q = make_int4(t[0], t[1], t[2], t[3]);
            f = make_float4(u[0], u[1], u[2], u[3]);
        }
    };
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) load_chunk(gw + d * wstride, idb[d], vb[d]);
    for (long long c = gw; c < n_chunks; c += wstride * DEPTH) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            const int4 q = idb[d];
            const float4 f = vb[d];
            load_chunk(c + (DEPTH + d) * wstride, idb[d], vb[d]);
            if (c + d * wstride < n_chunks) {
                tag_batch<MODE>(table, tag, q.x, f.x, k, kbits, lane);
                tag_batch<MODE>(table, tag, q.y, f.y, k, kbits, lane);
                tag_batch<MODE>(table, tag, q.z, f.z, k, kbits, lane);
                tag_batch<MODE>(table, tag, q.w, f.w, k, kbits, lane);
            }
        }
    }
    __syncthreads();
    for (int b = tid; b < k; b += kThreads) {
        float s = 0.f;
#pragma unroll 8
        for (int w = 0; w < WARPS; ++w) s += tables[static_cast<size_t>(w) * k + b];
        if (s != 0.f) atomicAdd(grad + b, s);
    }
}


// ---- variant H: tags + register-cached hot bins ------------------------------------------------------------
// HOT: number of per-warp hot ids whose contributions are summed in lane-private registers (no shared memory).
// SAMECHK: 0 none, 1 per batch (shfl + vote.all)
// NOTAG: timing probe only (wrong results when a batch has duplicates)
template <int WARPS, int DEPTH, int HOT, int SAMECHK, bool NOTAG, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
    accum_hot_kernel(const int32_t* __restrict__ idx, const float* __restrict__ val, long long n, float* grad, int k,
                     int kbits) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tables = reinterpret_cast<float*>(smem_raw);
    uint8_t* tags = reinterpret_cast<uint8_t*>(tables + static_cast<size_t>(WARPS) * k);
    constexpr int kThreads = WARPS * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < WARPS * k; i += kThreads) tables[i] = 0.f;
    __syncthreads();
    float* table = tables + static_cast<size_t>(warp) * k;
    uint8_t* tag = tags + static_cast<size_t>(warp) * k;

    const long long n_chunks = (n + 127) / 128;
    const long long gw = static_cast<long long>(blockIdx.x) * WARPS + warp;
    const long long wstride = static_cast<long long>(gridDim.x) * WARPS;

    int hot_id[HOT > 0 ? HOT : 1];
    float hot_acc[HOT > 0 ? HOT : 1];
#pragma unroll
    for (int h = 0; h < HOT; ++h) { hot_id[h] = -2 - h; hot_acc[h] = 0.f; }
    int hot_next = 0;

    auto batch = [&](int id, float v) {
        bool valid = static_cast<unsigned>(id) < static_cast<unsigned>(k);
        if (HOT > 0) {
            bool hit = false;
#pragma unroll
            for (int h = 0; h < HOT; ++h) {
                const bool p = (id == hot_id[h]);
                hot_acc[h] += p ? v : 0.f;
                hit |= p;
            }
            valid = valid && !hit;
            if (!__any_sync(kFull, valid)) return;
        }
        if (SAMECHK == 1) {
            const int id0 = __shfl_sync(kFull, id, 0);
            if (__all_sync(kFull, id == id0)) {
                const float s = warp_sum(valid ? v : 0.f);
                if (lane == 0 && valid) table[id] += s;
                __syncwarp();
                return;
            }
        }
        if (NOTAG) {
            if (valid) table[id] += v;
            __syncwarp();
            return;
        }
        if (valid) tag[id] = static_cast<uint8_t>(lane);
        __syncwarp();
        const bool won = valid && tag[id] == lane;
        if (won) table[id] += v;
        const bool mine = valid && !won;
        unsigned lost = __ballot_sync(kFull, mine);
        __syncwarp();
        if (lost == 0u) return;
        if (__popc(lost) <= 2) {
            while (lost) {
                const int l = __ffs(lost) - 1;
                if (lane == l) table[id] += v;
                __syncwarp();
                lost &= lost - 1u;
            }
            return;
        }
        unsigned peers = __match_any_sync(kFull, mine ? id : -1);
        if (!mine) peers = 1u << lane;
        v = reduce_peers(peers, v, lane);
        if (mine && lane == __ffs(peers) - 1) table[id] += v;
        __syncwarp();
        if (HOT > 0) {
            // adopt the id of the biggest group of losers as a hot id (replace round robin)
            const int cnt = mine ? __popc(peers) : 0;
            int best = cnt;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(kFull, best, o));
            if (best >= 3) {
                const unsigned who = __ballot_sync(kFull, mine && cnt == best);
                const int cand = __shfl_sync(kFull, id, __ffs(who) - 1);
                // flush the slot being replaced
#pragma unroll
                for (int h = 0; h < HOT; ++h) {
                    if (h == hot_next) {
                        const float s = warp_sum(hot_acc[h]);
                        if (lane == 0 && hot_id[h] >= 0) table[hot_id[h]] += s;
                        hot_acc[h] = 0.f;
                        hot_id[h] = cand;
                    }
                }
                __syncwarp();
                hot_next = (hot_next + 1 == HOT) ? 0 : hot_next + 1;
            }
        }
    };

    int4 idb[DEPTH];
    float4 vb[DEPTH];
    auto load_chunk = [&](long long c, int4& q, float4& f) {
        const long long e0 = c * 128 + lane * 4;
        if (c < n_chunks && e0 + 4 <= n) {
            q = __ldcs(reinterpret_cast<const int4*>(idx + e0));
            f = __ldcs(reinterpret_cast<const float4*>(val + e0));
        } else {
            int t[4]; float u[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long e = e0 + j;
                const bool in = (c < n_chunks) && (e < n);
                t[j] = in ? __ldg(idx + e) : -1;
                u[j] = in ? __ldg(val + e) : 0.f;
            }
            q = make_int4(t[0], t[1], t[2], t[3]);
            f = make_float4(u[0], u[1], u[2], u[3]);
        }
    };
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) load_chunk(gw + d * wstride, idb[d], vb[d]);
    for (long long c = gw; c < n_chunks; c += wstride * DEPTH) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            const int4 q = idb[d];
            const float4 f = vb[d];
            load_chunk(c + (DEPTH + d) * wstride, idb[d], vb[d]);
            if (c + d * wstride < n_chunks) {
                batch(q.x, f.x);
                batch(q.y, f.y);
                batch(q.z, f.z);
                batch(q.w, f.w);
            }
        }
    }
#pragma unroll
    for (int h = 0; h < HOT; ++h) {
        const float s = warp_sum(hot_acc[h]);
        if (lane == 0 && hot_id[h] >= 0) table[hot_id[h]] += s;
    }
    __syncthreads();
    for (int b = tid; b < k; b += kThreads) {
        float s = 0.f;
#pragma unroll 8
        for (int w = 0; w < WARPS; ++w) s += tables[static_cast<size_t>(w) * k + b];
        if (s != 0.f) atomicAdd(grad + b, s);
    }
}


// ---- variant A: tags + ADAPTIVE register-cached hot bins ---------------------------------------------------
template <int WARPS, int DEPTH, int HOT, int ADOPT, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB)
    accum_adaptive_kernel(const int32_t* __restrict__ idx, const float* __restrict__ val, long long n, float* grad, int k) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tables = reinterpret_cast<float*>(smem_raw);
    uint8_t* tags = reinterpret_cast<uint8_t*>(tables + static_cast<size_t>(WARPS) * k);
    constexpr int kThreads = WARPS * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < WARPS * k; i += kThreads) tables[i] = 0.f;
    __syncthreads();
    float* table = tables + static_cast<size_t>(warp) * k;
    uint8_t* tag = tags + static_cast<size_t>(warp) * k;

    const long long n_chunks = (n + 127) / 128;
    const long long gw = static_cast<long long>(blockIdx.x) * WARPS + warp;
    const long long wstride = static_cast<long long>(gridDim.x) * WARPS;

    int hot_id[HOT];
    float hot_acc[HOT];
#pragma unroll
    for (int h = 0; h < HOT; ++h) { hot_id[h] = -2; hot_acc[h] = 0.f; }
    bool hot_mode = false;
    int hits = 0, window = 0;

    auto flush_slot = [&](int id_, float acc_) {
        const float s = warp_sum(acc_);
        if (lane == 0 && id_ >= 0) table[id_] += s;
    };

    auto batch = [&](int id, float v) {
        bool valid = static_cast<unsigned>(id) < static_cast<unsigned>(k);
        if (hot_mode) {
            bool hit = false;
#pragma unroll
            for (int h = 0; h < HOT; ++h) {
                const bool p = (id == hot_id[h]);
                hot_acc[h] += p ? v : 0.f;
                hit |= p;
            }
            valid = valid && !hit;
            const unsigned rest = __ballot_sync(kFull, valid);
            hits += 32 - __popc(rest);
            if (++window == 32) {
                if (hits < 64) {  // fewer than 2 lanes per batch: not worth the compares
#pragma unroll
                    for (int h = 0; h < HOT; ++h) { flush_slot(hot_id[h], hot_acc[h]); hot_id[h] = -2; hot_acc[h] = 0.f; }
                    __syncwarp();
                    hot_mode = false;
                }
                hits = 0;
                window = 0;
            }
            if (rest == 0u) return;
        }
        if (valid) tag[id] = static_cast<uint8_t>(lane);
        __syncwarp();
        const bool won = valid && tag[id] == lane;
        if (won) table[id] += v;
        const bool mine = valid && !won;
        unsigned lost = __ballot_sync(kFull, mine);
        __syncwarp();
        if (lost == 0u) return;
        if (__popc(lost) <= 2) {
            while (lost) {
                const int l = __ffs(lost) - 1;
                if (lane == l) table[id] += v;
                __syncwarp();
                lost &= lost - 1u;
            }
            return;
        }
        unsigned peers = __match_any_sync(kFull, mine ? id : -1);
        if (!mine) peers = 1u << lane;
        v = reduce_peers(peers, v, lane);
        if (mine && lane == __ffs(peers) - 1) table[id] += v;
        __syncwarp();
        // a bin that lost >= ADOPT-1 lanes in one batch becomes a register-cached hot bin (FIFO replacement)
        const unsigned big = __ballot_sync(kFull, mine && __popc(peers) >= ADOPT - 1);
        if (big) {
            const int cand = __shfl_sync(kFull, id, __ffs(big) - 1);
            flush_slot(hot_id[HOT - 1], hot_acc[HOT - 1]);
            __syncwarp();
#pragma unroll
            for (int h = HOT - 1; h > 0; --h) { hot_id[h] = hot_id[h - 1]; hot_acc[h] = hot_acc[h - 1]; }
            hot_id[0] = cand;
            hot_acc[0] = 0.f;
            if (!hot_mode) { hot_mode = true; hits = 0; window = 0; }
        }
    };

    int4 idb[DEPTH];
    float4 vb[DEPTH];
    auto load_chunk = [&](long long c, int4& q, float4& f) {
        const long long e0 = c * 128 + lane * 4;
        if (c < n_chunks && e0 + 4 <= n) {
            q = __ldcs(reinterpret_cast<const int4*>(idx + e0));
            f = __ldcs(reinterpret_cast<const float4*>(val + e0));
        } else {
            int t[4]; float u[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long e = e0 + j;
                const bool in = (c < n_chunks) && (e < n);
                t[j] = in ? __ldg(idx + e) : -1;
                u[j] = in ? __ldg(val + e) : 0.f;
            }
            q = make_int4(t[0], t[1], t[2], t[3]);
            f = make_float4(u[0], u[1], u[2], u[3]);
        }
    };
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) load_chunk(gw + d * wstride, idb[d], vb[d]);
    for (long long c = gw; c < n_chunks; c += wstride * DEPTH) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            const int4 q = idb[d];
            const float4 f = vb[d];
            load_chunk(c + (DEPTH + d) * wstride, idb[d], vb[d]);
            if (c + d * wstride < n_chunks) {
                batch(q.x, f.x);
                batch(q.y, f.y);
                batch(q.z, f.z);
                batch(q.w, f.w);
            }
        }
    }
#pragma unroll
    for (int h = 0; h < HOT; ++h) flush_slot(hot_id[h], hot_acc[h]);
    __syncthreads();
    for (int b = tid; b < k; b += kThreads) {
        float s = 0.f;
#pragma unroll 8
        for (int w = 0; w < WARPS; ++w) s += tables[static_cast<size_t>(w) * k + b];
        if (s != 0.f) atomicAdd(grad + b, s);
    }
}


// ---- variant G: tags for most batches, plain global REDs (per-CTA table in L2) for every GLOBAL_EVERY-th batch ----
template <int WARPS, int GLOBAL_MASK>
__global__ void __launch_bounds__(WARPS * 32, 1)
    accum_mixed_kernel(const int32_t* __restrict__ idx, const float* __restrict__ val, long long n, float* grad, int k,
                       float* gtables) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tables = reinterpret_cast<float*>(smem_raw);
    uint8_t* tags = reinterpret_cast<uint8_t*>(tables + static_cast<size_t>(WARPS) * k);
    constexpr int kThreads = WARPS * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* gtab = gtables + static_cast<size_t>(blockIdx.x) * k;
    for (int i = tid; i < WARPS * k; i += kThreads) tables[i] = 0.f;
    for (int i = tid; i < k; i += kThreads) gtab[i] = 0.f;
    __syncthreads();
    float* table = tables + static_cast<size_t>(warp) * k;
    uint8_t* tag = tags + static_cast<size_t>(warp) * k;
    const long long n_chunks = (n + 127) / 128;
    const long long gw = static_cast<long long>(blockIdx.x) * WARPS + warp;
    const long long wstride = static_cast<long long>(gridDim.x) * WARPS;
    auto batch = [&](int id, float v) {
        const bool valid = static_cast<unsigned>(id) < static_cast<unsigned>(k);
        if (valid) tag[id] = static_cast<uint8_t>(lane);
        __syncwarp();
        const bool won = valid && tag[id] == lane;
        if (won) table[id] += v;
        const bool mine = valid && !won;
        unsigned lost = __ballot_sync(kFull, mine);
        __syncwarp();
        if (lost == 0u) return;
        if (__popc(lost) <= 2) {
            while (lost) {
                const int l = __ffs(lost) - 1;
                if (lane == l) table[id] += v;
                __syncwarp();
                lost &= lost - 1u;
            }
            return;
        }
        unsigned peers = __match_any_sync(kFull, mine ? id : -1);
        if (!mine) peers = 1u << lane;
        v = reduce_peers(peers, v, lane);
        if (mine && lane == __ffs(peers) - 1) table[id] += v;
        __syncwarp();
    };
    auto gbatch = [&](int id, float v) {
        if (static_cast<unsigned>(id) < static_cast<unsigned>(k)) atomicAdd(gtab + id, v);
    };
    for (long long c = gw; c < n_chunks; c += wstride) {
        const long long e0 = c * 128 + lane * 4;
        if (e0 + 4 <= n) {
            const int4 q = __ldcs(reinterpret_cast<const int4*>(idx + e0));
            const float4 f = __ldcs(reinterpret_cast<const float4*>(val + e0));
            if (GLOBAL_MASK & 1) gbatch(q.x, f.x); else batch(q.x, f.x);
            if (GLOBAL_MASK & 2) gbatch(q.y, f.y); else batch(q.y, f.y);
            if (GLOBAL_MASK & 4) gbatch(q.z, f.z); else batch(q.z, f.z);
            if (GLOBAL_MASK & 8) gbatch(q.w, f.w); else batch(q.w, f.w);
        }
    }
    __syncthreads();
    __threadfence();
    for (int b = tid; b < k; b += kThreads) {
        float s = 0.f;
#pragma unroll 8
        for (int w = 0; w < WARPS; ++w) s += tables[static_cast<size_t>(w) * k + b];
        s += __ldcg(gtab + b);
        if (s != 0.f) atomicAdd(grad + b, s);
    }
}


// ---- variant S: lane-striped tables ------------------------------------------------------------------------
// Every warp owns a table with 16 copies of every bin: copy = lane & 15, address (id * 16 + copy).  Lanes L and
// L ^ 16 are the only two that can touch the same word, so duplicates are found with ONE shuffle of the id (no tag
// table, no match.any) and the update costs 2 + 2 shared-memory wavefronts (the two half-warps hit the same 16 banks
// when their ids have the same parity) instead of ~14.  64 B per bin per warp -> 3 warps per SM at K = 1024; the
// input is staged by 1-D TMA bulk copies into a per-warp ring (the warp's lane 0 is its own producer).
namespace sv {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(__cvta_generic_to_global(gmem_src)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
}  // namespace sv

constexpr int kStripeChunk = 256;  // elements per TMA stage: 1 KB of ids + 1 KB of values

// one batch, general case: pv = the partner lane's value, p = its id
__device__ __forceinline__ void stripe_rmw_slow(float* mytab, int k, int a, float v, int p, float pv, int lane) {
    bool ok = static_cast<unsigned>(a) < static_cast<unsigned>(k);
    if (a == p) {
        if (lane < 16) v += pv; else ok = false;
    }
    if (ok) mytab[a * 16] += v;
    __syncwarp();
}

template <int STAGES>
__global__ void __launch_bounds__(128, 1)
    accum_stripe_kernel(const int32_t* __restrict__ idx, const float* __restrict__ val, long long n_chunks, float* grad, int k) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warps = blockDim.x >> 5;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* tables = reinterpret_cast<float*>(smem_raw);  // [warps][k][16]
    unsigned char* ring = smem_raw + static_cast<size_t>(warps) * k * 64;  // [warps][STAGES][2 KB]
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + static_cast<size_t>(warps) * STAGES * 2048);
    uint64_t* bar = bars + warp * STAGES;
    unsigned char* myring = ring + static_cast<size_t>(warp) * STAGES * 2048;
    const long long gw = static_cast<long long>(blockIdx.x) * warps + warp;
    const long long wstride = static_cast<long long>(gridDim.x) * warps;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) sv::mbar_init(bar + s, 1);
        sv::mbar_fence_init();
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            const long long c = gw + s * wstride;
            if (c < n_chunks) {
                sv::mbar_arrive_expect_tx(bar + s, 2048);
                sv::bulk_load(myring + s * 2048, idx + c * kStripeChunk, 1024, bar + s);
                sv::bulk_load(myring + s * 2048 + 1024, val + c * kStripeChunk, 1024, bar + s);
            }
        }
    }
    {
        float4* t4 = reinterpret_cast<float4*>(tables);
        const int n4 = warps * k * 4;
        for (int i = tid; i < n4; i += blockDim.x) t4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    float* mytab = tables + static_cast<size_t>(warp) * k * 16 + (lane & 15);
    int it = 0;
    for (long long c = gw; c < n_chunks; c += wstride, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = static_cast<uint32_t>(it / STAGES) & 1u;
        sv::mbar_wait(bar + s, ph);
        const int4* si = reinterpret_cast<const int4*>(myring + s * 2048);
        const float4* sf = reinterpret_cast<const float4*>(myring + s * 2048 + 1024);
        const int4 q0 = si[lane], q1 = si[32 + lane];
        const float4 f0 = sf[lane], f1 = sf[32 + lane];
        __syncwarp();
        if (lane == 0) {
            const long long cn = c + static_cast<long long>(STAGES) * wstride;
            if (cn < n_chunks) {
                sv::mbar_arrive_expect_tx(bar + s, 2048);
                sv::bulk_load(myring + s * 2048, idx + cn * kStripeChunk, 1024, bar + s);
                sv::bulk_load(myring + s * 2048 + 1024, val + cn * kStripeChunk, 1024, bar + s);
            }
        }
        const int a[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
        const float v[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
        int p[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) p[j] = __shfl_xor_sync(kFull, a[j], 16);
        bool slow[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int a0 = a[2 * t], a1 = a[2 * t + 1], p0 = p[2 * t], p1 = p[2 * t + 1];
            slow[t] = __any_sync(kFull, (a0 == p0) | (a1 == p1) | (a0 == a1) | (a0 == p1) | (a1 == p0));
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int a0 = a[2 * t], a1 = a[2 * t + 1];
            const float v0 = v[2 * t], v1 = v[2 * t + 1];
            if (!slow[t]) {
                const bool ok0 = static_cast<unsigned>(a0) < static_cast<unsigned>(k);
                const bool ok1 = static_cast<unsigned>(a1) < static_cast<unsigned>(k);
                float t0 = 0.f, t1 = 0.f;
                if (ok0) t0 = mytab[a0 * 16];
                if (ok1) t1 = mytab[a1 * 16];
                if (ok0) mytab[a0 * 16] = t0 + v0;
                if (ok1) mytab[a1 * 16] = t1 + v1;
                __syncwarp();
            } else {
                const float pv0 = __shfl_xor_sync(kFull, v0, 16), pv1 = __shfl_xor_sync(kFull, v1, 16);
                stripe_rmw_slow(mytab, k, a0, v0, p[2 * t], pv0, lane);
                stripe_rmw_slow(mytab, k, a1, v1, p[2 * t + 1], pv1, lane);
            }
        }
    }
    __syncthreads();
    // fold: bin b = 16 copies x warps; the float4 order is rotated by b / 2 so a quarter-warp reads 8 distinct bank groups
    for (int b = tid; b < k; b += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < warps; ++w) {
            const float4* row = reinterpret_cast<const float4*>(tables + (static_cast<size_t>(w) * k + b) * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 t = row[(j + (b >> 1)) & 3];
                s += (t.x + t.y) + (t.z + t.w);
            }
        }
        if (s != 0.f) atomicAdd(grad + b, s);
    }
}


// shared by the S2 / S3 variants: 8 batches held by the warp (lane holds elements 4 lane .. 4 lane + 3 of two 128-blocks)
template <bool RMW>
__device__ __forceinline__ void stripe_process8(float* mytab, int k, const int4 q0, const int4 q1, const float4 f0,
                                                const float4 f1, int lane) {
    const int a[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    const float v[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
    int p[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) p[j] = __shfl_xor_sync(kFull, a[j], 16);
    bool slow[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int a0 = a[2 * t], a1 = a[2 * t + 1], p0 = p[2 * t], p1 = p[2 * t + 1];
        slow[t] = __any_sync(kFull, (a0 == p0) | (a1 == p1) | (a0 == a1) | (a0 == p1) | (a1 == p0));
    }
    if (!RMW) {
        float acc = 0.f;
#pragma unroll
        for (int t = 0; t < 4; ++t) acc += slow[t] ? v[2 * t] : v[2 * t + 1];
        if (acc == 123.456f) mytab[0] = acc;
        return;
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int a0 = a[2 * t], a1 = a[2 * t + 1];
        const float v0 = v[2 * t], v1 = v[2 * t + 1];
        if (!slow[t]) {
            const bool ok0 = static_cast<unsigned>(a0) < static_cast<unsigned>(k);
            const bool ok1 = static_cast<unsigned>(a1) < static_cast<unsigned>(k);
            float t0 = 0.f, t1 = 0.f;
            if (ok0) t0 = mytab[a0 * 16];
            if (ok1) t1 = mytab[a1 * 16];
            if (ok0) mytab[a0 * 16] = t0 + v0;
            if (ok1) mytab[a1 * 16] = t1 + v1;
            __syncwarp();
        } else {
            const float pv0 = __shfl_xor_sync(kFull, v0, 16), pv1 = __shfl_xor_sync(kFull, v1, 16);
            stripe_rmw_slow(mytab, k, a0, v0, p[2 * t], pv0, lane);
            stripe_rmw_slow(mytab, k, a1, v1, p[2 * t + 1], pv1, lane);
        }
    }
}

__device__ __forceinline__ void stripe_fold(const float* tables, int warps, int k, float* grad, int tid, int nthreads) {
    for (int b = tid; b < k; b += nthreads) {
        float s = 0.f;
        for (int w = 0; w < warps; ++w) {
            const float4* row = reinterpret_cast<const float4*>(tables + (static_cast<size_t>(w) * k + b) * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 t = row[(j + (b >> 1)) & 3];
                s += (t.x + t.y) + (t.z + t.w);
            }
        }
        if (s != 0.f) atomicAdd(grad + b, s);
    }
}

// ---- variant S2: striped tables, per-lane cp.async (LDGSTS) ring ---------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sv::smem_u32(smem_dst)), "l"(__cvta_generic_to_global(gmem_src)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int STAGES, bool RMW>
__global__ void __launch_bounds__(128, 1)
    accum_stripe2_kernel(const int32_t* __restrict__ idx, const float* __restrict__ val, long long n_chunks, float* grad, int k) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warps = blockDim.x >> 5;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* tables = reinterpret_cast<float*>(smem_raw);
    unsigned char* ring = smem_raw + static_cast<size_t>(warps) * k * 64;
    unsigned char* mine = ring + static_cast<size_t>(warp) * STAGES * 2048 + lane * 16;  // this lane's 16-byte column
    const long long gw = static_cast<long long>(blockIdx.x) * warps + warp;
    const long long wstride = static_cast<long long>(gridDim.x) * warps;
    auto issue = [&](long long c, int s) {
        if (c < n_chunks) {
            const int32_t* gi = idx + c * kStripeChunk + lane * 4;
            const float* gv = val + c * kStripeChunk + lane * 4;
            unsigned char* d = mine + s * 2048;
            cp_async16(d, gi);
            cp_async16(d + 512, gi + 128);
            cp_async16(d + 1024, gv);
            cp_async16(d + 1536, gv + 128);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < STAGES; ++s) issue(gw + s * wstride, s);
    {
        float4* t4 = reinterpret_cast<float4*>(tables);
        const int n4 = warps * k * 4;
        for (int i = tid; i < n4; i += blockDim.x) t4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    float* mytab = tables + static_cast<size_t>(warp) * k * 16 + (lane & 15);
    int s = 0;
    for (long long c = gw; c < n_chunks; c += wstride) {
        cp_async_wait<STAGES - 1>();
        const unsigned char* d = mine + s * 2048;
        const int4 q0 = *reinterpret_cast<const int4*>(d), q1 = *reinterpret_cast<const int4*>(d + 512);
        const float4 f0 = *reinterpret_cast<const float4*>(d + 1024), f1 = *reinterpret_cast<const float4*>(d + 1536);
        issue(c + static_cast<long long>(STAGES) * wstride, s);  // same lane re-fills its own 64 bytes: no cross-lane hazard
        stripe_process8<RMW>(mytab, k, q0, q1, f0, f1, lane);
        s = (s + 1 == STAGES) ? 0 : s + 1;
    }
    __syncthreads();
    stripe_fold(tables, warps, k, grad, tid, blockDim.x);
}

// ---- variant S3: striped tables, register prefetch (DEPTH chunks of 256 elements in flight per warp) -----------
template <int DEPTH, bool RMW>
__global__ void __launch_bounds__(128, 1)
    accum_stripe3_kernel(const int32_t* __restrict__ idx, const float* __restrict__ val, long long n_chunks, float* grad, int k) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warps = blockDim.x >> 5;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* tables = reinterpret_cast<float*>(smem_raw);
    const long long gw = static_cast<long long>(blockIdx.x) * warps + warp;
    const long long wstride = static_cast<long long>(gridDim.x) * warps;
    int4 qb[DEPTH][2];
    float4 fb[DEPTH][2];
    auto load = [&](long long c, int4 (&q)[2], float4 (&f)[2]) {
        if (c < n_chunks) {
            const int4* gi = reinterpret_cast<const int4*>(idx + c * kStripeChunk) + lane;
            const float4* gv = reinterpret_cast<const float4*>(val + c * kStripeChunk) + lane;
            q[0] = __ldcs(gi); q[1] = __ldcs(gi + 32);
            f[0] = __ldcs(gv); f[1] = __ldcs(gv + 32);
        }
    };
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) load(gw + d * wstride, qb[d], fb[d]);
    {
        float4* t4 = reinterpret_cast<float4*>(tables);
        const int n4 = warps * k * 4;
        for (int i = tid; i < n4; i += blockDim.x) t4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    float* mytab = tables + static_cast<size_t>(warp) * k * 16 + (lane & 15);
    for (long long c = gw; c < n_chunks; c += wstride * DEPTH) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            const long long cur = c + d * wstride;
            if (cur < n_chunks) {
                const int4 q0 = qb[d][0], q1 = qb[d][1];
                const float4 f0 = fb[d][0], f1 = fb[d][1];
                load(cur + static_cast<long long>(DEPTH) * wstride, qb[d], fb[d]);
                stripe_process8<RMW>(mytab, k, q0, q1, f0, f1, lane);
            }
        }
    }
    __syncthreads();
    stripe_fold(tables, warps, k, grad, tid, blockDim.x);
}


// ---- variant S4: striped tables, exact in-register merge of groups of two batches (no votes, no branches) ------
// Items of a group in canonical order: (batch 0, lane < 16), (batch 0, lane >= 16), (batch 1, lower), (batch 1, upper)
// restricted to one lane pair (L, L ^ 16) -- the only lanes that share table words.  The first occurrence of an id
// absorbs the values of the later ones, so the two updates of a group never touch the same word and issue
// back to back (LDS LDS FADD FADD STS STS).  Timing is independent of the id distribution.
template <bool RMW>
__device__ __forceinline__ void stripe_group2(float* mytab, int k, bool lower, int a0, float v0, int a1, float v1, int p0,
                                              float pv0, int p1, float pv1) {
    const bool e00 = (p0 == a0), e01 = (a1 == a0), e0p1 = (p1 == a0), e1p0 = (p0 == a1), e11 = (p1 == a1);
    const bool ok0 = (static_cast<unsigned>(a0) < static_cast<unsigned>(k)) && (lower || !e00);
    const bool ok1 = (static_cast<unsigned>(a1) < static_cast<unsigned>(k)) && !e01 && !e1p0 && (lower || !e11);
    float acc0 = v0 + ((lower && e00) ? pv0 : 0.f);
    acc0 += e01 ? v1 : 0.f;
    acc0 += e0p1 ? pv1 : 0.f;
    const float acc1 = v1 + ((lower && e11) ? pv1 : 0.f);
    if (RMW) {
        float* s0 = mytab + a0 * 16;
        float* s1 = mytab + a1 * 16;
        float t0 = 0.f, t1 = 0.f;
        if (ok0) t0 = *s0;
        if (ok1) t1 = *s1;
        if (ok0) *s0 = t0 + acc0;
        if (ok1) *s1 = t1 + acc1;
        __syncwarp();
    } else {
        if (acc0 + acc1 == 123.456f && ok0 && ok1) mytab[0] = acc0;
    }
}

// RING slots of 128 elements (512 B ids + 512 B values) per warp, UNITS slots (4 UNITS batches) per iteration
template <int RING, int UNITS, bool RMW>
__global__ void __launch_bounds__(128, 1)
    accum_stripe4_kernel(const int32_t* __restrict__ idx, const float* __restrict__ val, long long n_slots, float* grad, int k) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warps = blockDim.x >> 5;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool lower = lane < 16;
    float* tables = reinterpret_cast<float*>(smem_raw);
    unsigned char* ring = smem_raw + static_cast<size_t>(warps) * k * 64;
    unsigned char* mine = ring + static_cast<size_t>(warp) * RING * 1024 + lane * 16;
    const long long gw = static_cast<long long>(blockIdx.x) * warps + warp;
    const long long wstride = static_cast<long long>(gridDim.x) * warps;
    // warp gw owns slot-groups gw, gw + wstride, ... ; a slot-group = UNITS consecutive slots (contiguous 4 UNITS batches)
    const long long n_groups = n_slots / UNITS;
    auto issue = [&](long long g, int slot0) {
        if (g < n_groups) {
#pragma unroll
            for (int u = 0; u < UNITS; ++u) {
                const long long e = (g * UNITS + u) * 128 + lane * 4;
                int sl = slot0 + u;
                if (sl >= RING) sl -= RING;
                cp_async16(mine + sl * 1024, idx + e);
                cp_async16(mine + sl * 1024 + 512, val + e);
            }
        }
        cp_async_commit();
    };
    constexpr int kGroups = RING / UNITS;  // commit groups in flight
#pragma unroll
    for (int s = 0; s < kGroups; ++s) issue(gw + s * wstride, s * UNITS);
    {
        float4* t4 = reinterpret_cast<float4*>(tables);
        const int n4 = warps * k * 4;
        for (int i = tid; i < n4; i += blockDim.x) t4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    float* mytab = tables + static_cast<size_t>(warp) * k * 16 + (lane & 15);
    int slot0 = 0;
    for (long long g = gw; g < n_groups; g += wstride) {
        cp_async_wait<kGroups - 1>();
        int a[4 * UNITS];
        float v[4 * UNITS];
#pragma unroll
        for (int u = 0; u < UNITS; ++u) {
            int sl = slot0 + u;
            if (sl >= RING) sl -= RING;
            const int4 q = *reinterpret_cast<const int4*>(mine + sl * 1024);
            const float4 f = *reinterpret_cast<const float4*>(mine + sl * 1024 + 512);
            a[4 * u] = q.x; a[4 * u + 1] = q.y; a[4 * u + 2] = q.z; a[4 * u + 3] = q.w;
            v[4 * u] = f.x; v[4 * u + 1] = f.y; v[4 * u + 2] = f.z; v[4 * u + 3] = f.w;
        }
        issue(g + static_cast<long long>(kGroups) * wstride, slot0);
        int p[4 * UNITS];
        float pv[4 * UNITS];
#pragma unroll
        for (int j = 0; j < 4 * UNITS; ++j) {
            p[j] = __shfl_xor_sync(kFull, a[j], 16);
            pv[j] = __shfl_xor_sync(kFull, v[j], 16);
        }
#pragma unroll
        for (int t = 0; t < 2 * UNITS; ++t)
            stripe_group2<RMW>(mytab, k, lower, a[2 * t], v[2 * t], a[2 * t + 1], v[2 * t + 1], p[2 * t], pv[2 * t], p[2 * t + 1],
                               pv[2 * t + 1]);
        slot0 += UNITS;
        if (slot0 >= RING) slot0 -= RING;
    }
    __syncthreads();
    stripe_fold(tables, warps, k, grad, tid, blockDim.x);
}


// ---- variant S6: striped tables shared by a GROUP of warps that take turns (token = mbarrier chain) --------------
// T tables per CTA (16 copies of each bin, copy = lane & 15, one dummy row), GW = warps / T warps per table.  A warp
// loads a unit of 16 batches into registers (prefetch depth 1), resolves duplicates inside groups of two batches
// in registers (stripe_front2: first occurrence absorbs the later ones, disabled items point at the dummy row), waits
// for its table's token, does 8 x (LDS LDS FADD FADD STS STS), passes the token on.  Loads + shuffles + merging of the
// other GW - 1 warps overlap the token holder's read-modify-write chain.
__device__ __forceinline__ void stripe_front2(unsigned base, int k, int lower, int a0, int a1, int p0, int p1, float v0, float v1,
                                              float pv0, float pv1, unsigned& addr0, unsigned& addr1, float& acc0, float& acc1) {
    asm("{\n"
        ".reg .pred lo, e00, e01, e0p1, e1p0, e11, ok0, ok1, t;\n"
        ".reg .s32 s0, s1;\n"
        "setp.ne.s32 lo, %12, 0;\n"
        "setp.eq.s32 e00, %6, %4;\n"
        "setp.eq.s32 e01, %5, %4;\n"
        "setp.eq.s32 e0p1, %7, %4;\n"
        "setp.eq.s32 e1p0, %6, %5;\n"
        "setp.eq.s32 e11, %7, %5;\n"
        "mov.f32 %2, %8;\n"
        "mov.f32 %3, %9;\n"
        "and.pred t, e00, lo;\n"
        "@t add.f32 %2, %2, %10;\n"
        "@e01 add.f32 %2, %2, %9;\n"
        "@e0p1 add.f32 %2, %2, %11;\n"
        "and.pred t, e11, lo;\n"
        "@t add.f32 %3, %3, %11;\n"
        "setp.lt.u32 ok0, %4, %13;\n"
        "not.pred t, e00;\n"
        "or.pred t, t, lo;\n"
        "and.pred ok0, ok0, t;\n"
        "setp.lt.u32 ok1, %5, %13;\n"
        "not.pred t, e11;\n"
        "or.pred t, t, lo;\n"
        "and.pred ok1, ok1, t;\n"
        "not.pred t, e01;\n"
        "and.pred ok1, ok1, t;\n"
        "not.pred t, e1p0;\n"
        "and.pred ok1, ok1, t;\n"
        "selp.s32 s0, %4, %13, ok0;\n"
        "selp.s32 s1, %5, %13, ok1;\n"
        "shl.b32 s0, s0, 6;\n"
        "shl.b32 s1, s1, 6;\n"
        "add.s32 %0, s0, %14;\n"
        "add.s32 %1, s1, %14;\n"
        "}\n"
        : "=r"(addr0), "=r"(addr1), "=f"(acc0), "=f"(acc1)
        : "r"(a0), "r"(a1), "r"(p0), "r"(p1), "f"(v0), "f"(v1), "f"(pv0), "f"(pv1), "r"(lower), "r"(k), "r"(base));
}
__device__ __forceinline__ float lds_f32(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f32(unsigned addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sv::smem_u32(bar)) : "memory");
}

template <bool RMW>
__global__ void __launch_bounds__(512, 1)
    accum_stripe6_kernel(const int32_t* __restrict__ idx, const float* __restrict__ val, long long n_units, float* grad, int k,
                         int T) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warps = blockDim.x >> 5;
    const int GW = warps / T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t = warp % T, j = warp / T;
    const int lower = lane < 16 ? 1 : 0;
    const size_t table_floats = static_cast<size_t>(k + 1) * 16;
    float* tables = reinterpret_cast<float*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(T) * table_floats * 4);  // [T][GW]
    if (tid == 0) {
        for (int i = 0; i < T * GW; ++i) sv::mbar_init(bars + i, 1);
        sv::mbar_fence_init();
        for (int i = 0; i < T; ++i) mbar_arrive(bars + i * GW);  // the first warp of every table starts with the token
    }
    const long long nstreams = static_cast<long long>(gridDim.x) * T;
    const long long stream = static_cast<long long>(blockIdx.x) * T + t;
    int4 nq[4];
    float4 nf[4];
    auto load = [&](long long u) {
        const int4* gi = reinterpret_cast<const int4*>(idx + u * 512) + lane;
        const float4* gv = reinterpret_cast<const float4*>(val + u * 512) + lane;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            nq[m] = __ldcs(gi + 32 * m);
            nf[m] = __ldcs(gv + 32 * m);
        }
    };
    long long u = stream + static_cast<long long>(j) * nstreams;
    const long long ustep = static_cast<long long>(GW) * nstreams;
    if (u < n_units) load(u);
    {
        float4* t4 = reinterpret_cast<float4*>(tables);
        const int n4 = static_cast<int>(T * table_floats / 4);
        for (int i = tid; i < n4; i += blockDim.x) t4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    const unsigned base = sv::smem_u32(tables + static_cast<size_t>(t) * table_floats + (lane & 15));
    uint64_t* my_bar = bars + t * GW + j;
    uint64_t* next_bar = bars + t * GW + (j + 1 == GW ? 0 : j + 1);
    unsigned round = 0;
    for (; u < n_units; u += ustep, ++round) {
        int a[16];
        float v[16];
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            a[4 * m] = nq[m].x; a[4 * m + 1] = nq[m].y; a[4 * m + 2] = nq[m].z; a[4 * m + 3] = nq[m].w;
            v[4 * m] = nf[m].x; v[4 * m + 1] = nf[m].y; v[4 * m + 2] = nf[m].z; v[4 * m + 3] = nf[m].w;
        }
        if (u + ustep < n_units) load(u + ustep);
        unsigned addr[16];
        float acc[16];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const int p0 = __shfl_xor_sync(kFull, a[2 * g], 16), p1 = __shfl_xor_sync(kFull, a[2 * g + 1], 16);
            const float pv0 = __shfl_xor_sync(kFull, v[2 * g], 16), pv1 = __shfl_xor_sync(kFull, v[2 * g + 1], 16);
            stripe_front2(base, k, lower, a[2 * g], a[2 * g + 1], p0, p1, v[2 * g], v[2 * g + 1], pv0, pv1, addr[2 * g],
                          addr[2 * g + 1], acc[2 * g], acc[2 * g + 1]);
        }
        sv::mbar_wait(my_bar, round & 1u);
        if (RMW) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const float t0 = lds_f32(addr[2 * g]);
                const float t1 = lds_f32(addr[2 * g + 1]);
                sts_f32(addr[2 * g], t0 + acc[2 * g]);
                sts_f32(addr[2 * g + 1], t1 + acc[2 * g + 1]);
                __syncwarp();
            }
        } else {
            float sacc = 0.f;
            unsigned sa = 0;
#pragma unroll
            for (int g = 0; g < 16; ++g) { sacc += acc[g]; sa ^= addr[g]; }
            if (sacc == 123.456f && sa == 77u) sts_f32(base, sacc);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(next_bar);
    }
    __syncthreads();
    for (int b = tid; b < k; b += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < T; ++w) {
            const float4* row = reinterpret_cast<const float4*>(tables + static_cast<size_t>(w) * table_floats + static_cast<size_t>(b) * 16);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 x = row[(q + (b >> 1)) & 3];
                s += (x.x + x.y) + (x.z + x.w);
            }
        }
        if (s != 0.f) atomicAdd(grad + b, s);
    }
}

// ---- variant L: load-only (roofline probe: how fast can this grid shape stream idx+val?) ----------------
template <int WARPS, int DEPTH>
__global__ void __launch_bounds__(WARPS * 32, 1)
    stream_only_kernel(const int32_t* __restrict__ idx, const float* __restrict__ val, long long n, float* grad) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long n_chunks = n / 128;
    const long long gw = static_cast<long long>(blockIdx.x) * WARPS + warp;
    const long long wstride = static_cast<long long>(gridDim.x) * WARPS;
    float acc = 0.f;
    int iacc = 0;
    for (long long c = gw; c < n_chunks; c += wstride * DEPTH) {
        int4 q[DEPTH]; float4 f[DEPTH];
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            const long long cc = c + d * wstride;
            if (cc < n_chunks) {
                q[d] = __ldcs(reinterpret_cast<const int4*>(idx + cc * 128 + lane * 4));
                f[d] = __ldcs(reinterpret_cast<const float4*>(val + cc * 128 + lane * 4));
            } else { q[d] = make_int4(0, 0, 0, 0); f[d] = make_float4(0, 0, 0, 0); }
        }
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
            acc += f[d].x + f[d].y + f[d].z + f[d].w;
            iacc ^= q[d].x ^ q[d].y ^ q[d].z ^ q[d].w;
        }
    }
    if (acc == 123.456f && iacc == 77) grad[0] = acc;
}

// ---- host -------------------------------------------------------------------------------------------
static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static inline uint64_t rng64() {
    uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double rngu() { return (rng64() >> 11) * (1.0 / 9007199254740992.0); }

struct Case { const char* name; std::vector<int32_t> idx; std::vector<float> val; std::vector<double> ref; std::vector<double> absref; };

static void make_case(Case& c, const char* name, long long n, int k, int dist) {
    c.name = name;
    c.idx.resize(n);
    c.val.resize(n);
    std::vector<double> cdf(k);
    if (dist == 1) {
        double s = 0;
        for (int i = 0; i < k; ++i) { s += 1.0 / pow(i + 1.0, 1.2); cdf[i] = s; }
        for (int i = 0; i < k; ++i) cdf[i] /= s;
    }
    for (long long i = 0; i < n; ++i) {
        if (dist == 0) c.idx[i] = static_cast<int32_t>(rng64() % k);
        else if (dist == 1) c.idx[i] = static_cast<int32_t>(std::lower_bound(cdf.begin(), cdf.end(), rngu()) - cdf.begin());
        else c.idx[i] = k / 3;
        if (c.idx[i] >= k) c.idx[i] = k - 1;
        c.val[i] = static_cast<float>(rngu() * 2.0 - 1.0);
    }
    c.ref.assign(k, 0.0);
    c.absref.assign(k, 0.0);
    for (long long i = 0; i < n; ++i) { c.ref[c.idx[i]] += c.val[i]; c.absref[c.idx[i]] += fabs(c.val[i]); }
}

typedef void (*LaunchFn)(const int32_t*, const float*, long long, float*, int, int);

template <int WARPS, int DEPTH, int MODE>
static void launch_tag(const int32_t* idx, const float* val, long long n, float* grad, int k, int grid) {
    int kbits = 1;
    while ((1 << kbits) < k) ++kbits;
    const size_t smem = static_cast<size_t>(WARPS) * k * 5;
    auto kern = accum_tag_kernel<WARPS, DEPTH, MODE>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, WARPS * 32, smem>>>(idx, val, n, grad, k, kbits);
}

template <int WARPS, int DEPTH, int HOT, int SAMECHK, bool NOTAG, int MINB>
static void launch_hot(const int32_t* idx, const float* val, long long n, float* grad, int k, int grid) {
    int kbits = 1;
    while ((1 << kbits) < k) ++kbits;
    const size_t smem = static_cast<size_t>(WARPS) * k * 5;
    auto kern = accum_hot_kernel<WARPS, DEPTH, HOT, SAMECHK, NOTAG, MINB>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid * MINB, WARPS * 32, smem>>>(idx, val, n, grad, k, kbits);
}

template <int WARPS, int DEPTH, int HOT, int ADOPT, int MINB>
static void launch_adaptive(const int32_t* idx, const float* val, long long n, float* grad, int k, int grid) {
    const size_t smem = static_cast<size_t>(WARPS) * k * 5;
    auto kern = accum_adaptive_kernel<WARPS, DEPTH, HOT, ADOPT, MINB>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid * MINB, WARPS * 32, smem>>>(idx, val, n, grad, k);
}

static float* g_gtables = nullptr;
template <int WARPS, int GLOBAL_MASK>
static void launch_mixed(const int32_t* idx, const float* val, long long n, float* grad, int k, int grid) {
    if (!g_gtables) CKV(cudaMalloc(&g_gtables, sizeof(float) * 4096 * 1024));
    const size_t smem = static_cast<size_t>(WARPS) * k * 5;
    auto kern = accum_mixed_kernel<WARPS, GLOBAL_MASK>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, WARPS * 32, smem>>>(idx, val, n, grad, k, g_gtables);
}
template <int WARPS, int DEPTH>
static void launch_stream(const int32_t* idx, const float* val, long long n, float* grad, int k, int grid) {
    stream_only_kernel<WARPS, DEPTH><<<grid, WARPS * 32>>>(idx, val, n, grad);
}

template <int WARPS, int STAGES>
static void launch_stripe(const int32_t* idx, const float* val, long long n, float* grad, int k, int grid) {
    const size_t smem = static_cast<size_t>(WARPS) * k * 64 + static_cast<size_t>(WARPS) * STAGES * 2048 + WARPS * STAGES * 8;
    auto kern = accum_stripe_kernel<STAGES>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, WARPS * 32, smem>>>(idx, val, n / kStripeChunk, grad, k);
}

template <int WARPS, int STAGES, bool RMW>
static void launch_stripe2(const int32_t* idx, const float* val, long long n, float* grad, int k, int grid) {
    const size_t smem = static_cast<size_t>(WARPS) * k * 64 + static_cast<size_t>(WARPS) * STAGES * 2048;
    auto kern = accum_stripe2_kernel<STAGES, RMW>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, WARPS * 32, smem>>>(idx, val, n / kStripeChunk, grad, k);
}
template <int WARPS, int DEPTH, bool RMW>
static void launch_stripe3(const int32_t* idx, const float* val, long long n, float* grad, int k, int grid) {
    const size_t smem = static_cast<size_t>(WARPS) * k * 64;
    auto kern = accum_stripe3_kernel<DEPTH, RMW>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, WARPS * 32, smem>>>(idx, val, n / kStripeChunk, grad, k);
}

template <int WARPS, int RING, int UNITS, bool RMW>
static void launch_stripe4(const int32_t* idx, const float* val, long long n, float* grad, int k, int grid) {
    const size_t smem = static_cast<size_t>(WARPS) * k * 64 + static_cast<size_t>(WARPS) * RING * 1024;
    auto kern = accum_stripe4_kernel<RING, UNITS, RMW>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, WARPS * 32, smem>>>(idx, val, n / 128, grad, k);
}

template <int WARPS, int TABLES, bool RMW>
static void launch_stripe6(const int32_t* idx, const float* val, long long n, float* grad, int k, int grid) {
    const size_t smem = static_cast<size_t>(TABLES) * (k + 1) * 64 + WARPS * 8;
    auto kern = accum_stripe6_kernel<RMW>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kern<<<grid, WARPS * 32, smem>>>(idx, val, n / 512, grad, k, TABLES);
}

int main(int argc, char** argv) {
    const long long n = 1LL << (argc > 3 ? atoi(argv[3]) : 24);
    const int k = 1024;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs\n", prop.name, sms);

    Case cases[3];
    make_case(cases[0], "uniform", n, k, 0);
    make_case(cases[1], "zipf", n, k, 1);
    make_case(cases[2], "same", n, k, 2);

    int32_t* d_idx; float* d_val; float* d_grad; unsigned char* d_flush;
    const int mode = argc > 4 ? atoi(argv[4]) : 0;  // 0 write flush, 1 write + read flush, 2 rotate over 8 input copies (no flush)
    const int copies = mode == 2 ? 8 : 1;
    CK(cudaMalloc(&d_idx, n * 4 * copies)); CK(cudaMalloc(&d_val, n * 4 * copies)); CK(cudaMalloc(&d_grad, k * 4));
    const size_t flush_bytes = 256u << 20;
    CK(cudaMalloc(&d_flush, flush_bytes));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));

    struct Variant { const char* name; LaunchFn fn; bool check; };
    Variant variants[] = {
        {"stream32x2", launch_stream<32, 2>, false},
        {"hot0 same0 w32 d1", launch_hot<32, 1, 0, 0, false, 1>, true},
        {"stripe4 w3 r10 u2", launch_stripe4<3, 10, 2, true>, true},
        {"stripe6 w15 t3", launch_stripe6<15, 3, true>, true},
        {"stripe6 w15 t3 normw", launch_stripe6<15, 3, false>, false},
        {"stripe6 w12 t3", launch_stripe6<12, 3, true>, true},
        {"stripe6 w9 t3", launch_stripe6<9, 3, true>, true},
        {"stripe6 w6 t3", launch_stripe6<6, 3, true>, true},
        {"stripe6 w16 t2", launch_stripe6<16, 2, true>, true},
    };
    const char* filter = argc > 1 ? argv[1] : nullptr;
    const int iters = argc > 2 ? atoi(argv[2]) : 13;
    for (Case& c : cases) {
        for (int r = 0; r < copies; ++r) {
            CK(cudaMemcpy(d_idx + r * n, c.idx.data(), n * 4, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(d_val + r * n, c.val.data(), n * 4, cudaMemcpyHostToDevice));
        }
        for (const Variant& v : variants) {
            if (filter && !strstr(v.name, filter)) continue;
            std::vector<float> ts;
            std::vector<float> got(k);
            double worst = 0;
            for (int it = 0; it < iters; ++it) {
                if (mode != 2) CK(cudaMemset(d_flush, it, flush_bytes));
                if (mode == 1)  // read it back: evicts the dirty lines the memset left in L2
                    stream_only_kernel<32, 2><<<sms * 2, 1024>>>(reinterpret_cast<const int32_t*>(d_flush), reinterpret_cast<const float*>(d_flush + flush_bytes / 2), static_cast<long long>(flush_bytes / 8), d_grad);
                CK(cudaMemset(d_grad, 0, k * 4));
                CK(cudaEventRecord(e0));
                v.fn(d_idx + (it % copies) * n, d_val + (it % copies) * n, n, d_grad, k, sms);
                CK(cudaEventRecord(e1));
                CK(cudaDeviceSynchronize());
                CK(cudaGetLastError());
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                if (it >= 3 || iters <= 3) ts.push_back(ms);
                if (it == 0 && v.check) {
                    CK(cudaMemcpy(got.data(), d_grad, k * 4, cudaMemcpyDeviceToHost));
                    for (int b = 0; b < k; ++b) {
                        const double err = fabs(got[b] - c.ref[b]) / (c.absref[b] + 1e-30);
                        worst = std::max(worst, err);
                    }
                }
            }
            std::sort(ts.begin(), ts.end());
            const float med = ts[ts.size() / 2];
            printf("%-8s %-22s median %8.2f us  min %8.2f us  %7.1f GB/s  relerr(sum|x|) %.2e %s\n", c.name, v.name,
                   med * 1e3, ts[0] * 1e3, 8.0 * n / (med * 1e-3) / 1e9, worst,
                   (v.check && worst > 1e-5) ? "FAIL" : "");
        }
    }
    return 0;
}
