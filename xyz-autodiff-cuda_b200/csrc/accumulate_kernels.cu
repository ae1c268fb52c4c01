// csrc/accumulate_kernels.cu -- C2: accumulation of per-element gradients into K shared
// parameters: grad[idx[i]] += val[i].
//
// What it replaces (reference): every thread calling VariableRef::add_grad = atomicAdd on the
// shared parameter's gradient (include/xyz_autodiff/variable.cuh:48-50), i.e. E same-address
// L2 atomics (tests/test_parallel_gradient_accumulation.cu:25-49; on __shared__ memory:
// tests/test_shared_memory_atomic.cu:29-64).
//
// Three kernels, picked by accumulate() below (no floating-point atomics on shared memory anywhere -- fp32 shared
// atomics are CAS loops on sm_100a, ATOMS.CAST.SPIN):
//   1. accumulate_striped_kernel   fp32 (accumulate_striped64_kernel: fp64), 16-byte aligned arrays, n >= 2^16, K <= 3626
//                                  in one pass and K <= 16384 in passes of 1810 bins: lane-striped tables shared by
//                                  turn-taking warps, exact in-register duplicate merge; same cost for every id
//                                  distribution; deterministic per CTA by construction (the hot path, see below);
//   2. accumulate_tagged_kernel    fp64 beyond its striped range: one value table + one byte TAG table per warp,
//                                  arbitration by "store my lane id, read it back";
//   3. accumulate_kernel           small n / unaligned bases: per-warp tables, duplicates grouped with match.any and
//                                  folded by a fixed-order shuffle tree;
//   (K too large for shared memory: accumulate_global_kernel, warp-aggregated global REDs.)
// Every table kernel ends the same way: the CTA folds its tables in a fixed order, then either issues one RED per
// bin (fast) or writes a partial row that a finishing kernel adds in CTA order (XYZ_FLAG_DETERMINISTIC, and the
// multi-GPU path whose finishing kernel also exchanges the row over NVLink mailboxes).
// 8 algorithmic bytes per element (4 with implicit ids) -> HBM bound once contention is gone.
#include "common.cuh"

namespace xyzb {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kCtasPerSM = 4;
constexpr unsigned kFull = 0xffffffffu;

// Fold the values of all lanes in `peers` (lanes holding the same bin) onto the lowest lane of the
// group.  Fixed pairing order: deterministic.  Every lane of the warp must call it.
template <class T>
__device__ __forceinline__ T reduce_peers(unsigned peers, T x, int lane) {
    int rel = __popc(peers & ((1u << lane) - 1u));   // my rank inside the group
    unsigned above = peers & ~((2u << lane) - 1u);   // group members in higher lanes
    while (__any_sync(kFull, above != 0u)) {
        const int next = __ffs(above);               // 1-based lane of the next higher member, 0 = none
        const T t = __shfl_sync(kFull, x, (next - 1) & 31);
        if (next) x += t;
        // members with an odd rank have been absorbed by their lower neighbour: drop them everywhere
        const unsigned alive = __ballot_sync(kFull, (rel & 1) == 0);
        above &= alive;
        rel >>= 1;
    }
    return x;
}

template <class T>
__device__ __forceinline__ void warp_batch(T* table, int id, T v, int k, int lane) {
    const bool valid = (id >= 0) && (id < k);
    const int key = valid ? id : (-1 - lane);        // invalid lanes never match anybody
    const unsigned peers = __match_any_sync(kFull, key);
    if (peers == kFull) {                            // whole batch on one bin (the reference's own pattern)
        const T s = warp_sum(v);
        if (lane == 0 && valid) table[id] += s;
    } else {
        const bool alone = (peers == (1u << lane));
        if (!__all_sync(kFull, alone)) v = reduce_peers(peers, v, lane);
        if (valid && lane == __ffs(peers) - 1) table[id] += v;
    }
    __syncwarp();
}

template <class T> struct Vec4;
template <> struct Vec4<float> { using type = float4; };
template <> struct Vec4<double> { using type = double4; };

// tables: kWarps x k of T in dynamic shared memory.
// kVec: idx/val 16-byte aligned (vector loads).  kImplicit: id = i mod k, idx ignored.
template <class T, bool kVec, bool kImplicit, bool kDeterministic>
__global__ void __launch_bounds__(kThreads, kCtasPerSM)
    accumulate_kernel(const int32_t* __restrict__ idx, const T* __restrict__ val, long long n, T* grad, int k,
                      T* partial_rows) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* tables = reinterpret_cast<T*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kWarps * k; i += kThreads) tables[i] = T(0);
    __syncthreads();
    T* table = tables + static_cast<size_t>(warp) * k;

    const long long n_chunks = (n + 127) / 128;
    const long long gw = static_cast<long long>(blockIdx.x) * kWarps + warp;
    const long long wstride = static_cast<long long>(gridDim.x) * kWarps;
    for (long long c = gw; c < n_chunks; c += wstride) {
        const long long e0 = c * 128 + lane * 4;
        int id[4];
        T v[4];
        if (kVec && e0 + 4 <= n) {
            if constexpr (kImplicit) {
#pragma unroll
                for (int j = 0; j < 4; ++j) id[j] = static_cast<int>((e0 + j) % k);
            } else {
                const int4 q = __ldcs(reinterpret_cast<const int4*>(idx + e0));
                id[0] = q.x; id[1] = q.y; id[2] = q.z; id[3] = q.w;
            }
            if constexpr (sizeof(T) == 4) {
                const float4 f = __ldcs(reinterpret_cast<const float4*>(val + e0));
                v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
            } else {
                const double2 f0 = __ldcs(reinterpret_cast<const double2*>(val + e0));
                const double2 f1 = __ldcs(reinterpret_cast<const double2*>(val + e0 + 2));
                v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long e = e0 + j;
                if (e < n) {
                    id[j] = kImplicit ? static_cast<int>(e % k) : __ldg(idx + e);
                    v[j] = __ldg(val + e);
                } else {
                    id[j] = -1;
                    v[j] = T(0);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) warp_batch<T>(table, id[j], v[j], k, lane);
    }
    __syncthreads();

    // CTA: add the warps' tables in warp order, 4 bins per thread-step
    if constexpr (kDeterministic) {
        T* row = partial_rows + static_cast<size_t>(blockIdx.x) * k;
        for (int b = tid; b < k; b += kThreads) {
            T s = T(0);
#pragma unroll
            for (int w = 0; w < kWarps; ++w) s += tables[static_cast<size_t>(w) * k + b];
            row[b] = s;
        }
    } else {
        if constexpr (sizeof(T) == 4) {
            const bool v4 = ((k & 3) == 0) && ((reinterpret_cast<uintptr_t>(grad) & 15u) == 0);
            if (v4) {
                for (int b = tid * 4; b < k; b += kThreads * 4) {
                    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int w = 0; w < kWarps; ++w) {
                        const float4 t = *reinterpret_cast<const float4*>(tables + static_cast<size_t>(w) * k + b);
                        s[0] += t.x; s[1] += t.y; s[2] += t.z; s[3] += t.w;
                    }
                    if (s[0] != 0.f || s[1] != 0.f || s[2] != 0.f || s[3] != 0.f)
                        red_add_v4(reinterpret_cast<float*>(grad) + b, s[0], s[1], s[2], s[3]);
                }
                return;
            }
        }
        for (int b = tid; b < k; b += kThreads) {
            T s = T(0);
#pragma unroll
            for (int w = 0; w < kWarps; ++w) s += tables[static_cast<size_t>(w) * k + b];
            if (s != T(0)) atomicAdd(grad + b, s);
        }
    }
}


// ---- tagged tables: the fast path ------------------------------------------------------------------------
// ncu on the kernel above: `match.any` is its top stall when ids are spread, and a design with R replica
// columns per bin (collision check by shuffles) leaves room for only 6 warps per SM -- latency bound (102 us
// for 2^24 uniform ids).  The hot kernel therefore arbitrates through shared memory itself:
//   * every warp owns a K-bin value table plus a K-byte TAG table; 32 warps per SM (160 KB for K = 1024);
//   * a batch of 32 elements: every lane stores its lane id to tag[id], reads it back, and the lane whose
//     store landed ("winner", exactly one per distinct id) does the plain LDS/FADD/STS on the value table;
//   * losers (duplicates inside the batch: 38 % of uniform batches have one) are replayed one lane at a time
//     when there are <= 2 of them, else grouped with `match.any` over the losers only (few distinct keys =>
//     cheap), folded with the fixed-order shuffle tree and added by each group's lowest lane;
//   * when half a batch or more lost (the reference's own all-threads-one-parameter pattern) the warp turns on
//     a register-level check "whole batch on one bin" (one shuffle + vote, then a shuffle-tree sum and ONE
//     table update); it turns itself off at the first batch that fails the check.
// Which duplicate wins the tag store is the hardware's choice; the kDet flavour (XYZ_FLAG_DETERMINISTIC) therefore
// never lets the winner add first: every duplicate group is folded in lane order and added once (see below).
// ncu (profiles/): 15.4 shared-memory wavefronts per batch, 14 of them the 4 table/tag accesses at the
// ~3.5-way bank conflict degree of 32 random banks; l1tex 93 % busy -- the kernel sits on the shared-memory
// pipe, 41 us for 2^24 uniform ids against 29 us for a kernel that only streams idx+val with this grid.
template <class T, int kWarpsT, bool kImplicit, bool kRows, bool kDet>
__global__ void __launch_bounds__(kWarpsT * 32, 1)
    accumulate_tagged_kernel(const int32_t* __restrict__ idx, const T* __restrict__ val, long long n, T* grad, int k,
                             T* partial_rows, int aligned) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* tables = reinterpret_cast<T*>(smem_raw);
    uint8_t* tags = reinterpret_cast<uint8_t*>(tables + static_cast<size_t>(kWarpsT) * k);
    constexpr int kThreadsT = kWarpsT * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kWarpsT * k; i += kThreadsT) tables[i] = T(0);
    __syncthreads();
    T* table = tables + static_cast<size_t>(warp) * k;
    uint8_t* tag = tags + static_cast<size_t>(warp) * k;

    const long long n_chunks = (n + 127) / 128;
    const long long gw = static_cast<long long>(blockIdx.x) * kWarpsT + warp;
    const long long wstride = static_cast<long long>(gridDim.x) * kWarpsT;
    bool same_mode = false;  // warp-uniform

    auto batch = [&](int id, T v) {
        const bool valid = static_cast<unsigned>(id) < static_cast<unsigned>(k);
        if (same_mode) {
            const int id0 = __shfl_sync(kFull, id, 0);
            if (__all_sync(kFull, id == id0)) {
                const T s = warp_sum(v);
                if (lane == 0 && valid) table[id] += s;
                __syncwarp();
                return;
            }
            same_mode = false;
        }
        if (valid) tag[id] = static_cast<uint8_t>(lane);
        __syncwarp();
        if constexpr (kDet) {
            // Deterministic flavour: WHICH duplicate's tag store landed must not matter.  rep = the lane whose store
            // landed (the same value for every member of a duplicate group); groups with losers are folded in lane
            // order by the fixed shuffle tree and added once by their lowest lane; everybody else adds directly.
            const unsigned int rep = valid ? tag[id] : static_cast<unsigned int>(lane);
            const bool mine = valid && rep != static_cast<unsigned int>(lane);
            const unsigned lost = __ballot_sync(kFull, mine);
            if (lost == 0u) {
                if (valid) table[id] += v;
                __syncwarp();
                return;
            }
            same_mode = __popc(lost) >= 16;
            const unsigned repmask = __reduce_or_sync(kFull, mine ? (1u << rep) : 0u);
            const bool ingroup = mine || ((repmask >> lane) & 1u);
            if (valid && !ingroup) table[id] += v;
            unsigned peers = __match_any_sync(kFull, ingroup ? static_cast<int>(rep) : 64);
            if (!ingroup) peers = 1u << lane;
            v = reduce_peers(peers, v, lane);
            if (ingroup && lane == __ffs(peers) - 1) table[id] += v;
            __syncwarp();
            return;
        }
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
        same_mode = __popc(lost) >= 16;
        unsigned peers = __match_any_sync(kFull, mine ? id : -1);
        if (!mine) peers = 1u << lane;
        v = reduce_peers(peers, v, lane);
        if (mine && lane == __ffs(peers) - 1) table[id] += v;
        __syncwarp();
    };

    int id_nxt[4];
    T v_nxt[4];
    auto load_chunk = [&](long long c, int (&id)[4], T (&v)[4]) {
        const long long e0 = c * 128 + lane * 4;
        if (aligned && c < n_chunks && e0 + 4 <= n) {  // 16-byte aligned bases: one vector load per array
            if constexpr (kImplicit) {
#pragma unroll
                for (int j = 0; j < 4; ++j) id[j] = static_cast<int>((e0 + j) % k);
            } else {
                const int4 q = __ldcs(reinterpret_cast<const int4*>(idx + e0));
                id[0] = q.x; id[1] = q.y; id[2] = q.z; id[3] = q.w;
            }
            if constexpr (sizeof(T) == 4) {
                const float4 f = __ldcs(reinterpret_cast<const float4*>(val + e0));
                v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
            } else {
                const double2 f0 = __ldcs(reinterpret_cast<const double2*>(val + e0));
                const double2 f1 = __ldcs(reinterpret_cast<const double2*>(val + e0 + 2));
                v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long e = e0 + j;
                const bool in = (c < n_chunks) && (e < n);
                id[j] = in ? (kImplicit ? static_cast<int>(e % k) : __ldg(idx + e)) : -1;
                v[j] = in ? __ldg(val + e) : T(0);
            }
        }
    };

    // one chunk (1 KB of idx+val per warp) in flight per warp while the previous one is accumulated:
    // 32 warps x 1 KB per SM covers the HBM latency (a second chunk in flight measured slower)
    load_chunk(gw, id_nxt, v_nxt);
    for (long long c = gw; c < n_chunks; c += wstride) {
        int id_cur[4];
        T v_cur[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            id_cur[j] = id_nxt[j];
            v_cur[j] = v_nxt[j];
        }
        load_chunk(c + wstride, id_nxt, v_nxt);
#pragma unroll
        for (int j = 0; j < 4; ++j) batch(id_cur[j], v_cur[j]);
    }
    __syncthreads();

    // fold the warps' tables in warp order
    for (int b = tid; b < k; b += kThreadsT) {
        T s = T(0);
#pragma unroll 8
        for (int w = 0; w < kWarpsT; ++w) s += tables[static_cast<size_t>(w) * k + b];
        if constexpr (kRows) {
            partial_rows[static_cast<size_t>(blockIdx.x) * k + b] = s;  // summed (and exchanged) by the finishing kernel
        } else {
            if (s != T(0)) atomicAdd(grad + b, s);
        }
    }
}


// ---- lane-striped tables shared by warps that take turns: the fp32 fast path ------------------------------------
// ncu on the tagged kernel above: 15.4 shared-memory wavefronts per batch of 32 elements, 14 of them the tag / table
// accesses at the ~3.5-way conflict degree of 32 random banks; 43 us for 2^24 uniform ids, 76 us for Zipf ids.
// This kernel removes both the tags and the random bank pattern:
//   * a table holds 16 COPIES of every bin, copy = lane & 15, word (id * 16 + copy): lanes L and L ^ 16 are the only
//     two lanes of a warp that can touch the same word, and the 16 lanes of a half-warp always hit 16 different
//     banks => an update costs 2 + 2 wavefronts, and duplicates are found with ONE shuffle of the id;
//   * 64 B per bin => T = 3 tables per SM at K = 1024.  A table is shared by a GROUP of GW = 12 / T warps that take
//     turns (token = a ring of named barriers: the predecessor does bar.arrive, the successor bar.sync).  While one warp holds the token the others
//     load their next unit (16 batches = 512 elements, straight into registers, prefetch depth 1) and resolve
//     duplicates in registers;
//   * duplicates are resolved exactly inside groups of two batches (stripe_front2: per lane pair the first
//     occurrence of an id absorbs the later ones in a fixed order, absorbed items are pointed at a dummy row), so a
//     group's two updates never touch the same word and issue back to back: LDS LDS FADD FADD STS STS, no votes, no
//     branches, no match.any.  The cost is the same for every id distribution;
//   * element -> (CTA, table, turn) is static and every sum has a fixed order, so per-CTA rows are bit-identical run
//     to run: XYZ_FLAG_DETERMINISTIC only swaps the final REDs for rows + the finishing kernel.
// Measured (dev/accum_lab.cu, B200): 30.9 us for 2^24 uniform, Zipf(1.2) and all-equal ids alike = the time of a
// kernel that only streams idx + val with a one-CTA-per-SM grid.
constexpr int kStripeWarps = 12;
constexpr int kStripeUnit = 512;  // elements per warp turn: 16 batches of 32

__device__ __forceinline__ void stripe_front2(unsigned base, int k, int dummy_row, int lower, int a0, int a1, int p0, int p1,
                                              float v0, float v1, float pv0, float pv1, unsigned& addr0, unsigned& addr1,
                                              float& acc0, float& acc1) {
    // items in canonical order: (batch 0, lower lane) (batch 0, upper lane) (batch 1, lower) (batch 1, upper);
    // a* / v* are this lane's, p* / pv* the partner lane's (lane ^ 16).
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
        "setp.lt.u32 ok0, %4, %13;\n"   // valid id and (lower lane or the partner's batch-0 id differs)
        "not.pred t, e00;\n"
        "or.pred t, t, lo;\n"
        "and.pred ok0, ok0, t;\n"
        "setp.lt.u32 ok1, %5, %13;\n"   // valid, not absorbed by either batch-0 item, and first of batch 1
        "not.pred t, e11;\n"
        "or.pred t, t, lo;\n"
        "and.pred ok1, ok1, t;\n"
        "not.pred t, e01;\n"
        "and.pred ok1, ok1, t;\n"
        "not.pred t, e1p0;\n"
        "and.pred ok1, ok1, t;\n"
        "selp.s32 s0, %4, %15, ok0;\n"  // absorbed / invalid items go to this half-warp's dummy row
        "selp.s32 s1, %5, %15, ok1;\n"
        "shl.b32 s0, s0, 6;\n"
        "shl.b32 s1, s1, 6;\n"
        "add.s32 %0, s0, %14;\n"
        "add.s32 %1, s1, %14;\n"
        "}\n"
        : "=r"(addr0), "=r"(addr1), "=f"(acc0), "=f"(acc1)
        : "r"(a0), "r"(a1), "r"(p0), "r"(p1), "f"(v0), "f"(v1), "f"(pv0), "f"(pv1), "r"(lower), "r"(k), "r"(base),
          "r"(dummy_row));
}
// Named barriers (bar.sync / bar.arrive with an id and a thread count) as the turn token between two warps: the
// predecessor ARRIVES (does not wait), the successor SYNCs; 64 = both warps.  compute-sanitizer's racecheck tracks
// these, unlike hand-rolled mbarrier polling.
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int threads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ float lds_f32(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f32(unsigned addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

template <bool kImplicit, bool kRows>
__global__ void __launch_bounds__(kStripeWarps * 32, 1)
    accumulate_striped_kernel(const int32_t* __restrict__ idx, const float* __restrict__ val, long long n, float* grad,
                              int k, float* partial_rows, int T, int id_base, int k_ids) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int kThreadsS = kStripeWarps * 32;
    const int GW = kStripeWarps / T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t = warp % T, j = warp / T;
    const int lower = lane < 16 ? 1 : 0;
    // rows k and k + 1 = dummies, one per half-warp (lanes L and L ^ 16 share a copy, so they must not share a dummy)
    const size_t table_floats = static_cast<size_t>(k + 2) * 16;
    float* tables = reinterpret_cast<float*>(smem_raw);
    const int dummy_row = k + (lane >> 4);
    const long long n_units = (n + kStripeUnit - 1) / kStripeUnit;
    const long long nstreams = static_cast<long long>(gridDim.x) * T;
    const long long stream = static_cast<long long>(blockIdx.x) * T + t;
    int na[16];
    float nv[16];
    auto load = [&](long long u) {
        const long long e0 = u * kStripeUnit + lane * 4;
        if ((u + 1) * kStripeUnit <= n) {
            const float4* gv = reinterpret_cast<const float4*>(val + e0);
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const float4 f = __ldcs(gv + 32 * m);
                nv[4 * m] = f.x; nv[4 * m + 1] = f.y; nv[4 * m + 2] = f.z; nv[4 * m + 3] = f.w;
            }
            if constexpr (kImplicit) {
                const unsigned uk = static_cast<unsigned>(k_ids);  // implicit id = element index mod the TOTAL bin count
                const unsigned r0 = static_cast<unsigned>(static_cast<unsigned long long>(e0) % uk);
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    unsigned r = (r0 + 128u * m) % uk;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        na[4 * m + c] = static_cast<int>(r);
                        r = (r + 1u == uk) ? 0u : r + 1u;
                    }
                }
            } else {
                const int4* gi = reinterpret_cast<const int4*>(idx + e0);
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    const int4 q = __ldcs(gi + 32 * m);
                    na[4 * m] = q.x; na[4 * m + 1] = q.y; na[4 * m + 2] = q.z; na[4 * m + 3] = q.w;
                }
            }
        } else {  // the last, partial unit
#pragma unroll
            for (int m = 0; m < 4; ++m) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const long long e = e0 + 128 * m + c;
                    const bool in = e < n;
                    na[4 * m + c] = in ? (kImplicit ? static_cast<int>(e % k_ids) : __ldg(idx + e)) : -1;
                    nv[4 * m + c] = in ? __ldg(val + e) : 0.f;
                }
            }
        }
    };
    long long u = stream + static_cast<long long>(j) * nstreams;
    const long long ustep = static_cast<long long>(GW) * nstreams;
    if (u < n_units) load(u);
    {
        float4* t4 = reinterpret_cast<float4*>(tables);
        const int n4 = static_cast<int>(T * table_floats / 4);
        for (int i = tid; i < n4; i += kThreadsS) t4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    const unsigned base = smem_u32(tables + static_cast<size_t>(t) * table_floats + (lane & 15));
    // turn token of table t: named barrier 1 + t * GW + j belongs to warp j of the group (ids 1 .. 12; 0 is
    // __syncthreads).  Warp j's predecessor arrives on it when its turn ends; warp 0 owns the first turn.
    const int my_bar = 1 + t * GW + j;
    const int next_bar = 1 + t * GW + (j + 1 == GW ? 0 : j + 1);
    unsigned round = 0;
    for (; u < n_units; u += ustep, ++round) {
        int a[16];
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            a[i] = na[i] - id_base;  // bins [id_base, id_base + k) of this pass; everything else fails the range check
            v[i] = nv[i];
        }
        if (u + ustep < n_units) load(u + ustep);
        unsigned addr[16];
        float acc[16];
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const int p0 = __shfl_xor_sync(kFull, a[2 * g], 16), p1 = __shfl_xor_sync(kFull, a[2 * g + 1], 16);
            const float pv0 = __shfl_xor_sync(kFull, v[2 * g], 16), pv1 = __shfl_xor_sync(kFull, v[2 * g + 1], 16);
            stripe_front2(base, k, dummy_row, lower, a[2 * g], a[2 * g + 1], p0, p1, v[2 * g], v[2 * g + 1], pv0, pv1, addr[2 * g],
                          addr[2 * g + 1], acc[2 * g], acc[2 * g + 1]);
        }
        if (GW > 1 && (round | static_cast<unsigned>(j)) != 0u) named_bar_sync(my_bar, 64);  // my turn on table t
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const float t0 = lds_f32(addr[2 * g]);
            const float t1 = lds_f32(addr[2 * g + 1]);
            sts_f32(addr[2 * g], t0 + acc[2 * g]);
            sts_f32(addr[2 * g + 1], t1 + acc[2 * g + 1]);
            __syncwarp();  // lane L ^ 16 may read these words in the next group
        }
        if (GW > 1) named_bar_arrive(next_bar, 64);  // hand the table on (all 32 lanes arrive; nobody waits here)
    }
    __syncthreads();
    // fold: bin b = 16 copies x T tables, fixed order; the float4 order is rotated by b / 2 so that a quarter-warp
    // reads eight different bank groups
    for (int b = tid; b < k; b += kThreadsS) {
        float s = 0.f;
        for (int w = 0; w < T; ++w) {
            const float4* row =
                reinterpret_cast<const float4*>(tables + static_cast<size_t>(w) * table_floats + static_cast<size_t>(b) * 16);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 x = row[(q + (b >> 1)) & 3];
                s += (x.x + x.y) + (x.z + x.w);
            }
        }
        if constexpr (kRows) {
            partial_rows[static_cast<size_t>(blockIdx.x) * k + b] = s;
        } else {
            if (s != 0.f) atomicAdd(grad + b, s);
        }
    }
}

// number of striped tables that fit one SM for K bins (a divisor of kStripeWarps; 1 table up to K = 3626: 57 us for 2^24
// elements, the 12 warps take turns on it), 0 = does not fit
inline int stripe_tables(int k) {
    const size_t budget = 227 * 1024 - 256;
    const size_t per_table = (static_cast<size_t>(k) + 2) * 64;
    const int choices[6] = {12, 6, 4, 3, 2, 1};
    for (int c : choices)
        if (per_table * c <= budget) return c;
    return 0;
}

// One pass over the elements for the bins [id_base, id_base + k) (k small enough for stripe_tables(k) >= 2).
// rows == nullptr: REDs into grad + id_base; else one row of k sums per CTA (returns the number of rows in *n_rows).
template <bool kImplicit>
int launch_striped(const int32_t* idx, const float* val, long long n, float* grad, int k, cudaStream_t st, float* rows,
                   int* n_rows, int id_base = 0, int k_total = 0) {
    const int k_ids = k_total > 0 ? k_total : k;
    const int T = stripe_tables(k);
    const size_t smem = static_cast<size_t>(T) * (k + 2) * 64;
    const long long n_units = (n + kStripeUnit - 1) / kStripeUnit;
    const long long want = (n_units + kStripeWarps - 1) / kStripeWarps;
    const int sms = sm_count();
    const int grid = static_cast<int>(want < sms ? want : sms);
    if (rows) {
        auto kern = accumulate_striped_kernel<kImplicit, true>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        kern<<<grid, kStripeWarps * 32, smem, st>>>(idx, val, n, grad ? grad + id_base : nullptr, k, rows, T, id_base, k_ids);
    } else {
        auto kern = accumulate_striped_kernel<kImplicit, false>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        kern<<<grid, kStripeWarps * 32, smem, st>>>(idx, val, n, grad + id_base, k, nullptr, T, id_base, k_ids);
    }
    count_launch();
    if (n_rows) *n_rows = grid;
    return last_error();
}

// Bins per pass when K does not fit two striped tables: the largest count that does (1810), and the K up to which
// re-reading the elements once per 1810 bins beats the alternatives on their WORST case (measured, 2^24 elements:
// K = 8192 in 5 passes = 155 us for any distribution; the tagged tables need 88 / 201 us at K = 4096 (uniform / Zipf),
// plain global atomics 233 us / 5.9 ms at K = 8192).
constexpr int kStripePassBins = 1810;
constexpr int kStripeMultiPassMaxK = 16384;

// ---- fp64 flavour of the striped tables ----------------------------------------------------------------------------
// Same design with 8-byte words: a bin row is still 64 bytes = 8 COPIES (copy = lane & 7), so T = 3 tables still fit at
// K = 1024.  Now four lanes (L, L ^ 8, L ^ 16, L ^ 24) share a word.  The warp is treated as two half-warps that use
// the table one after the other: inside a half only L and L ^ 8 collide -- the pair structure of the fp32 kernel with
// partner lane ^ 8 -- and the two halves are ordered by a __syncwarp between their update phases.  Duplicate merging
// (stripe_front2_f64) therefore runs for all 32 lanes at once, and a group's update is
// [LDS LDS DADD DADD STS STS](lanes 0-15) [the same](lanes 16-31).  Units are 8 batches (256 elements) to keep the
// register footprint of the 64-bit values at the fp32 kernel's level.
#ifndef XYZ_STRIPE64_BATCHES
#define XYZ_STRIPE64_BATCHES 8
#endif
constexpr int kB64 = XYZ_STRIPE64_BATCHES;   // batches of 32 elements per warp turn
constexpr int kStripeUnit64 = 32 * kB64;

__device__ __forceinline__ void stripe_front2_f64(unsigned base, int k, int dummy_row, int lower, int a0, int a1, int p0,
                                                  int p1, double v0, double v1, double pv0, double pv1, unsigned& addr0,
                                                  unsigned& addr1, double& acc0, double& acc1) {
    asm("{\n"
        ".reg .pred lo, e00, e01, e0p1, e1p0, e11, ok0, ok1, t;\n"
        ".reg .s32 s0, s1;\n"
        "setp.ne.s32 lo, %12, 0;\n"
        "setp.eq.s32 e00, %6, %4;\n"
        "setp.eq.s32 e01, %5, %4;\n"
        "setp.eq.s32 e0p1, %7, %4;\n"
        "setp.eq.s32 e1p0, %6, %5;\n"
        "setp.eq.s32 e11, %7, %5;\n"
        "mov.f64 %2, %8;\n"
        "mov.f64 %3, %9;\n"
        "and.pred t, e00, lo;\n"
        "@t add.f64 %2, %2, %10;\n"
        "@e01 add.f64 %2, %2, %9;\n"
        "@e0p1 add.f64 %2, %2, %11;\n"
        "and.pred t, e11, lo;\n"
        "@t add.f64 %3, %3, %11;\n"
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
        "selp.s32 s0, %4, %15, ok0;\n"
        "selp.s32 s1, %5, %15, ok1;\n"
        "shl.b32 s0, s0, 6;\n"
        "shl.b32 s1, s1, 6;\n"
        "add.s32 %0, s0, %14;\n"
        "add.s32 %1, s1, %14;\n"
        "}\n"
        : "=r"(addr0), "=r"(addr1), "=d"(acc0), "=d"(acc1)
        : "r"(a0), "r"(a1), "r"(p0), "r"(p1), "d"(v0), "d"(v1), "d"(pv0), "d"(pv1), "r"(lower), "r"(k), "r"(base),
          "r"(dummy_row));
}
__device__ __forceinline__ double lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64(unsigned addr, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}

template <bool kImplicit, bool kRows>
__global__ void __launch_bounds__(kStripeWarps * 32, 1)
    accumulate_striped64_kernel(const int32_t* __restrict__ idx, const double* __restrict__ val, long long n, double* grad,
                                int k, double* partial_rows, int T, int id_base, int k_ids) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int kThreadsS = kStripeWarps * 32;
    const int GW = kStripeWarps / T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t = warp % T, j = warp / T;
    const int lower = (lane & 8) ? 0 : 1;   // the lower lane of the pair (L, L ^ 8)
    const bool first_half = lane < 16;
    // rows k .. k + 3 = dummies: one per (half-warp, pair side), so no two lanes ever share a dummy word
    const size_t table_doubles = static_cast<size_t>(k + 4) * 8;
    double* tables = reinterpret_cast<double*>(smem_raw);
    const int dummy_row = k + (lane >> 3);
    const long long n_units = (n + kStripeUnit64 - 1) / kStripeUnit64;
    const long long nstreams = static_cast<long long>(gridDim.x) * T;
    const long long stream = static_cast<long long>(blockIdx.x) * T + t;
    int na[kB64];
    double nv[kB64];
    auto load = [&](long long u) {
        const long long e0 = u * kStripeUnit64 + lane * 4;
        if ((u + 1) * kStripeUnit64 <= n) {
#pragma unroll
            for (int m = 0; m < kB64 / 4; ++m) {
                const double2* gv = reinterpret_cast<const double2*>(val + e0 + 128 * m);
                const double2 f0 = __ldcs(gv), f1 = __ldcs(gv + 1);
                nv[4 * m] = f0.x; nv[4 * m + 1] = f0.y; nv[4 * m + 2] = f1.x; nv[4 * m + 3] = f1.y;
            }
            if constexpr (kImplicit) {
                const unsigned uk = static_cast<unsigned>(k_ids);
                const unsigned r0 = static_cast<unsigned>(static_cast<unsigned long long>(e0) % uk);
#pragma unroll
                for (int m = 0; m < kB64 / 4; ++m) {
                    unsigned r = (r0 + 128u * m) % uk;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        na[4 * m + c] = static_cast<int>(r);
                        r = (r + 1u == uk) ? 0u : r + 1u;
                    }
                }
            } else {
#pragma unroll
                for (int m = 0; m < kB64 / 4; ++m) {
                    const int4 q = __ldcs(reinterpret_cast<const int4*>(idx + e0 + 128 * m));
                    na[4 * m] = q.x; na[4 * m + 1] = q.y; na[4 * m + 2] = q.z; na[4 * m + 3] = q.w;
                }
            }
        } else {
#pragma unroll
            for (int m = 0; m < kB64 / 4; ++m) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const long long e = e0 + 128 * m + c;
                    const bool in = e < n;
                    na[4 * m + c] = in ? (kImplicit ? static_cast<int>(e % k_ids) : __ldg(idx + e)) : -1;
                    nv[4 * m + c] = in ? __ldg(val + e) : 0.0;
                }
            }
        }
    };
    long long u = stream + static_cast<long long>(j) * nstreams;
    const long long ustep = static_cast<long long>(GW) * nstreams;
    if (u < n_units) load(u);
    {
        float4* t4 = reinterpret_cast<float4*>(tables);
        const int n4 = static_cast<int>(T * table_doubles / 2);
        for (int i = tid; i < n4; i += kThreadsS) t4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    const unsigned base = smem_u32(tables + static_cast<size_t>(t) * table_doubles + (lane & 7));
    const int my_bar = 1 + t * GW + j;
    const int next_bar = 1 + t * GW + (j + 1 == GW ? 0 : j + 1);
    unsigned round = 0;
    for (; u < n_units; u += ustep, ++round) {
        int a[kB64];
        double v[kB64];
#pragma unroll
        for (int i = 0; i < kB64; ++i) {
            a[i] = na[i] - id_base;
            v[i] = nv[i];
        }
        if (u + ustep < n_units) load(u + ustep);
        unsigned addr[kB64];
        double acc[kB64];
#pragma unroll
        for (int g = 0; g < kB64 / 2; ++g) {
            const int p0 = __shfl_xor_sync(kFull, a[2 * g], 8), p1 = __shfl_xor_sync(kFull, a[2 * g + 1], 8);
            const double pv0 = __shfl_xor_sync(kFull, v[2 * g], 8), pv1 = __shfl_xor_sync(kFull, v[2 * g + 1], 8);
            stripe_front2_f64(base, k, dummy_row, lower, a[2 * g], a[2 * g + 1], p0, p1, v[2 * g], v[2 * g + 1], pv0, pv1,
                              addr[2 * g], addr[2 * g + 1], acc[2 * g], acc[2 * g + 1]);
        }
        if (GW > 1 && (round | static_cast<unsigned>(j)) != 0u) named_bar_sync(my_bar, 64);
#pragma unroll
        for (int g = 0; g < kB64 / 2; ++g) {
            if (first_half) {
                const double t0 = lds_f64(addr[2 * g]);
                const double t1 = lds_f64(addr[2 * g + 1]);
                sts_f64(addr[2 * g], t0 + acc[2 * g]);
                sts_f64(addr[2 * g + 1], t1 + acc[2 * g + 1]);
            }
            __syncwarp();  // lanes 16-31 may touch the words lanes 0-15 just wrote
            if (!first_half) {
                const double t0 = lds_f64(addr[2 * g]);
                const double t1 = lds_f64(addr[2 * g + 1]);
                sts_f64(addr[2 * g], t0 + acc[2 * g]);
                sts_f64(addr[2 * g + 1], t1 + acc[2 * g + 1]);
            }
            __syncwarp();
        }
        if (GW > 1) named_bar_arrive(next_bar, 64);
    }
    __syncthreads();
    // fold: bin b = 8 copies x T tables, fixed order (copy order rotated by b so that a quarter-warp spreads over banks)
    for (int b = tid; b < k; b += kThreadsS) {
        double s = 0.0;
        for (int w = 0; w < T; ++w) {
            const double2* row = reinterpret_cast<const double2*>(tables + static_cast<size_t>(w) * table_doubles + static_cast<size_t>(b) * 8);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double2 x = row[(q + (b >> 1)) & 3];
                s += x.x + x.y;
            }
        }
        if constexpr (kRows) {
            partial_rows[static_cast<size_t>(blockIdx.x) * k + b] = s;
        } else {
            if (s != 0.0) atomicAdd(grad + b, s);
        }
    }
}

inline int stripe_tables64(int k) {
    const size_t budget = 227 * 1024 - 256;
    const size_t per_table = (static_cast<size_t>(k) + 4) * 64;
    const int choices[6] = {12, 6, 4, 3, 2, 1};
    for (int c : choices)
        if (per_table * c <= budget) return c;
    return 0;
}

template <bool kImplicit>
int launch_striped64(const int32_t* idx, const double* val, long long n, double* grad, int k, cudaStream_t st, double* rows,
                     int* n_rows, int id_base = 0, int k_total = 0) {
    const int k_ids = k_total > 0 ? k_total : k;
    const int T = stripe_tables64(k);
    const size_t smem = static_cast<size_t>(T) * (k + 4) * 64;
    const long long n_units = (n + kStripeUnit64 - 1) / kStripeUnit64;
    const long long want = (n_units + kStripeWarps - 1) / kStripeWarps;
    const int sms = sm_count();
    const int grid = static_cast<int>(want < sms ? want : sms);
    if (rows) {
        auto kern = accumulate_striped64_kernel<kImplicit, true>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        kern<<<grid, kStripeWarps * 32, smem, st>>>(idx, val, n, grad ? grad + id_base : nullptr, k, rows, T, id_base, k_ids);
    } else {
        auto kern = accumulate_striped64_kernel<kImplicit, false>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        kern<<<grid, kStripeWarps * 32, smem, st>>>(idx, val, n, grad + id_base, k, nullptr, T, id_base, k_ids);
    }
    count_launch();
    if (n_rows) *n_rows = grid;
    return last_error();
}

// Multi-GPU finish (ONE CTA): add the CTAs' rows in CTA order, store the result into every rank's mailbox over
// NVLink, publish the sequence number, wait for all ranks, add the rows in rank order: grad[b] += global sum.
__global__ void __launch_bounds__(1024)
    accumulate_finish_allreduce_kernel(const float* __restrict__ rows, int n_rows, float* grad, int k, PeerArgs pa) {
    const int tid = threadIdx.x;
    const int par = static_cast<int>(pa.seq & 1ull);
    for (int b = tid; b < k; b += 1024) {
        float s = 0.f;
        for (int r = 0; r < n_rows; ++r) s += rows[static_cast<size_t>(r) * k + b];
        for (int p = 0; p < pa.world; ++p) pa.box[p]->vec[par][pa.rank][b] = s;
    }
    __threadfence_system();
    const bool ok = peer_publish_and_wait_cta(pa, tid);
    for (int b = tid; b < k; b += 1024) {
        float s = 0.f;
        for (int q = 0; q < pa.world; ++q) s += ld_relaxed_sys_f32(&pa.box[pa.rank]->vec[par][q][b]);
        grad[b] += ok ? s : __int_as_float(0x7fc00000);
    }
}

// one extra partial row holding the (< 4) head elements peeled off a misaligned shard
__global__ void __launch_bounds__(256)
    accumulate_head_row_kernel(const int32_t* __restrict__ idx, const float* __restrict__ val, int head, float* row, int k) {
    for (int b = threadIdx.x; b < k; b += 256) {
        float s = 0.f;
        for (int e = 0; e < head; ++e) {
            if (idx[e] == b) s += val[e];
        }
        row[b] = s;
    }
}

// deterministic finish: grad[b] += sum over the rows in a FIXED order.  Block = 32 bins x 8 row groups: thread (b, g)
// adds rows g, g + 8, g + 16, ... (coalesced 128-byte reads per row), the 8 group sums are then added in group order.
template <class T>
__global__ void __launch_bounds__(256) accumulate_finish_kernel(const T* __restrict__ rows, int n_rows, T* grad, int k) {
    __shared__ T part[8][33];
    const int bl = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int b = blockIdx.x * 32 + bl;
    T s = T(0);
    if (b < k) {
        T s0 = T(0), s1 = T(0);
        int r = g;
        for (; r + 8 < n_rows; r += 16) {
            s0 += rows[static_cast<size_t>(r) * k + b];
            s1 += rows[static_cast<size_t>(r + 8) * k + b];
        }
        if (r < n_rows) s0 += rows[static_cast<size_t>(r) * k + b];
        s = s0 + s1;
    }
    part[g][bl] = s;
    __syncthreads();
    if (g == 0 && b < k) {
        T t = part[0][bl];
#pragma unroll
        for (int q = 1; q < 8; ++q) t += part[q][bl];
        grad[b] += t;
    }
}

// K too large for shared-memory tables: coalesced loads + native global REDs (what the reference does, minus its
// strided access), warp-aggregated: lanes of a warp that hit the same bin are grouped with match.any, their values
// are summed with shuffles and ONE atomic per distinct bin per warp goes to L2 -- skewed ids (Zipf) would otherwise
// serialise millions of same-address atomics (5.9 ms for 2^24 elements before, measured).  Never deterministic.
template <class T>
__global__ void accumulate_global_kernel(const int32_t* __restrict__ idx, const T* __restrict__ val, long long n,
                                         T* grad, int k) {
    const int lane = threadIdx.x & 31;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long n_up = (n + 31) / 32 * 32;  // whole warps iterate together
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n_up; i += stride) {
        int id = -1;
        T v = T(0);
        if (i < n) {
            id = idx ? __ldg(idx + i) : static_cast<int>(i % k);
            v = __ldg(val + i);
        }
        const bool valid = id >= 0 && id < k;
        const unsigned peers = __match_any_sync(kFull, valid ? id : (-1 - lane));
        const int leader = __ffs(static_cast<int>(peers)) - 1;
        unsigned rest = peers & (peers - 1u);
        while (rest) {  // same trip count for every lane of a group
            const int src = __ffs(static_cast<int>(rest)) - 1;
            const T other = __shfl_sync(peers, v, src);
            if (lane == leader) v += other;
            rest &= rest - 1u;
        }
        if (valid && lane == leader) atomicAdd(grad + id, v);
    }
}

template <class T, bool kVec, bool kImplicit>
int launch_tables(const int32_t* idx, const T* val, long long n, T* grad, int k, cudaStream_t st, bool deterministic,
                  int grid, size_t smem) {
    if (deterministic) {
        void* scratch = nullptr;
        int err = scratch_get(SCRATCH_REDUCE, 256 + static_cast<size_t>(grid) * k * sizeof(T), &scratch, st);
        if (err) return err;
        T* rows = reinterpret_cast<T*>(reinterpret_cast<unsigned char*>(scratch) + 256);
        auto kern = accumulate_kernel<T, kVec, kImplicit, true>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        kern<<<grid, kThreads, smem, st>>>(idx, val, n, grad, k, rows);
        accumulate_finish_kernel<T><<<(k + 31) / 32, 256, 0, st>>>(rows, grid, grad, k);
        count_launch(2);
    } else {
        auto kern = accumulate_kernel<T, kVec, kImplicit, false>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        kern<<<grid, kThreads, smem, st>>>(idx, val, n, grad, k, nullptr);
        count_launch();
    }
    return last_error();
}

template <class T, int W, bool kImplicit>
int launch_tagged(const int32_t* idx, const T* val, long long n, T* grad, int k, cudaStream_t st, bool deterministic) {
    const size_t smem = static_cast<size_t>(W) * k * (sizeof(T) + 1);
    const long long n_chunks = (n + 127) / 128;
    const long long want = (n_chunks + W - 1) / W;
    const int sms = sm_count();
    const int grid = static_cast<int>(want < sms ? want : sms);  // one CTA per SM (the tables fill shared memory)
    if (deterministic) {
        // rows per CTA (static element -> warp -> CTA assignment, lane-ordered duplicate folding), summed in CTA order
        void* scratch = nullptr;
        int err = scratch_get(SCRATCH_REDUCE, 256 + static_cast<size_t>(grid) * k * sizeof(T), &scratch, st);
        if (err) return err;
        T* rows = reinterpret_cast<T*>(reinterpret_cast<unsigned char*>(scratch) + 256);
        auto kern = accumulate_tagged_kernel<T, W, kImplicit, true, true>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        kern<<<grid, W * 32, smem, st>>>(idx, val, n, grad, k, rows, 1);
        accumulate_finish_kernel<T><<<(k + 31) / 32, 256, 0, st>>>(rows, grid, grad, k);
        count_launch(2);
        return last_error();
    }
    auto kern = accumulate_tagged_kernel<T, W, kImplicit, false, false>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    kern<<<grid, W * 32, smem, st>>>(idx, val, n, grad, k, nullptr, 1);
    count_launch();
    return last_error();
}

// rows flavour for the multi-GPU path (fp32, K <= XYZ_PEER_VEC_FLOATS): returns the number of rows written
template <int W, bool kImplicit>
int launch_tagged_rows(const int32_t* idx, const float* val, long long n, int k, cudaStream_t st, float* rows, int* n_rows) {
    const size_t smem = static_cast<size_t>(W) * k * 5;
    const long long n_chunks = (n + 127) / 128;
    const long long want = (n_chunks + W - 1) / W;
    const int sms = sm_count();
    const int grid = static_cast<int>(want < sms ? want : sms);
    auto kern = accumulate_tagged_kernel<float, W, kImplicit, true, false>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    kern<<<grid, W * 32, smem, st>>>(idx, val, n, nullptr, k, rows,
                                     (aligned16(val) && (kImplicit || aligned16(idx))) ? 1 : 0);
    count_launch();
    *n_rows = grid;
    return last_error();
}

template <class T, bool kImplicit>
int launch_tagged_any(const int32_t* idx, const T* val, long long n, T* grad, int k, cudaStream_t st, bool deterministic,
                      bool* done) {
    const size_t per_warp = static_cast<size_t>(k) * (sizeof(T) + 1);
    const size_t budget = 200 * 1024;
    *done = true;
    if (per_warp * 32 <= budget) return launch_tagged<T, 32, kImplicit>(idx, val, n, grad, k, st, deterministic);
    if (per_warp * 16 <= budget) return launch_tagged<T, 16, kImplicit>(idx, val, n, grad, k, st, deterministic);
    if (per_warp * 8 <= budget) return launch_tagged<T, 8, kImplicit>(idx, val, n, grad, k, st, deterministic);
    *done = false;
    return 0;
}

template <class T>
int accumulate(const int32_t* idx, const T* val, long long n, T* grad, int k, void* stream, int flags) {
    if (n < 0 || k <= 0 || !grad) return XYZ_ERR_INVALID_ARGUMENT;
    if (n == 0) return 0;
    if (!val) return XYZ_ERR_INVALID_ARGUMENT;
    const bool implicit = (idx == nullptr);
    if (implicit && !(flags & XYZ_FLAG_IMPLICIT_IDS)) return XYZ_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool deterministic = (flags & XYZ_FLAG_DETERMINISTIC) != 0;
    const bool vec_ok = aligned16(val) && (implicit || aligned16(idx));
    // A slice [b, e) of two 16-byte aligned arrays (a multi-GPU shard) is misaligned by the SAME number of elements in
    // idx and val: peel the head elements (< 16 bytes' worth) with the small kernel and run the vector paths on the
    // aligned remainder.  Two launches, fixed order, so XYZ_FLAG_DETERMINISTIC still holds.
    if (!vec_ok && !implicit && n >= (1 << 16) + 4) {
        const uintptr_t mv = reinterpret_cast<uintptr_t>(val) & 15u, mi = reinterpret_cast<uintptr_t>(idx) & 15u;
        if (mv % sizeof(T) == 0 && mi % sizeof(int32_t) == 0) {
            const long long head_v = static_cast<long long>(((16u - mv) & 15u) / sizeof(T));
            const long long head_i = static_cast<long long>(((16u - mi) & 15u) / sizeof(int32_t));
            if (head_v == head_i && head_v > 0) {
                const int err = accumulate<T>(idx, val, head_v, grad, k, stream, flags);
                if (err) return err;
                return accumulate<T>(idx + head_v, val + head_v, n - head_v, grad, k, stream, flags);
            }
        }
    }
    // tagged tables (fast path; with XYZ_FLAG_DETERMINISTIC its lane-ordered flavour): as many warps as 200 KB of
    // shared memory hold
    if constexpr (sizeof(T) == 4) {
        // lane-striped tables with turn-taking warps: fp32, aligned arrays, K small enough for two tables per SM
        if (vec_ok && n >= (1 << 16) && stripe_tables(k) > 0) {
            float* rows = nullptr;
            int n_rows = 0;
            if (deterministic) {
                void* scratch = nullptr;
                const int err = scratch_get(SCRATCH_REDUCE, 256 + static_cast<size_t>(sm_count()) * k * sizeof(float), &scratch, st);
                if (err) return err;
                rows = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(scratch) + 256);
            }
            const int err = implicit ? launch_striped<true>(idx, val, n, grad, k, st, rows, &n_rows)
                                     : launch_striped<false>(idx, val, n, grad, k, st, rows, &n_rows);
            if (err || !deterministic) return err;
            accumulate_finish_kernel<float><<<(k + 31) / 32, 256, 0, st>>>(rows, n_rows, grad, k);
            count_launch();
            return last_error();
        }
        // K beyond two tables per SM: one pass of the same kernel per kStripePassBins bins (ids outside the pass fail the
        // range check and go to the dummy rows).  Distribution-independent and deterministic like the single pass.
        // (also for short inputs when a fixed order is requested and the per-warp tables of the small kernel do not fit)
        if (vec_ok && k <= kStripeMultiPassMaxK &&
            (n >= (1 << 16) || (deterministic && static_cast<size_t>(kWarps) * k * sizeof(T) > 200 * 1024))) {
            float* rows = nullptr;
            if (deterministic) {
                void* scratch = nullptr;
                const int err =
                    scratch_get(SCRATCH_REDUCE, 256 + static_cast<size_t>(sm_count()) * kStripePassBins * sizeof(float), &scratch, st);
                if (err) return err;
                rows = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(scratch) + 256);
            }
            for (int base = 0; base < k; base += kStripePassBins) {
                const int kp = (k - base < kStripePassBins) ? k - base : kStripePassBins;
                int n_rows = 0;
                const int err = implicit ? launch_striped<true>(idx, val, n, grad, kp, st, rows, &n_rows, base, k)
                                         : launch_striped<false>(idx, val, n, grad, kp, st, rows, &n_rows, base, k);
                if (err) return err;
                if (deterministic) {
                    accumulate_finish_kernel<float><<<(kp + 31) / 32, 256, 0, st>>>(rows, n_rows, grad + base, kp);
                    count_launch();
                }
            }
            return last_error();
        }
    }
    if constexpr (sizeof(T) == 8) {
        // fp64: the same striped design with 8 copies per bin and two half-warp update phases
        if (vec_ok && n >= (1 << 16) && stripe_tables64(k) > 0) {
            double* rows = nullptr;
            int n_rows = 0;
            if (deterministic) {
                void* scratch = nullptr;
                const int err = scratch_get(SCRATCH_REDUCE, 256 + static_cast<size_t>(sm_count()) * k * sizeof(double), &scratch, st);
                if (err) return err;
                rows = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(scratch) + 256);
            }
            const int err = implicit ? launch_striped64<true>(idx, val, n, grad, k, st, rows, &n_rows)
                                     : launch_striped64<false>(idx, val, n, grad, k, st, rows, &n_rows);
            if (err || !deterministic) return err;
            accumulate_finish_kernel<double><<<(k + 31) / 32, 256, 0, st>>>(rows, n_rows, grad, k);
            count_launch();
            return last_error();
        }
        if (vec_ok && n >= (1 << 16) && k <= kStripeMultiPassMaxK) {  // passes of kStripePassBins bins, as in fp32
            double* rows = nullptr;
            if (deterministic) {
                void* scratch = nullptr;
                const int err =
                    scratch_get(SCRATCH_REDUCE, 256 + static_cast<size_t>(sm_count()) * kStripePassBins * sizeof(double), &scratch, st);
                if (err) return err;
                rows = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(scratch) + 256);
            }
            constexpr int kBins = kStripePassBins - 4;  // four dummy rows in the fp64 layout
            for (int base = 0; base < k; base += kBins) {
                const int kp = (k - base < kBins) ? k - base : kBins;
                int n_rows = 0;
                const int err = implicit ? launch_striped64<true>(idx, val, n, grad, kp, st, rows, &n_rows, base, k)
                                         : launch_striped64<false>(idx, val, n, grad, kp, st, rows, &n_rows, base, k);
                if (err) return err;
                if (deterministic) {
                    accumulate_finish_kernel<double><<<(kp + 31) / 32, 256, 0, st>>>(rows, n_rows, grad + base, kp);
                    count_launch();
                }
            }
            return last_error();
        }
    }
    if (vec_ok && n >= (1 << 16)) {
        bool done = false;
        const int err = implicit ? launch_tagged_any<T, true>(idx, val, n, grad, k, st, deterministic, &done)
                                 : launch_tagged_any<T, false>(idx, val, n, grad, k, st, deterministic, &done);
        if (done) return err;
    }
    const size_t smem = static_cast<size_t>(kWarps) * k * sizeof(T);
    const int sms = sm_count();
    if (smem > 200 * 1024) {
        if (deterministic) return XYZ_ERR_INVALID_ARGUMENT;  // no fixed-order path for K this large
        accumulate_global_kernel<T><<<sms * 8, 256, 0, st>>>(idx, val, n, grad, k);
        count_launch();
        return last_error();
    }
    int ctas_per_sm = static_cast<int>((220 * 1024) / (smem + 1024));
    if (ctas_per_sm > kCtasPerSM) ctas_per_sm = kCtasPerSM;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    const long long n_chunks = (n + 127) / 128;
    long long want = (n_chunks + kWarps - 1) / kWarps;
    const long long max_ctas = static_cast<long long>(sms) * ctas_per_sm;
    const int grid = static_cast<int>(want < max_ctas ? want : max_ctas);
    const bool vec = aligned16(val) && (implicit || aligned16(idx));
    if (vec) {
        return implicit ? launch_tables<T, true, true>(idx, val, n, grad, k, st, deterministic, grid, smem)
                        : launch_tables<T, true, false>(idx, val, n, grad, k, st, deterministic, grid, smem);
    }
    return implicit ? launch_tables<T, false, true>(idx, val, n, grad, k, st, deterministic, grid, smem)
                    : launch_tables<T, false, false>(idx, val, n, grad, k, st, deterministic, grid, smem);
}

}  // namespace
}  // namespace xyzb

extern "C" int xyz_accumulate_f32(const int32_t* idx, const float* val, long long n, float* grad, int k, void* stream,
                                  int flags) {
    return xyzb::accumulate<float>(idx, val, n, grad, k, stream, flags);
}
extern "C" int xyz_accumulate_f64(const int32_t* idx, const double* val, long long n, double* grad, int k,
                                  void* stream, int flags) {
    return xyzb::accumulate<double>(idx, val, n, grad, k, stream, flags);
}

extern "C" int xyz_accumulate_f32_allreduce(const int32_t* idx, const float* val, long long n, float* grad, int k,
                                            const xyz_peer_group* group, unsigned long long seq, void* stream, int flags) {
    using namespace xyzb;
    if (!group || group->world < 1 || group->world > XYZ_PEER_MAX_WORLD || group->rank < 0 || group->rank >= group->world ||
        seq == 0 || n < 0 || k <= 0 || k > XYZ_PEER_VEC_FLOATS || !grad)
        return XYZ_ERR_INVALID_ARGUMENT;
    const bool implicit = (idx == nullptr);
    if (n > 0 && (!val || (implicit && !(flags & XYZ_FLAG_IMPLICIT_IDS)))) return XYZ_ERR_INVALID_ARGUMENT;
    PeerArgs pa{};
    pa.rank = group->rank;
    pa.world = group->world;
    pa.seq = seq;
    for (int i = 0; i < group->world; ++i) {
        if (!group->mailbox[i]) return XYZ_ERR_INVALID_ARGUMENT;
        pa.box[i] = static_cast<PeerMailbox*>(group->mailbox[i]);
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    void* scratch = nullptr;
    int err = scratch_get(SCRATCH_REDUCE, 256 + static_cast<size_t>(sm_count() + 1) * k * sizeof(float), &scratch, st);
    if (err) return err;
    float* rows = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(scratch) + 256);
    int n_rows = 0;
    if (n > 0) {
        // tables of 8 / 16 / 32 warps as shared memory allows (K <= 4096 always fits 8 warps); unaligned slices
        // (shard boundaries) take the same kernel with scalar loads
        const size_t per_warp = static_cast<size_t>(k) * 5;
        // misaligned shard of aligned arrays: head elements into one extra row, the vector kernel on the rest
        const uintptr_t mv = reinterpret_cast<uintptr_t>(val) & 15u, mi = reinterpret_cast<uintptr_t>(idx) & 15u;
        const int head = (!implicit && mv == mi && mv != 0 && (mv & 3u) == 0) ? static_cast<int>((16u - mv) / 4u) : 0;
        if (head > 0 && n >= (1 << 16) + 4 && stripe_tables(k) > 0) {
            err = launch_striped<false>(idx + head, val + head, n - head, nullptr, k, st, rows, &n_rows);
            if (err) return err;
            accumulate_head_row_kernel<<<1, 256, 0, st>>>(idx, val, head, rows + static_cast<size_t>(n_rows) * k, k);
            count_launch();
            ++n_rows;
        } else if (n >= (1 << 16) && aligned16(val) && (implicit || aligned16(idx)) && stripe_tables(k) > 0)
            err = implicit ? launch_striped<true>(idx, val, n, nullptr, k, st, rows, &n_rows)
                           : launch_striped<false>(idx, val, n, nullptr, k, st, rows, &n_rows);
        else if (per_warp * 32 <= 200 * 1024)
            err = implicit ? launch_tagged_rows<32, true>(idx, val, n, k, st, rows, &n_rows)
                           : launch_tagged_rows<32, false>(idx, val, n, k, st, rows, &n_rows);
        else if (per_warp * 16 <= 200 * 1024)
            err = implicit ? launch_tagged_rows<16, true>(idx, val, n, k, st, rows, &n_rows)
                           : launch_tagged_rows<16, false>(idx, val, n, k, st, rows, &n_rows);
        else
            err = implicit ? launch_tagged_rows<8, true>(idx, val, n, k, st, rows, &n_rows)
                           : launch_tagged_rows<8, false>(idx, val, n, k, st, rows, &n_rows);
        if (err) return err;
    }
    accumulate_finish_allreduce_kernel<<<1, 1024, 0, st>>>(rows, n_rows, grad, k, pa);
    count_launch();
    return last_error();
}
