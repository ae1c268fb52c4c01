// xyz_autodiff/batched.cuh -- batched evaluation of a user graph: one element per thread, forward + reverse, with the
// adjoints of the SHARED parameters accumulated on chip.  sm_100a (B200), nvcc only.
//
// The reference's usage pattern (examples/optimization/linear_regression_sgd.cu:86-123,
// tests/test_parallel_gradient_accumulation.cu:25-49): a hand-written kernel in which every thread builds the graph
// on VariableRef leaves that point at ONE parameter block, so every leaf adjoint is a same-address atomicAdd
// (include/xyz_autodiff/variable.cuh:48-50), and per-element inputs are read with strided scalar loads.
// for_each() is that kernel written once, for any graph:
//
//   struct In  { ... };            // per-element inputs, trivially copyable (AoS row, contiguous in memory)
//   struct Out { ... };            // per-element results (values and/or per-element adjoints); NoOutput if none
//   struct MyGraph {
//       __device__ void operator()(const In& x, Out& y, accum::RegisterLeaf<NS, T>& shared) const {
//           auto a = accum::slice<0, 1>(shared);               // views of the shared parameters: values in
//           auto W = accum::slice<1, 9>(shared);               // registers, add_grad is a plain +=
//           Variable<6, T> J(x.J);                             // per-element leaves
//           auto node = op::matmul<2, 3, 3>(J, W); ... node.run();
//           y.gJ[i] = J.grad(i);
//       }
//   };
//   batched::Workspace ws;         // once
//   batched::for_each<In, Out, NS, T>(in, out, n, shared_values, shared_grads, MyGraph{}, ws, stream);
//
// Data path: persistent CTAs of 128 threads; a tile = 128 consecutive elements = one contiguous block of In, moved
// by ONE 1-D TMA bulk copy (cp.async.bulk ... mbarrier::complete_tx) into a 4-stage shared-memory ring; results go
// through a 2-stage ring and one TMA bulk store per tile.  The shared parameters' adjoints stay in the thread's
// registers for the whole persistent loop (RegisterLeaf), then shuffle tree -> shared memory -> one row per CTA ->
// the last CTA to finish (ticket) adds the rows in CTA order and does shared_grads[i] += total: no floating-point
// atomics, bit-identical run to run.  Unaligned bases and the tail (n % 128) take plain loads / stores in the same
// kernel.
#pragma once

#if defined(__CUDACC__)

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <type_traits>

#include "accumulate.cuh"

namespace xyz_autodiff {
namespace batched {

struct NoOutput {};  // Out of graphs without per-element results (only shared-parameter gradients)

constexpr int kThreads = 128;  // threads per CTA == elements per tile
constexpr int kInStages = 4;
constexpr int kOutStages = 2;

// Device scratch of for_each (one row of partial sums per CTA + the ticket).  Create once, reuse for every call on
// the same device; calls sharing a Workspace must be stream-ordered.
struct Workspace {
    void* ptr = nullptr;
    std::size_t bytes = 0;
    Workspace() = default;
    Workspace(const Workspace&) = delete;
    Workspace& operator=(const Workspace&) = delete;
    ~Workspace() {
        if (ptr) cudaFree(ptr);
    }
    cudaError_t reserve(std::size_t need) {
        if (need <= bytes) return cudaSuccess;
        if (ptr) {
            cudaError_t e = cudaDeviceSynchronize();  // launches that still use the old buffer
            if (e != cudaSuccess) return e;
            cudaFree(ptr);
            ptr = nullptr;
            bytes = 0;
        }
        cudaError_t e = cudaMalloc(&ptr, need);
        if (e != cudaSuccess) {
            ptr = nullptr;
            return e;
        }
        bytes = need;
        return cudaMemset(ptr, 0, need);  // the ticket starts at zero; every launch leaves it at zero
    }
};

namespace detail {

__device__ __forceinline__ std::uint32_t smem_u32(const void* p) {
    return static_cast<std::uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(std::uint64_t* bar, std::uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(std::uint64_t* bar, std::uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(std::uint64_t* bar, std::uint32_t parity) {
    std::uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, std::uint32_t bytes, std::uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(__cvta_generic_to_global(gmem_src)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, std::uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(gmem_dst)),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <class In, class Out, std::size_t NL, typename T>
struct Smem {
    static constexpr bool kHasOut = !std::is_same_v<Out, NoOutput>;
    alignas(128) unsigned char in[kInStages][kThreads * sizeof(In)];
    alignas(128) unsigned char out[kOutStages][kHasOut ? kThreads * sizeof(Out) : 16];
    std::uint64_t full[kInStages];
    T red[kThreads / 32][NL];
    int is_last;
};

template <class In, class Out, std::size_t NS, typename T, class F, bool kTma>
__global__ void __launch_bounds__(kThreads)
    batched_kernel(const In* __restrict__ in, Out* __restrict__ out, long long n, const T* __restrict__ shared_values,
                   T* shared_grads, T* rows, unsigned int* ticket, F f) {
    constexpr std::size_t NL = NS ? NS : 1;
    constexpr bool kHasOut = !std::is_same_v<Out, NoOutput>;
    using SmemT = Smem<In, Out, NL, T>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SmemT& sm = *reinterpret_cast<SmemT*>(smem_raw);
    const int tid = threadIdx.x;

    T init[NL];
#pragma unroll
    for (std::size_t i = 0; i < NL; ++i) init[i] = (i < NS) ? shared_values[i] : T(0);
    accum::RegisterLeaf<NL, T> shared(init);

    const long long n_tiles = kTma ? n / kThreads : 0;
    if constexpr (kTma) {
        constexpr std::uint32_t kInBytes = kThreads * sizeof(In);
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < kInStages; ++s) mbar_init(&sm.full[s], 1);
            mbar_fence_init();
        }
        __syncthreads();
        const long long first = blockIdx.x, stride = gridDim.x;
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < kInStages; ++s) {
                const long long t = first + s * stride;
                if (t < n_tiles) {
                    mbar_arrive_expect_tx(&sm.full[s], kInBytes);
                    bulk_load(sm.in[s], in + t * kThreads, kInBytes, &sm.full[s]);
                }
            }
        }
        int it = 0;
        for (long long tile = first; tile < n_tiles; tile += stride, ++it) {
            const int s = it % kInStages;
            mbar_wait(&sm.full[s], (it / kInStages) & 1);
            const In x = *reinterpret_cast<const In*>(sm.in[s] + tid * sizeof(In));
            if constexpr (kHasOut) {
                if (tid == 0) bulk_wait_read<kOutStages - 1>();  // the output stage about to be refilled has left
            }
            __syncthreads();  // input stage consumed by everyone
            if (tid == 0) {
                const long long nt = tile + static_cast<long long>(kInStages) * stride;
                if (nt < n_tiles) {
                    mbar_arrive_expect_tx(&sm.full[s], kInBytes);
                    bulk_load(sm.in[s], in + nt * kThreads, kInBytes, &sm.full[s]);
                }
            }
            Out y;
            f(x, y, shared);
            if constexpr (kHasOut) {
                unsigned char* o = sm.out[it % kOutStages];
                *reinterpret_cast<Out*>(o + tid * sizeof(Out)) = y;
                fence_proxy_async_smem();
                __syncthreads();
                if (tid == 0) {
                    bulk_store(out + tile * kThreads, o, kThreads * sizeof(Out));
                    bulk_commit();
                }
            }
        }
        if constexpr (kHasOut) {
            if (tid == 0) bulk_wait_all();  // shared memory must outlive the last stores
        }
    }
    // tail, and the whole range when a base pointer is not 16-byte aligned
    for (long long e = n_tiles * kThreads + blockIdx.x * static_cast<long long>(kThreads) + tid; e < n;
         e += static_cast<long long>(gridDim.x) * kThreads) {
        const In x = in[e];
        Out y;
        f(x, y, shared);
        if constexpr (kHasOut) out[e] = y;
    }

    if constexpr (NS > 0) {
        T g[NL];
#pragma unroll
        for (std::size_t i = 0; i < NL; ++i) g[i] = shared.grad(i);
        accum::warp_reduce_add(g);
        if ((tid & 31) == 0) {
#pragma unroll
            for (std::size_t i = 0; i < NL; ++i) sm.red[tid >> 5][i] = g[i];
        }
        __syncthreads();
        if (tid < static_cast<int>(NL)) {
            T s = T(0);
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) s += sm.red[w][tid];
            rows[static_cast<std::size_t>(blockIdx.x) * NL + tid] = s;
            __threadfence();
        }
        __syncthreads();
        if (tid == 0) sm.is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
        __syncthreads();
        if (!sm.is_last) return;
        __threadfence();
        // last CTA: thread t adds rows t, t + 128, ... ; then the same fixed-order tree
#pragma unroll
        for (std::size_t i = 0; i < NL; ++i) g[i] = T(0);
        for (unsigned int r = tid; r < gridDim.x; r += kThreads) {
#pragma unroll
            for (std::size_t i = 0; i < NL; ++i) g[i] += __ldcg(rows + static_cast<std::size_t>(r) * NL + i);
        }
        accum::warp_reduce_add(g);
        __syncthreads();
        if ((tid & 31) == 0) {
#pragma unroll
            for (std::size_t i = 0; i < NL; ++i) sm.red[tid >> 5][i] = g[i];
        }
        __syncthreads();
        if (tid < static_cast<int>(NL)) {
            T s = T(0);
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) s += sm.red[w][tid];
            shared_grads[tid] += s;
        }
        if (tid == 0) *ticket = 0u;
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<std::uintptr_t>(p) & 15u) == 0; }

}  // namespace detail

// out may be nullptr when Out is NoOutput; shared_values / shared_grads hold NS values of T (device memory,
// shared_grads is ACCUMULATED into); returns the launch error.  Asynchronous on `stream`.
template <class In, class Out, std::size_t NS, typename T, class F>
cudaError_t for_each(const In* in, Out* out, long long n, const T* shared_values, T* shared_grads, F f, Workspace& ws,
                     cudaStream_t stream = nullptr) {
    static_assert(std::is_trivially_copyable_v<In> && std::is_trivially_copyable_v<Out>, "In / Out must be PODs");
    static_assert(NS <= kThreads, "at most 128 shared parameters");
    constexpr std::size_t NL = NS ? NS : 1;
    using SmemT = detail::Smem<In, Out, NL, T>;
    constexpr bool kHasOut = !std::is_same_v<Out, NoOutput>;
    if (n < 0 || (n > 0 && !in) || (n > 0 && kHasOut && !out) || (NS > 0 && (!shared_values || !shared_grads)))
        return cudaErrorInvalidValue;
    if (n == 0) return cudaSuccess;
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    const bool tma = n >= kThreads && detail::aligned16(in) && (!kHasOut || detail::aligned16(out)) &&
                     (kThreads * sizeof(In)) % 16 == 0 && (!kHasOut || (kThreads * sizeof(Out)) % 16 == 0);
    auto kern = tma ? detail::batched_kernel<In, Out, NS, T, F, true> : detail::batched_kernel<In, Out, NS, T, F, false>;
    const std::size_t smem = sizeof(SmemT);
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    if (per_sm > 8) per_sm = 8;
    const long long max_ctas = static_cast<long long>(sms) * per_sm;
    const long long want = (n + kThreads - 1) / kThreads;
    const int grid = static_cast<int>(want < max_ctas ? want : max_ctas);
    e = ws.reserve(256 + static_cast<std::size_t>(max_ctas) * NL * sizeof(T));
    if (e != cudaSuccess) return e;
    unsigned int* ticket = static_cast<unsigned int*>(ws.ptr);
    T* rows = reinterpret_cast<T*>(static_cast<unsigned char*>(ws.ptr) + 256);
    kern<<<grid, kThreads, smem, stream>>>(in, out, n, shared_values, shared_grads, rows, ticket, f);
    return cudaGetLastError();
}

}  // namespace batched
}  // namespace xyz_autodiff

#endif  // __CUDACC__
