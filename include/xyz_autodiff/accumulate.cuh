// xyz_autodiff/accumulate.cuh -- gradient-accumulation helpers for leaves shared by many threads.
//
// The reference has exactly one mechanism: VariableRef::add_grad = one scalar atomicAdd per component per
// thread on the shared parameter (include/xyz_autodiff/variable.cuh:48-50), exercised on global memory by
// tests/test_parallel_gradient_accumulation.cu:25-49 and on __shared__ memory by
// tests/test_shared_memory_atomic.cu:29-64.  With E threads and K parameters that is E*K same-address
// L2 atomics.  These helpers keep the per-thread graph code unchanged and move the reduction on chip:
//
//   RegisterLeaf<N, T>   a leaf (satisfies DifferentiableVariableConcept) whose VALUES are read from memory
//                        once and whose ADJOINTS accumulate in the thread's registers across as many
//                        graph evaluations as the thread performs;
//   LeafSlice / slice    a view of some components of a RegisterLeaf as a leaf of its own;
//   warp_reduce_add      shuffle-tree sum of N per-thread values over the 32 lanes (fixed order);
//   block_accumulate     warp shuffle -> one shared-memory row per warp -> fixed-order sum -> ONE
//                        atomic/RED per component per CTA (vector red.global.add.v4.f32 when N % 4 == 0
//                        and the target is 16-byte aligned).
//
// Device-only parts are guarded; on the host RegisterLeaf works as a plain accumulator and
// block_accumulate degenerates to `target[i] += value[i]`.
#pragma once

#include "concept/variable.cuh"
#include "detail/config.cuh"

namespace xyz_autodiff {
namespace accum {

template <std::size_t N, typename T>
    requires FloatingPointConcept<T>
class RegisterLeaf {
public:
    using value_type = T;
    static constexpr std::size_t size = N;

    // `values` may point to global, shared or host memory; they are copied into registers once
    XYZ_HD explicit RegisterLeaf(const T* values) {
#pragma unroll
        for (std::size_t i = 0; i < N; ++i) {
            v_[i] = values[i];
            g_[i] = T(0);
        }
    }

    XYZ_HD T& operator[](std::size_t i) noexcept { return v_[i]; }
    XYZ_HD const T& operator[](std::size_t i) const noexcept { return v_[i]; }
    XYZ_HD const T& grad(std::size_t i) const noexcept { return g_[i]; }
    XYZ_HD void add_grad(std::size_t i, T value) noexcept { g_[i] += value; }
    XYZ_HD void zero_grad() noexcept {
#pragma unroll
        for (std::size_t i = 0; i < N; ++i) g_[i] = T(0);
    }
    XYZ_HD T* grad() noexcept { return g_; }
    XYZ_HD const T* grad() const noexcept { return g_; }
    XYZ_HD T* data() noexcept { return v_; }
    XYZ_HD const T* data() const noexcept { return v_; }

private:
    T v_[N];
    T g_[N];
};

// View of components [Offset, Offset + N) of a RegisterLeaf (or of any leaf with data() / grad() arrays): lets ONE
// parameter block be used as several graph leaves (a, b, c, d of the least-squares example; W of the matrix chain).
// Values and adjoints stay where the parent keeps them (registers); add_grad is the parent's plain +=.
template <std::size_t Offset, std::size_t N, class Leaf>
class LeafSlice {
public:
    using value_type = typename Leaf::value_type;
    static constexpr std::size_t size = N;
    static_assert(Offset + N <= Leaf::size, "slice exceeds the leaf");

    XYZ_HD explicit LeafSlice(Leaf& leaf) : leaf_(&leaf) {}
    XYZ_HD value_type& operator[](std::size_t i) const noexcept { return (*leaf_)[Offset + i]; }
    XYZ_HD const value_type& grad(std::size_t i) const noexcept { return leaf_->grad(Offset + i); }
    XYZ_HD void add_grad(std::size_t i, value_type value) const noexcept { leaf_->add_grad(Offset + i, value); }
    XYZ_HD void zero_grad() const noexcept {
#pragma unroll
        for (std::size_t i = 0; i < N; ++i) leaf_->grad()[Offset + i] = value_type(0);
    }

private:
    Leaf* leaf_;
};

template <std::size_t Offset, std::size_t N, class Leaf>
XYZ_HD LeafSlice<Offset, N, Leaf> slice(Leaf& leaf) {
    return LeafSlice<Offset, N, Leaf>(leaf);
}

#if defined(__CUDACC__)
// Sum `values[0..N)` over the 32 lanes of the calling warp; every lane receives the totals.
template <std::size_t N, typename T>
__device__ __forceinline__ void warp_reduce_add(T (&values)[N]) {
#pragma unroll
    for (int offset = 16; offset > 0; offset >>= 1) {
#pragma unroll
        for (std::size_t i = 0; i < N; ++i) values[i] += __shfl_xor_sync(0xffffffffu, values[i], offset);
    }
}
#endif

// Reduce the per-thread `values` of a whole CTA and add the totals to target[0..N): one atomic per
// component per CTA.  All threads of the CTA must call it (it synchronises).  `scratch` is shared
// memory for (blockDim / 32) * N values of T.  blockDim.x * blockDim.y * blockDim.z must be a
// multiple of 32.
template <std::size_t N, typename T>
XYZ_HD void block_accumulate(T (&values)[N], T* target, T* scratch) {
#if defined(__CUDA_ARCH__)
    const unsigned tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const unsigned nwarps = (blockDim.x * blockDim.y * blockDim.z) >> 5;
    warp_reduce_add(values);
    if ((tid & 31u) == 0u) {
#pragma unroll
        for (std::size_t i = 0; i < N; ++i) scratch[(tid >> 5) * N + i] = values[i];
    }
    __syncthreads();
    if (tid == 0) {
        T total[N];
#pragma unroll
        for (std::size_t i = 0; i < N; ++i) total[i] = T(0);
        for (unsigned w = 0; w < nwarps; ++w) {
#pragma unroll
            for (std::size_t i = 0; i < N; ++i) total[i] += scratch[w * N + i];
        }
        bool done = false;
        if constexpr (sizeof(T) == 4 && (N % 4 == 0)) {
            if ((reinterpret_cast<unsigned long long>(target) & 15ull) == 0ull) {
#pragma unroll
                for (std::size_t i = 0; i < N; i += 4) {
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(__cvta_generic_to_global(target + i)),
                                 "f"(total[i]), "f"(total[i + 1]), "f"(total[i + 2]), "f"(total[i + 3])
                                 : "memory");
                }
                done = true;
            }
        }
        if (!done) {
#pragma unroll
            for (std::size_t i = 0; i < N; ++i) atomicAdd(target + i, total[i]);
        }
    }
    __syncthreads();
#else
    (void)scratch;
    for (std::size_t i = 0; i < N; ++i) target[i] += values[i];
#endif
}

}  // namespace accum
}  // namespace xyz_autodiff
