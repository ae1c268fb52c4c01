// xyz_autodiff/detail/config.cuh -- build-mode glue of the B200 header set.
//
// Every function of this library is usable from device code AND from host code: the same
// translation unit compiles under nvcc (sm_100a) and under a plain C++20 host compiler.  That is
// what makes the "host-compiled path" of the reference (which needs a macro shim, SURVEY.md
// section 0) a first-class citizen here.
#pragma once

#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define XYZ_HD __host__ __device__ __forceinline__
#define XYZ_HD_NOINLINE __host__ __device__
#else
#define XYZ_HD inline
#define XYZ_HD_NOINLINE
#endif

namespace xyz_autodiff::detail {

// Leaf-gradient accumulation primitive behind VariableRef::add_grad
// (reference: include/xyz_autodiff/variable.cuh:48-50 -- an unconditional atomicAdd).
// Device: one native RED (red.global.add / atoms) when the pointer is in global or shared memory; a plain
// += when it points into the calling thread's own local storage (atomics on the local window are illegal --
// the reference documents "CANNOT USE BUFFER ON A LOCAL VARIABLE", variable.cuh:9-10; here it just works).
// Host: plain read-modify-write (a host thread owns its accumulators).
//
// Warp aggregation (device, default; define XYZ_AUTODIFF_PLAIN_ATOMICS to get one atomic per thread like the
// reference): the lanes of the warp that execute the add_grad together are grouped by target address with ONE
// match.any; each group's values are summed with shuffles and its lowest lane issues ONE atomic.  The reference's
// own usage -- every thread of a kernel holding a VariableRef onto the same parameter
// (tests/test_parallel_gradient_accumulation.cu:32-43, examples/optimization/linear_regression_sgd.cu:93-122,
// gaussian_splatting_kernel.cu:79-110) -- thus sends 1 instead of 32 same-address atomics per warp to L2, where
// same-address atomics are serialised (measured on B200, dev/addgrad_lab.cu).  Unmodified user kernels get this
// by compiling against these headers.
#if defined(__CUDA_ARCH__) && !defined(XYZ_AUTODIFF_PLAIN_ATOMICS)
template <typename T>
__device__ __forceinline__ void accumulate_aggregated(T* address, T value) noexcept {
    const unsigned active = __activemask();
    const unsigned peers = __match_any_sync(active, reinterpret_cast<unsigned long long>(address));
    unsigned lane;
    asm("mov.u32 %0, %%laneid;" : "=r"(lane));
    if (peers == 0xffffffffu) {  // the whole warp on one address: butterfly, lane 0 adds
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) value += __shfl_xor_sync(0xffffffffu, value, o);
        if (lane == 0) atomicAdd(address, value);
        return;
    }
    const unsigned leader = static_cast<unsigned>(__ffs(static_cast<int>(peers)) - 1);
    unsigned rest = peers & (peers - 1u);  // the group without its leader; same trip count for all its lanes
    while (rest) {
        const int src = __ffs(static_cast<int>(rest)) - 1;
        const T other = __shfl_sync(peers, value, src);
        if (lane == leader) value += other;
        rest &= rest - 1u;
    }
    if (lane == leader) atomicAdd(address, value);
}
#endif

template <typename T>
XYZ_HD void accumulate(T* address, T value) noexcept {
#if defined(__CUDA_ARCH__)
    if (__isLocal(address)) {
        *address += value;
    } else {
#if defined(XYZ_AUTODIFF_PLAIN_ATOMICS)
        atomicAdd(address, value);
#else
        accumulate_aggregated(address, value);
#endif
    }
#else
    *address += value;
#endif
}

}  // namespace xyz_autodiff::detail
