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
template <typename T>
XYZ_HD void accumulate(T* address, T value) noexcept {
#if defined(__CUDA_ARCH__)
    if (__isLocal(address)) {
        *address += value;
    } else {
        atomicAdd(address, value);
    }
#else
    *address += value;
#endif
}

}  // namespace xyz_autodiff::detail
