// xyz_autodiff/util/cuda_unique_ptr.cuh -- RAII ownership of cudaMalloc'ed device memory.
// Contract of reference include/xyz_autodiff/util/cuda_unique_ptr.cuh:10-53: allocation and release
// failures throw (CHECK_CUDA_ERROR); the names live in the global namespace.
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <memory>
#include <type_traits>

#include "error_checker.cuh"

template <typename T>
struct CudaDeleter {
    using element = std::remove_extent_t<T>;
    void operator()(element* p) const {
        if (p != nullptr) CHECK_CUDA_ERROR(cudaFree(p));
    }
};

template <typename T>
using cuda_unique_ptr = std::unique_ptr<T, CudaDeleter<T>>;

namespace xyz_autodiff::detail {
template <typename T>
inline T* device_alloc(std::size_t count) {
    void* raw = nullptr;
    CHECK_CUDA_ERROR(cudaMalloc(&raw, sizeof(T) * count));
    return static_cast<T*>(raw);
}
}  // namespace xyz_autodiff::detail

// `count` objects of T, owned as a single-object pointer (the reference's convention for test buffers)
template <typename T>
cuda_unique_ptr<T> makeCudaUnique(std::size_t count = 1) {
    return cuda_unique_ptr<T>(xyz_autodiff::detail::device_alloc<T>(count));
}

template <typename T>
cuda_unique_ptr<T[]> makeCudaUniqueArray(std::size_t count) {
    return cuda_unique_ptr<T[]>(xyz_autodiff::detail::device_alloc<T>(count));
}
