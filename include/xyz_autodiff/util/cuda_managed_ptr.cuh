// xyz_autodiff/util/cuda_managed_ptr.cuh -- RAII ownership of cudaMallocManaged memory.
// Contract of reference include/xyz_autodiff/util/cuda_managed_ptr.cuh:8-62: allocation failure is
// reported on std::cerr and yields a null pointer (it does NOT throw); release errors are ignored.
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <iostream>
#include <memory>
#include <type_traits>

template <typename T>
struct CudaManagedDeleter {
    void operator()(std::remove_extent_t<T>* p) const {
        if (p != nullptr) cudaFree(p);
    }
};

template <typename T>
using cuda_managed_ptr = std::unique_ptr<T, CudaManagedDeleter<T>>;

namespace xyz_autodiff::detail {
template <typename T>
inline T* managed_alloc(std::size_t count, const char* who) {
    void* raw = nullptr;
    const std::size_t bytes = sizeof(T) * count;
    const cudaError_t status = cudaMallocManaged(&raw, bytes);
    if (status != cudaSuccess) {
        std::cerr << "[CUDA ERROR] " << who << " failed to allocate " << bytes << " bytes: " << cudaGetErrorString(status)
                  << "\n";
        return nullptr;
    }
    return static_cast<T*>(raw);
}
}  // namespace xyz_autodiff::detail

template <typename T>
cuda_managed_ptr<T> makeCudaManagedUnique(std::size_t count = 1) {
    return cuda_managed_ptr<T>(xyz_autodiff::detail::managed_alloc<T>(count, "makeCudaManagedUnique"));
}

template <typename T>
cuda_managed_ptr<T[]> makeCudaManagedArray(std::size_t count) {
    return cuda_managed_ptr<T[]>(xyz_autodiff::detail::managed_alloc<T>(count, "makeCudaManagedArray"));
}
