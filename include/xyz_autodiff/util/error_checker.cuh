// xyz_autodiff/util/error_checker.cuh -- CHECK_CUDA_ERROR(expr): throw std::runtime_error on failure.
// Contract of reference include/xyz_autodiff/util/error_checker.cuh:7-20 (namespace cuda, same macro).
#pragma once

#include <cuda_runtime_api.h>

#include <stdexcept>
#include <string>

namespace cuda {

template <typename File, typename Line>
inline void check_error(const ::cudaError_t status, File&& file, Line&& line) {
    if (status == ::cudaSuccess) return;
    std::string message = ::cudaGetErrorName(status);
    message += " (" + std::to_string(static_cast<int>(status)) + ")@" + std::string(file) + "#L" + std::to_string(line) +
               ": " + ::cudaGetErrorString(status);
    throw std::runtime_error(message);
}

}  // namespace cuda

#define CHECK_CUDA_ERROR(e) (cuda::check_error((e), __FILE__, __LINE__))
