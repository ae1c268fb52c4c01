// oracle/shim/cuda_runtime.h -- TEST INFRASTRUCTURE, not product code.
//
// A stand-in for <cuda_runtime.h> that lets g++ compile the reference's header-only library
// (/root/reference/include/xyz_autodiff/**) and the body of its splat kernel for the HOST.
// It blanks the CUDA qualifiers, supplies the built-in index variables as thread_local
// globals (set by the driver before each "thread" is run as a plain function call), and
// turns atomicAdd into a plain read-modify-write (every oracle thread owns private
// accumulation buffers, so no real atomicity is needed).
//
// Used only by oracle/ref_driver.cpp (-> oracle/_ref/libxyz_ref.so). Nothing shipped links it.
#pragma once

#include <math.h>
#include <stdlib.h>
#include <cmath>
#include <concepts>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <type_traits>
#include <utility>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)

struct uint3 {
    unsigned x, y, z;
};
struct dim3 {
    unsigned x, y, z;
    constexpr dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

extern thread_local uint3 threadIdx;
extern thread_local uint3 blockIdx;
extern thread_local dim3 blockDim;
extern thread_local dim3 gridDim;

template <class T>
inline T atomicAdd(T* p, T v) {
    T old = *p;
    *p = old + v;
    return old;
}

typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "host shim"; }
// used by util/cuda_unique_ptr.cuh (pulled in through gaussian_parameters.h); never called.
inline cudaError_t cudaMalloc(void** p, std::size_t n) { *p = ::malloc(n); return *p ? 0 : 2; }
inline cudaError_t cudaFree(void* p) { ::free(p); return cudaSuccess; }
