// oracle/shim/cuda_runtime_api.h -- TEST INFRASTRUCTURE. See cuda_runtime.h in this directory.
#pragma once
#include "cuda_runtime.h"
inline const char* cudaGetErrorName(cudaError_t) { return "host shim"; }
