// include/xyz_b200_compat.hpp -- source-level drop-in for the reference's only host launch function.
//
// The reference trainer (examples/mini-gaussian-splatting/gaussian_splatting_training.cu:138-147) calls
//   launch_gaussian_splatting(device_params, device_grads, device_target, device_output, device_loss, W, H, N);
// declared in examples/mini-gaussian-splatting/gaussian_splatting_kernel.cuh:38-47.  Including this header
// instead of gaussian_splatting_kernel.cuh and linking libxyz_b200.so keeps that call site unchanged.
// GaussianParams / GaussianGrads / PixelOutput are only forward-declared here: their layouts (9 floats, 9 floats,
// 3 floats) are the ones of gaussian_parameters.h:12-41 and ConstArray<float, 3>.
//
// Error behaviour mirrors the reference (gaussian_splatting_kernel.cu:144-148): the function returns void and
// reports a launch failure on std::cerr.
#pragma once

#include <iostream>

#include "xyz_b200.h"

struct GaussianParams;
struct GaussianGrads;
namespace xyz_autodiff {
template <typename T, int N>
struct ConstArray;
}
using PixelOutput = xyz_autodiff::ConstArray<float, 3>;

#ifndef XYZ_B200_COMPAT_FLAGS
#define XYZ_B200_COMPAT_FLAGS 0  // fast-math flavour, atomics: what the reference training app builds
#endif

inline void launch_gaussian_splatting(const GaussianParams* device_gaussians, GaussianGrads* device_gradients,
                                      const PixelOutput* device_target_image, PixelOutput* device_output_image,
                                      float* device_total_loss, int image_width, int image_height, int num_gaussians) {
    const int rc = xyz_launch_gaussian_splatting(
        reinterpret_cast<const xyz_gaussian_params*>(device_gaussians),
        reinterpret_cast<xyz_gaussian_grads*>(device_gradients), reinterpret_cast<const float*>(device_target_image),
        reinterpret_cast<float*>(device_output_image), device_total_loss, image_width, image_height, num_gaussians,
        /*stream=*/nullptr, XYZ_B200_COMPAT_FLAGS);
    if (rc != 0) std::cerr << "Kernel launch error: xyz_launch_gaussian_splatting returned " << rc << std::endl;
}
