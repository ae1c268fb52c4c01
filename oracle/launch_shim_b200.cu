// oracle/launch_shim_b200.cu -- TEST/BENCH INFRASTRUCTURE: the ONE symbol the reference's training application needs
// from its kernel file (launch_gaussian_splatting, examples/mini-gaussian-splatting/gaussian_splatting_kernel.cuh:38-47),
// defined on top of libxyz_b200.so.  Linked with the reference's own, unmodified gaussian_splatting_training.cu,
// gaussian_parameters.cu, image_utils.cpp and training_config.cpp (oracle/Makefile -> oracle/_ref/ref_trainer_on_b200)
// it turns the reference application into a user of this repo's kernels: the drop-in boundary of SURVEY 8b, exercised by
// the reference's own main().  Declarations come from the reference header, so a signature drift fails the build.
#include <iostream>

#include "gaussian_splatting_kernel.cuh"  // the reference's declaration and types

#include <xyz_b200.h>

static_assert(sizeof(GaussianParams) == sizeof(xyz_gaussian_params), "GaussianParams layout");
static_assert(sizeof(GaussianGrads) == sizeof(xyz_gaussian_grads), "GaussianGrads layout");
static_assert(sizeof(PixelOutput) == 3 * sizeof(float), "PixelOutput layout");

void launch_gaussian_splatting(const GaussianParams* device_gaussians, GaussianGrads* device_gradients,
                               const PixelOutput* device_target_image, PixelOutput* device_output_image,
                               float* device_total_loss, int image_width, int image_height, int num_gaussians) {
    const int rc = xyz_launch_gaussian_splatting(reinterpret_cast<const xyz_gaussian_params*>(device_gaussians),
                                                 reinterpret_cast<xyz_gaussian_grads*>(device_gradients),
                                                 reinterpret_cast<const float*>(device_target_image),
                                                 reinterpret_cast<float*>(device_output_image), device_total_loss, image_width,
                                                 image_height, num_gaussians, /*stream=*/nullptr, /*flags=*/0);
    if (rc != 0)  // the reference prints the launch error and returns (gaussian_splatting_kernel.cu:144-148)
        std::cerr << "CUDA kernel launch error: " << rc << std::endl;
}
