// oracle/ref_cuda_lsq_shipped.cu -- TEST INFRASTRUCTURE: the reference's SHIPPED least-squares example
// (examples/optimization/linear_regression_sgd.cu) compiled UNMODIFIED from where it lies -- this file only renames its
// main() and adds launchers for its three file-local kernels, so that tests can run
// parallel_gradient_computation_kernel (:86-123), select_batch_kernel (:68-81) and update_parameters_kernel (:126-134)
// themselves on the B200.  Part of oracle/_ref/libxyz_ref_cuda.so; never linked by the product.
#define main xyz_reference_lsq_example_main
#include "linear_regression_sgd.cu"  // -I$(REF)/examples/optimization
#undef main

extern "C" {
// params = {value[4], grad[4]} (Parameters, :36-39); grad += like the kernel does.  Any batch size, 256-thread blocks.
int refcuda_lsq_shipped(const double* data, long long n, double* params) {
    parallel_gradient_computation_kernel<<<static_cast<unsigned>((n + 255) / 256), 256>>>(
        reinterpret_cast<const DataPoint*>(data), static_cast<int>(n), reinterpret_cast<Parameters*>(params));
    return static_cast<int>(cudaGetLastError());
}
int refcuda_lsq_update(double* params, double learning_rate, int batch_size) {
    update_parameters_kernel<<<1, 1>>>(reinterpret_cast<Parameters*>(params), learning_rate, batch_size);
    return static_cast<int>(cudaGetLastError());
}
}
