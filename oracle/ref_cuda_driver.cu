// oracle/ref_cuda_driver.cu -- TEST/BENCH INFRASTRUCTURE: the REFERENCE's own CUDA path built for sm_100a, so
// that bench.py can time "the reference's CUDA kernels on one B200" next to this repo's kernels
// (BASELINE.md section 4.1).  Built only where /root/reference exists (oracle/Makefile ->
// oracle/_ref/libxyz_ref_cuda.so); never linked by the product.
//
//   refcuda_splat       calls the reference's launch_gaussian_splatting; gaussian_splatting_kernel.cu is
//                       compiled UNMODIFIED from where it lies, with the reference's own flags
//                       (examples/mini-gaussian-splatting/CMakeLists.txt:22-31).
//   refcuda_lsq         the graph of examples/optimization/tests/test_linear_regression_gradient.cu:44-78
//                       (gtest-bound file, so the kernel body is restated from the reference's op:: factories),
//                       256-thread blocks as at :253-257: 4 same-address fp64 atomics per thread.
//   refcuda_accumulate  VariableRef::add_grad per element (tests/test_parallel_gradient_accumulation.cu:32-43).
//   refcuda_covproj     the tree of reference op::matmul nodes per thread (same composition as
//                       oracle/ref_driver.cpp::covproj_one), reading and writing global memory directly.
// All pointers are device pointers; everything runs on the legacy default stream like the reference.
#include <cuda_runtime.h>

#include <xyz_autodiff/variable.cuh>
#include <xyz_autodiff/operations/operation.cuh>
#include <xyz_autodiff/operations/binary/add_logic.cuh>
#include <xyz_autodiff/operations/binary/mul_logic.cuh>
#include <xyz_autodiff/operations/binary/matmul_logic.cuh>
#include <xyz_autodiff/operations/unary/sub_constant_logic.cuh>
#include <xyz_autodiff/operations/unary/squared_logic.cuh>

#include "gaussian_parameters.h"
#include <xyz_autodiff/const_array.cuh>
using PixelOutputT = xyz_autodiff::ConstArray<float, 3>;
void launch_gaussian_splatting(const GaussianParams*, GaussianGrads*, const PixelOutputT*, PixelOutputT*, float*, int, int,
                               int);

using namespace xyz_autodiff;

struct RefDataPoint { double x1, x2, y; };
struct RefParameters { double value[4]; double grad[4]; };

__global__ void ref_lsq_kernel(const RefDataPoint* batch, long long n, RefParameters* params) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= n) return;
    VariableRef<1, double> a_var(&params->value[0], &params->grad[0]);
    VariableRef<1, double> b_var(&params->value[1], &params->grad[1]);
    VariableRef<1, double> c_var(&params->value[2], &params->grad[2]);
    VariableRef<1, double> d_var(&params->value[3], &params->grad[3]);
    const RefDataPoint& data = batch[idx];
    auto x1_minus_a = op::sub_constant(a_var, data.x1);
    auto x1_term = op::squared(x1_minus_a);
    auto x2_minus_c = op::sub_constant(c_var, data.x2);
    auto x2_squared = op::squared(x2_minus_c);
    auto x2_term = op::mul(b_var, x2_squared);
    auto combined_terms = op::add(x1_term, x2_term);
    auto y_pred = op::add(combined_terms, d_var);
    auto y_diff = op::sub_constant(y_pred, data.y);
    auto loss = op::squared(y_diff);
    loss.run();
}

__global__ void ref_accumulate_kernel(const int* idx, const float* val, long long n, float* grad) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    float v = val[i];
    VariableRef<1, float> ref(&v, grad + idx[i]);
    ref.add_grad(0, v);
}

__global__ void ref_covproj_kernel(const float* J, const float* W, const float* S, const float* g, float* out, float* gJ,
                                   float* gW, float* gS, long long n) {
    const long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (e >= n) return;
    const float* j = J + 6 * e; const float* w = W + 9 * e; const float* s = S + 6 * e; const float* u = g + 3 * e;
    float Sfull[9] = {s[0], s[1], s[2], s[1], s[3], s[4], s[2], s[4], s[5]};
    float Wt[9], Jt[6];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Wt[r * 3 + c] = w[c * 3 + r];
    for (int r = 0; r < 2; ++r) for (int c = 0; c < 3; ++c) Jt[c * 2 + r] = j[r * 3 + c];
    Variable<6, float> vJ(j), vJt(Jt);
    Variable<9, float> vW(w), vS(Sfull), vWt(Wt);
    auto Tm = op::matmul<2, 3, 3>(vJ, vW);
    auto U = op::matmul<2, 3, 3>(Tm, vS);
    auto Tt = op::matmul<3, 3, 2>(vWt, vJt);
    auto P = op::matmul<2, 3, 2>(U, Tt);
    P.forward();
    out[3 * e] = P[0]; out[3 * e + 1] = P[1]; out[3 * e + 2] = P[3];
    P.zero_grad();
    P.add_grad(0, u[0]); P.add_grad(1, u[1]); P.add_grad(3, u[2]);
    P.backward();
    for (int r = 0; r < 2; ++r) for (int c = 0; c < 3; ++c) gJ[6 * e + r * 3 + c] = vJ.grad(r * 3 + c) + vJt.grad(c * 2 + r);
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) gW[9 * e + r * 3 + c] = vW.grad(r * 3 + c) + vWt.grad(c * 3 + r);
    gS[6 * e] = vS.grad(0); gS[6 * e + 1] = vS.grad(1) + vS.grad(3); gS[6 * e + 2] = vS.grad(2) + vS.grad(6);
    gS[6 * e + 3] = vS.grad(4); gS[6 * e + 4] = vS.grad(5) + vS.grad(7); gS[6 * e + 5] = vS.grad(8);
}

extern "C" {
int refcuda_splat(const float* params, float* grads, const float* target, float* output, float* loss, int W, int H, int N) {
    launch_gaussian_splatting(reinterpret_cast<const GaussianParams*>(params), reinterpret_cast<GaussianGrads*>(grads),
                              reinterpret_cast<const PixelOutputT*>(target), reinterpret_cast<PixelOutputT*>(output), loss,
                              W, H, N);
    return static_cast<int>(cudaGetLastError());
}
int refcuda_lsq(const double* data, long long n, double* params) {
    ref_lsq_kernel<<<static_cast<unsigned>((n + 255) / 256), 256>>>(reinterpret_cast<const RefDataPoint*>(data), n,
                                                                    reinterpret_cast<RefParameters*>(params));
    return static_cast<int>(cudaGetLastError());
}
int refcuda_accumulate(const int* idx, const float* val, long long n, float* grad) {
    ref_accumulate_kernel<<<static_cast<unsigned>((n + 255) / 256), 256>>>(idx, val, n, grad);
    return static_cast<int>(cudaGetLastError());
}
int refcuda_covproj(const float* J, const float* W, const float* S, const float* g, float* out, float* gJ, float* gW,
                    float* gS, long long n) {
    ref_covproj_kernel<<<static_cast<unsigned>((n + 255) / 256), 256>>>(J, W, S, g, out, gJ, gW, gS, n);
    return static_cast<int>(cudaGetLastError());
}
}
