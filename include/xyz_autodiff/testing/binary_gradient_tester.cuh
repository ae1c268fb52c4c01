// xyz_autodiff/testing/binary_gradient_tester.cuh -- analytic vs central-difference gradients of a binary Logic.
// API of reference include/xyz_autodiff/testing/binary_gradient_tester.cuh:109-... (BinaryGradientTester<Logic, In1,
// In2, Out>::test / ::test_custom), without gtest and batched (see gradient_report.cuh).
#pragma once

#if !defined(__CUDACC__)
#error "the gradient testers launch kernels: compile with nvcc"
#endif

#include <cuda_runtime.h>

#include <cmath>
#include <string>
#include <vector>

#include "../operations/operation.cuh"
#include "../util/cuda_unique_ptr.cuh"
#include "../variable.cuh"
#include "gradient_report.cuh"

namespace xyz_autodiff {
namespace testing {

template <typename LogicType, std::size_t In1Dim, std::size_t In2Dim, std::size_t OutDim>
__global__ void test_binary_gradient_kernel(LogicType logic, std::size_t num_cases, std::uint64_t seed, double delta,
                                            double lo, double hi, double* analytical, double* numerical, double* work) {
    using T = double;
    constexpr std::size_t kIn = In1Dim + In2Dim;
    const std::size_t c = blockIdx.x * static_cast<std::size_t>(blockDim.x) + threadIdx.x;
    if (c >= num_cases) return;
    // leaf storage lives in global memory like the reference's test buffers (VariableRef::add_grad is an atomicAdd)
    T* x1 = work + c * (2 * kIn);
    T* g1 = x1 + In1Dim;
    T* x2 = g1 + In1Dim;
    T* g2 = x2 + In2Dim;
    T gout[OutDim];
    for (std::size_t i = 0; i < In1Dim; ++i) x1[i] = detail::uniform_at(seed, c, i, lo, hi);
    for (std::size_t i = 0; i < In2Dim; ++i) x2[i] = detail::uniform_at(seed, c, In1Dim + i, lo, hi);
    for (std::size_t j = 0; j < OutDim; ++j) gout[j] = detail::uniform_at(seed, c, kIn + j, lo, hi);
    VariableRef<In1Dim, T> v1(x1, g1);
    VariableRef<In2Dim, T> v2(x2, g2);
    for (int pass = 0; pass < 2; ++pass) {
        for (std::size_t i = 0; i < In1Dim; ++i) g1[i] = T(0);
        for (std::size_t i = 0; i < In2Dim; ++i) g2[i] = T(0);
        auto op = BinaryOperation<OutDim, LogicType, VariableRef<In1Dim, T>, VariableRef<In2Dim, T>>(logic, v1, v2);
        op.forward();
        op.zero_grad();
        for (std::size_t j = 0; j < OutDim; ++j) op.add_grad(j, gout[j]);
        if (pass == 0) op.backward(); else op.backward_numerical(delta);
        double* dst = (pass == 0 ? analytical : numerical) + c * kIn;
        for (std::size_t i = 0; i < In1Dim; ++i) dst[i] = v1.grad(i);
        for (std::size_t i = 0; i < In2Dim; ++i) dst[In1Dim + i] = v2.grad(i);
    }
}

template <typename Logic, std::size_t Input1Dim, std::size_t Input2Dim, std::size_t OutputDim>
class BinaryGradientTester {
    static constexpr std::size_t NUM_TESTS = 100;
    static constexpr double TOLERANCE = 1e-5;
    static constexpr double DELTA = 1e-5;
    static constexpr std::size_t kIn = Input1Dim + Input2Dim;

public:
    static GradientReport run(const std::string& operation_name, std::size_t num_tests, double tolerance, double delta,
                              double input_min, double input_max, Logic logic = Logic{}, std::uint64_t seed = 42) {
        GradientReport rep;
        rep.name = operation_name;
        rep.num_tests = num_tests;
        rep.tolerance = tolerance;
        rep.delta = delta;
        if (num_tests == 0) return rep;
        try {
            auto d_a = makeCudaUniqueArray<double>(num_tests * kIn);
            auto d_n = makeCudaUniqueArray<double>(num_tests * kIn);
            auto d_w = makeCudaUniqueArray<double>(num_tests * 2 * kIn);
            const unsigned threads = 64, blocks = static_cast<unsigned>((num_tests + threads - 1) / threads);
            test_binary_gradient_kernel<Logic, Input1Dim, Input2Dim, OutputDim><<<blocks, threads>>>(
                logic, num_tests, seed, delta, input_min, input_max, d_a.get(), d_n.get(), d_w.get());
            CHECK_CUDA_ERROR(cudaGetLastError());
            CHECK_CUDA_ERROR(cudaDeviceSynchronize());
            std::vector<double> a(num_tests * kIn), n(num_tests * kIn);
            CHECK_CUDA_ERROR(cudaMemcpy(a.data(), d_a.get(), a.size() * sizeof(double), cudaMemcpyDeviceToHost));
            CHECK_CUDA_ERROR(cudaMemcpy(n.data(), d_n.get(), n.size() * sizeof(double), cudaMemcpyDeviceToHost));
            for (std::size_t c = 0; c < num_tests; ++c)
                for (std::size_t i = 0; i < kIn; ++i) {
                    const double av = a[c * kIn + i], nv = n[c * kIn + i];
                    const double err = compute_error_min(av, nv);
                    if (!(err <= tolerance)) ++rep.num_failures;
                    if (err > rep.max_error || std::isnan(err)) {
                        rep.max_error = err;
                        rep.max_error_case = c;
                        rep.max_error_index = i;  // input1 components first, then input2
                        rep.max_error_analytical = av;
                        rep.max_error_numerical = nv;
                    }
                }
        } catch (const std::exception& e) {
            ++rep.num_failures;
            rep.message = std::string("CUDA error: ") + e.what();
        }
        if (!rep.passed()) detail::report_failure(rep);
        return rep;
    }

    static GradientReport test(const std::string& operation_name) {
        return run(operation_name, NUM_TESTS, TOLERANCE, DELTA, -2.0, 2.0);
    }

    // reference :190-..., with the same forbidden-tolerance rule (:199-202) and summary block
    static GradientReport test_custom(const std::string& operation_name, std::size_t num_tests, double tolerance, double delta,
                                      double input_min = -2.0, double input_max = 2.0, Logic logic = Logic{}) {
        if (tolerance < 0.0 || tolerance > 1e-5) {
            GradientReport rep;
            rep.name = operation_name;
            rep.tolerance = tolerance;
            rep.num_failures = 1;
            rep.message = "FORBIDDEN: tolerance outside [0, 1e-5]; the maximum tolerance for double precision tests is 1e-5.";
            detail::report_failure(rep);
            return rep;
        }
        GradientReport rep = run(operation_name, num_tests, tolerance, delta, input_min, input_max, logic);
        rep.print();
        return rep;
    }
};

#ifdef GTEST_INCLUDE_GTEST_GTEST_H_
#define TEST_BINARY_GRADIENT(LogicType, Input1Dim, Input2Dim, OutputDim, TestName)                                      \
    TEST(GradientTest, TestName) {                                                                                      \
        xyz_autodiff::testing::BinaryGradientTester<LogicType, Input1Dim, Input2Dim, OutputDim>::test(#TestName);       \
    }
#endif

}  // namespace testing
}  // namespace xyz_autodiff
