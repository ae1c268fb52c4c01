// xyz_autodiff/testing/unary_gradient_tester.cuh -- analytic vs central-difference gradients of a unary Logic.
// API of reference include/xyz_autodiff/testing/unary_gradient_tester.cuh:96-248 (UnaryGradientTester<Logic, In, Out>
// ::test / ::test_custom), without gtest and batched: one launch, one thread per random case (see
// gradient_report.cuh).  Both entry points also RETURN the GradientReport (the reference returns void).
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

// The reference's kernel body (unary_gradient_tester.cuh:35-93) for case `c` of a batch: inputs and upstream
// gradients come from the counter-based stream, both passes run on thread-local storage.
template <typename LogicType, std::size_t InDim, std::size_t OutDim>
__global__ void test_unary_gradient_kernel(LogicType logic, std::size_t num_cases, std::uint64_t seed, double delta,
                                           double lo, double hi, double* analytical, double* numerical, double* work) {
    using T = double;
    const std::size_t c = blockIdx.x * static_cast<std::size_t>(blockDim.x) + threadIdx.x;
    if (c >= num_cases) return;
    // leaf storage lives in global memory like the reference's test buffers (VariableRef::add_grad is an atomicAdd)
    T* x = work + c * (2 * InDim);
    T* gx = x + InDim;
    T gout[OutDim];
    for (std::size_t i = 0; i < InDim; ++i) x[i] = detail::uniform_at(seed, c, i, lo, hi);
    for (std::size_t j = 0; j < OutDim; ++j) gout[j] = detail::uniform_at(seed, c, InDim + j, lo, hi);
    VariableRef<InDim, T> input_var(x, gx);
    for (int pass = 0; pass < 2; ++pass) {
        for (std::size_t i = 0; i < InDim; ++i) gx[i] = T(0);
        auto op = UnaryOperation<OutDim, LogicType, VariableRef<InDim, T>>(logic, input_var);
        op.forward();
        op.zero_grad();
        for (std::size_t j = 0; j < OutDim; ++j) op.add_grad(j, gout[j]);
        if (pass == 0) op.backward(); else op.backward_numerical(delta);
        double* dst = (pass == 0 ? analytical : numerical) + c * InDim;
        for (std::size_t i = 0; i < InDim; ++i) dst[i] = input_var.grad(i);
    }
}

template <typename Logic, std::size_t InputDim, std::size_t OutputDim>
class UnaryGradientTester {
    static constexpr std::size_t NUM_TESTS = 100;
    static constexpr double TOLERANCE = 1e-5;
    static constexpr double DELTA = 1e-5;

public:
    // every case of one launch; `logic` lets parameterised Logics (a constant, a ConstArray) be tested too
    static GradientReport run(const std::string& operation_name, std::size_t num_tests, double tolerance, double delta,
                              double input_min, double input_max, Logic logic = Logic{}, std::uint64_t seed = 42) {
        GradientReport rep;
        rep.name = operation_name;
        rep.num_tests = num_tests;
        rep.tolerance = tolerance;
        rep.delta = delta;
        if (num_tests == 0) return rep;
        try {
            auto d_a = makeCudaUniqueArray<double>(num_tests * InputDim);
            auto d_n = makeCudaUniqueArray<double>(num_tests * InputDim);
            auto d_w = makeCudaUniqueArray<double>(num_tests * 2 * InputDim);
            const unsigned threads = 64, blocks = static_cast<unsigned>((num_tests + threads - 1) / threads);
            test_unary_gradient_kernel<Logic, InputDim, OutputDim><<<blocks, threads>>>(logic, num_tests, seed, delta, input_min,
                                                                                        input_max, d_a.get(), d_n.get(), d_w.get());
            CHECK_CUDA_ERROR(cudaGetLastError());
            CHECK_CUDA_ERROR(cudaDeviceSynchronize());
            std::vector<double> a(num_tests * InputDim), n(num_tests * InputDim);
            CHECK_CUDA_ERROR(cudaMemcpy(a.data(), d_a.get(), a.size() * sizeof(double), cudaMemcpyDeviceToHost));
            CHECK_CUDA_ERROR(cudaMemcpy(n.data(), d_n.get(), n.size() * sizeof(double), cudaMemcpyDeviceToHost));
            for (std::size_t c = 0; c < num_tests; ++c)
                for (std::size_t i = 0; i < InputDim; ++i) {
                    const double av = a[c * InputDim + i], nv = n[c * InputDim + i];
                    const double err = compute_error_min(av, nv);
                    if (!(err <= tolerance)) ++rep.num_failures;  // NaN counts as a failure
                    if (err > rep.max_error || std::isnan(err)) {
                        rep.max_error = err;
                        rep.max_error_case = c;
                        rep.max_error_index = i;
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

    // reference :102-154: 100 cases, inputs and upstream gradients ~ U(-2, 2), tolerance 1e-5, delta 1e-5
    static GradientReport test(const std::string& operation_name) {
        return run(operation_name, NUM_TESTS, TOLERANCE, DELTA, -2.0, 2.0);
    }

    // reference :157-247, including its rule that a tolerance outside [0, 1e-5] is forbidden (:166-169) and the
    // summary block it prints
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
#define TEST_UNARY_GRADIENT(LogicType, InputDim, OutputDim, TestName) \
    TEST(GradientTest, TestName) { xyz_autodiff::testing::UnaryGradientTester<LogicType, InputDim, OutputDim>::test(#TestName); }
#define TEST_UNARY_GRADIENT_CUSTOM(LogicType, InputDim, OutputDim, TestName, NumTests, Tolerance, Delta)                      \
    TEST(GradientTest, TestName) {                                                                                            \
        xyz_autodiff::testing::UnaryGradientTester<LogicType, InputDim, OutputDim>::test_custom(#TestName, NumTests, Tolerance, \
                                                                                                Delta);                       \
    }
#endif

}  // namespace testing
}  // namespace xyz_autodiff
