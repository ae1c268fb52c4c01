// xyz_autodiff/testing/network_gradient_tester.cuh -- analytic vs numerical gradients of a whole user network.
// API of reference include/xyz_autodiff/testing/network_gradient_tester.cuh:41-192: the network is a functor with
//     template <GradientTag tag> __device__ void operator()(ParameterStruct* value, ParameterStruct* diff, double delta)
// that builds its graph over `value`/`diff` and calls run() (Analytical) or run_numerical(delta) (Numerical);
// ParameterStruct is a struct of doubles.  No gtest: the verdict is the reference's (passed, max_error, details)
// tuple; test_random_cases runs all cases in one launch per tag (one thread per case).
#pragma once

#if !defined(__CUDACC__)
#error "the gradient testers launch kernels: compile with nvcc"
#endif

#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <string>
#include <tuple>
#include <vector>

#include "../util/cuda_unique_ptr.cuh"
#include "gradient_report.cuh"

namespace xyz_autodiff {
namespace testing {

enum class GradientTag { Analytical, Numerical };

template <typename ParameterStruct>
struct NetworkTestBuffer {
    ParameterStruct value;  // parameter values
    ParameterStruct diff;   // parameter gradients
};

template <GradientTag tag, typename NetworkFunction, typename ParameterStruct>
__global__ void run_network_kernel(NetworkFunction network, NetworkTestBuffer<ParameterStruct>* buffers, std::size_t n,
                                   double delta) {
    const std::size_t c = blockIdx.x * static_cast<std::size_t>(blockDim.x) + threadIdx.x;
    if (c < n) network.template operator()<tag>(&buffers[c].value, &buffers[c].diff, delta);
}

template <typename ParameterStruct, typename NetworkFunction>
class NetworkGradientTester {
    static_assert(sizeof(ParameterStruct) % sizeof(double) == 0, "ParameterStruct must be a struct of doubles");
    static constexpr int kParams = static_cast<int>(sizeof(ParameterStruct) / sizeof(double));
    using Buffer = NetworkTestBuffer<ParameterStruct>;

    static void gradients(NetworkFunction network, const std::vector<ParameterStruct>& params, double delta, bool numerical,
                          std::vector<ParameterStruct>& out) {
        const std::size_t n = params.size();
        std::vector<Buffer> host(n);
        for (std::size_t c = 0; c < n; ++c) {
            host[c].value = params[c];
            std::memset(&host[c].diff, 0, sizeof(ParameterStruct));
        }
        auto dev = makeCudaUniqueArray<Buffer>(n);
        CHECK_CUDA_ERROR(cudaMemcpy(dev.get(), host.data(), n * sizeof(Buffer), cudaMemcpyHostToDevice));
        const unsigned threads = 64, blocks = static_cast<unsigned>((n + threads - 1) / threads);
        if (numerical)
            run_network_kernel<GradientTag::Numerical><<<blocks, threads>>>(network, dev.get(), n, delta);
        else
            run_network_kernel<GradientTag::Analytical><<<blocks, threads>>>(network, dev.get(), n, delta);
        CHECK_CUDA_ERROR(cudaGetLastError());
        CHECK_CUDA_ERROR(cudaDeviceSynchronize());
        CHECK_CUDA_ERROR(cudaMemcpy(host.data(), dev.get(), n * sizeof(Buffer), cudaMemcpyDeviceToHost));
        out.resize(n);
        for (std::size_t c = 0; c < n; ++c) out[c] = host[c].diff;
    }

    // the reference's comparison (:96-123): relative error against the NUMERICAL value, + 1e-10
    static std::tuple<bool, double, std::string> compare(const ParameterStruct& a, const ParameterStruct& n, double tolerance) {
        const double* ap = reinterpret_cast<const double*>(&a);
        const double* np = reinterpret_cast<const double*>(&n);
        bool passed = true;
        double max_error = 0.0;
        std::string details;
        for (int i = 0; i < kParams; ++i) {
            const double error = std::fabs(ap[i] - np[i]);
            const double relative = error / (std::fabs(np[i]) + 1e-10);
            if (!(std::fmin(error, relative) <= tolerance)) {
                passed = false;
                if (error > max_error || details.empty()) {
                    max_error = error;
                    details = "Parameter " + std::to_string(i) + ": analytical=" + std::to_string(ap[i]) +
                              ", numerical=" + std::to_string(np[i]) + ", error=" + std::to_string(error);
                }
            }
        }
        return {passed, max_error, details};
    }

public:
    static std::tuple<bool, double, std::string> test_single_case(NetworkFunction network, const ParameterStruct& initial_params,
                                                                  double tolerance = 1e-5, double delta = 1e-7, int /*seed*/ = 42) {
        std::vector<ParameterStruct> a, n;
        gradients(network, {initial_params}, delta, false, a);
        gradients(network, {initial_params}, delta, true, n);
        return compare(a[0], n[0], tolerance);
    }

    // every parameter of every case ~ U(param_min, param_max); returns the number of failing cases in the report
    static GradientReport test_random_cases(NetworkFunction network, const std::string& name, std::size_t num_tests = 100,
                                            double tolerance = 1e-5, double delta = 1e-7, double param_min = -2.0,
                                            double param_max = 2.0, std::uint64_t seed = 42) {
        GradientReport rep;
        rep.name = name;
        rep.num_tests = num_tests;
        rep.tolerance = tolerance;
        rep.delta = delta;
        try {
            std::vector<ParameterStruct> params(num_tests), a, n;
            for (std::size_t c = 0; c < num_tests; ++c) {
                double* p = reinterpret_cast<double*>(&params[c]);
                for (int i = 0; i < kParams; ++i) p[i] = detail::uniform_at(seed, c, static_cast<std::uint64_t>(i), param_min, param_max);
            }
            gradients(network, params, delta, false, a);
            gradients(network, params, delta, true, n);
            for (std::size_t c = 0; c < num_tests; ++c) {
                auto [ok, err, msg] = compare(a[c], n[c], tolerance);
                if (!ok) {
                    ++rep.num_failures;
                    if (err >= rep.max_error) {
                        rep.max_error = err;
                        rep.max_error_case = c;
                        rep.message = msg;
                    }
                }
            }
        } catch (const std::exception& e) {
            ++rep.num_failures;
            rep.message = std::string("CUDA error: ") + e.what();
        }
        if (!rep.passed()) detail::report_failure(rep);
        return rep;
    }
};

}  // namespace testing
}  // namespace xyz_autodiff
