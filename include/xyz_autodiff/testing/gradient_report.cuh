// xyz_autodiff/testing/gradient_report.cuh -- result type and acceptance rule of the gradient testers.
//
// The reference's testers (include/xyz_autodiff/testing/*.cuh) are gtest fixtures: they ASSERT/EXPECT inside
// static member functions and launch one <<<1,1>>> kernel per random case with a host round trip each.  This
// re-host keeps the class names, template parameters, default constants (100 cases, tolerance 1e-5, delta 1e-5,
// inputs and upstream gradients ~ U(-2, 2)) and the acceptance rule
//     min(|a - n|, |a - n| / (|a| + 1e-15)) <= tolerance          (unary_gradient_tester.cuh:26-32)
// but needs no gtest: every case is one THREAD of a single launch and the verdict comes back as a GradientReport.
// With gtest on the include path (GTEST_INCLUDE_GTEST_GTEST_H_ defined before this header) failures are also
// reported through ADD_FAILURE(), so the reference's TEST_UNARY_GRADIENT-style macros keep working.
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <iostream>
#include <string>

#include "../detail/config.cuh"

namespace xyz_autodiff {
namespace testing {

template <typename T>
XYZ_HD T compute_error_min(T analytical, T numerical) {
    const T abs_error = analytical > numerical ? analytical - numerical : numerical - analytical;
    const T mag = analytical < T(0) ? -analytical : analytical;
    const T rel_error = abs_error / (mag + T(1e-15));
    return abs_error < rel_error ? abs_error : rel_error;
}

struct GradientReport {
    std::string name;
    std::size_t num_tests = 0;
    std::size_t num_failures = 0;  // components over tolerance (plus 1 for a forbidden tolerance / a CUDA error)
    double tolerance = 0.0, delta = 0.0;
    double max_error = 0.0;
    std::size_t max_error_case = 0, max_error_index = 0;
    double max_error_analytical = 0.0, max_error_numerical = 0.0;
    std::string message;
    bool passed() const { return num_failures == 0; }
    explicit operator bool() const { return passed(); }

    void print(std::ostream& os = std::cout) const {  // the reference's summary block
        os << "=== GRADIENT TEST SUMMARY for " << name << " ===\n"
           << "Number of tests: " << num_tests << "\nTolerance: " << tolerance << "\nDelta: " << delta
           << "\nMaximum error: " << max_error << "\nMax error location: test case " << max_error_case << ", input["
           << max_error_index << "]\nMax error values: analytical=" << max_error_analytical
           << ", numerical=" << max_error_numerical << "\n";
        if (max_error > tolerance) os << "RECOMMENDATION: Use tolerance >= " << max_error * 1.1 << " for this operation\n";
        if (!message.empty()) os << message << "\n";
        os << "=========================================" << std::endl;
    }
};

namespace detail {

// counter-based uniform doubles: case c, slot s -> U(lo, hi); the same stream on the host and in the kernel
XYZ_HD double uniform_at(std::uint64_t seed, std::uint64_t c, std::uint64_t s, double lo, double hi) {
    std::uint64_t x = seed + 0x9E3779B97F4A7C15ull * (c * 131ull + s + 1ull);
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    x ^= x >> 31;
    return lo + (hi - lo) * (static_cast<double>(x >> 11) * (1.0 / 9007199254740992.0));
}

struct CaseWorst {  // per-launch reduction target (device memory)
    unsigned long long failures;
    unsigned long long max_error_bits;  // double bits: non-negative doubles order like unsigned integers
    unsigned long long where;           // (case << 8) | component of the current maximum (last writer wins among ties)
};

inline void report_failure(const GradientReport& r) {
#ifdef GTEST_INCLUDE_GTEST_GTEST_H_
    ADD_FAILURE() << r.name << ": " << r.num_failures << " gradient component(s) over tolerance " << r.tolerance
                  << " (max error " << r.max_error << " at case " << r.max_error_case << ", input[" << r.max_error_index
                  << "]: analytical=" << r.max_error_analytical << ", numerical=" << r.max_error_numerical << ") "
                  << r.message;
#else
    (void)r;
#endif
}

}  // namespace detail
}  // namespace testing
}  // namespace xyz_autodiff
