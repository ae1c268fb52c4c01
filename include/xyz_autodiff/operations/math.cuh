// xyz_autodiff/operations/math.cuh -- the math dispatcher: one spelling per function for float and
// double, host and device.  Same surface as reference include/xyz_autodiff/operations/math.cuh:14-298.
//
// float arguments go to the single-precision C functions (expf, sinf, ...), double arguments to the
// double-precision ones, in BOTH compilation modes.  On the device, building a translation unit with
// -use_fast_math maps the float ones to MUFU (ex2.approx / sin.approx / rcp.approx, FTZ) exactly as it
// does for the reference's training app; without it they are the IEEE-accurate libdevice versions.
#pragma once

#include <cmath>
#include <type_traits>

#include "../detail/config.cuh"

namespace xyz_autodiff {
namespace math {

#define XYZ_MATH_UNARY(NAME, F32, F64)                      \
    template <typename T>                                   \
    XYZ_HD T NAME(const T& x) {                             \
        if constexpr (std::is_same_v<T, float>) {           \
            return F32(x);                                  \
        } else {                                            \
            static_assert(std::is_same_v<T, double>);       \
            return F64(x);                                  \
        }                                                   \
    }

XYZ_MATH_UNARY(exp, ::expf, ::exp)
XYZ_MATH_UNARY(log, ::logf, ::log)
XYZ_MATH_UNARY(sin, ::sinf, ::sin)
XYZ_MATH_UNARY(cos, ::cosf, ::cos)
XYZ_MATH_UNARY(tan, ::tanf, ::tan)
XYZ_MATH_UNARY(sinh, ::sinhf, ::sinh)
XYZ_MATH_UNARY(cosh, ::coshf, ::cosh)
XYZ_MATH_UNARY(tanh, ::tanhf, ::tanh)
XYZ_MATH_UNARY(sqrt, ::sqrtf, ::sqrt)
XYZ_MATH_UNARY(abs, ::fabsf, ::fabs)
#undef XYZ_MATH_UNARY

template <typename T>
XYZ_HD T pow(const T& base, const T& exponent) {
    if constexpr (std::is_same_v<T, float>) {
        return ::powf(base, exponent);
    } else {
        return ::pow(base, exponent);
    }
}

template <typename T>
XYZ_HD T max(const T& a, const T& b) {
    if constexpr (std::is_same_v<T, float>) {
        return ::fmaxf(a, b);
    } else {
        return ::fmax(a, b);
    }
}

template <typename T>
XYZ_HD T min(const T& a, const T& b) {
    if constexpr (std::is_same_v<T, float>) {
        return ::fminf(a, b);
    } else {
        return ::fmin(a, b);
    }
}

// ---- activations (reference math.cuh:200-243) -----------------------------------------------------
template <typename T>
XYZ_HD T sigmoid(const T& x) {
    return T(1) / (T(1) + exp(-x));
}

template <typename T>
XYZ_HD T relu(const T& x) {
    return x > T(0) ? x : T(0);
}

template <typename T>
XYZ_HD T leaky_relu(const T& x, const T& alpha = T(0.01)) {
    return x > T(0) ? x : alpha * x;
}

template <typename T>
XYZ_HD T elu(const T& x, const T& alpha = T(1)) {
    return x >= T(0) ? x : alpha * (exp(x) - T(1));
}

template <typename T>
XYZ_HD T softplus(const T& x) {
    return log(T(1) + exp(x));
}

template <typename T>
XYZ_HD T swish(const T& x) {
    return x * sigmoid(x);
}

template <typename T>
XYZ_HD T gelu(const T& x) {  // tanh approximation, constants of reference math.cuh:236-243
    const T k = T(0.7978845608028654);
    const T c = T(0.044715);
    return T(0.5) * x * (T(1) + tanh(k * (x + c * pow(x, T(3)))));
}

}  // namespace math
}  // namespace xyz_autodiff
