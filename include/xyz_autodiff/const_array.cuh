// xyz_autodiff/const_array.cuh -- ConstArray<T, N>: a plain value array (no adjoint) with element-wise
// arithmetic against itself and against anything ConstArrayLike (Variables, VariableRefs, operation
// nodes).  Used as the pixel type of the splat example (PixelOutput = ConstArray<float, 3>).
// Contract of reference include/xyz_autodiff/const_array.cuh:7-223.
#pragma once

#include "concept/const_array.cuh"
#include "detail/config.cuh"

namespace xyz_autodiff {

template <typename T, int N>
struct ConstArray {
    using value_type = T;
    static constexpr std::size_t size = N;

    T data[N];

    constexpr ConstArray() = default;
    XYZ_HD constexpr ConstArray(const T (&init)[N]) {
        for (int i = 0; i < N; ++i) data[i] = init[i];
    }

    XYZ_HD constexpr T& operator[](std::size_t i) noexcept { return data[i]; }
    XYZ_HD constexpr const T& operator[](std::size_t i) const noexcept { return data[i]; }

#define XYZ_CONST_ARRAY_COMPOUND(SYMBOL)                                              \
    XYZ_HD constexpr ConstArray& operator SYMBOL(const ConstArray& other) {           \
        for (int i = 0; i < N; ++i) data[i] SYMBOL other.data[i];                     \
        return *this;                                                                 \
    }                                                                                 \
    template <ConstArrayLike Other>                                                   \
        requires ConstArrayCompatible<ConstArray, Other> && (Other::size == N)        \
    XYZ_HD constexpr ConstArray& operator SYMBOL(const Other& other) {                \
        for (int i = 0; i < N; ++i) data[i] SYMBOL other[i];                          \
        return *this;                                                                 \
    }
    XYZ_CONST_ARRAY_COMPOUND(+=)
    XYZ_CONST_ARRAY_COMPOUND(-=)
    XYZ_CONST_ARRAY_COMPOUND(*=)
    XYZ_CONST_ARRAY_COMPOUND(/=)
#undef XYZ_CONST_ARRAY_COMPOUND
};

// array (op) array, array (op) array-like, array-like (op) array -> a new ConstArray
#define XYZ_CONST_ARRAY_BINARY(SYMBOL)                                                                     \
    template <typename T, int N>                                                                           \
    XYZ_HD constexpr ConstArray<T, N> operator SYMBOL(const ConstArray<T, N>& a, const ConstArray<T, N>& b) { \
        ConstArray<T, N> r{};                                                                              \
        for (int i = 0; i < N; ++i) r[i] = a[i] SYMBOL b[i];                                               \
        return r;                                                                                          \
    }                                                                                                      \
    template <typename T, int N, ConstArrayLike Other>                                                     \
        requires ConstArrayCompatible<ConstArray<T, N>, Other> && (Other::size == N)                       \
    XYZ_HD constexpr ConstArray<T, N> operator SYMBOL(const ConstArray<T, N>& a, const Other& b) {         \
        ConstArray<T, N> r{};                                                                              \
        for (int i = 0; i < N; ++i) r[i] = a[i] SYMBOL b[i];                                               \
        return r;                                                                                          \
    }                                                                                                      \
    template <typename T, int N, ConstArrayLike Other>                                                     \
        requires ConstArrayCompatible<ConstArray<T, N>, Other> && (Other::size == N)                       \
    XYZ_HD constexpr ConstArray<T, N> operator SYMBOL(const Other& a, const ConstArray<T, N>& b) {         \
        ConstArray<T, N> r{};                                                                              \
        for (int i = 0; i < N; ++i) r[i] = a[i] SYMBOL b[i];                                               \
        return r;                                                                                          \
    }
XYZ_CONST_ARRAY_BINARY(+)
XYZ_CONST_ARRAY_BINARY(-)
XYZ_CONST_ARRAY_BINARY(*)
XYZ_CONST_ARRAY_BINARY(/)
#undef XYZ_CONST_ARRAY_BINARY

}  // namespace xyz_autodiff
