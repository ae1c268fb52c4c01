// xyz_autodiff/dense_matrix.cuh -- DenseMatrix<T, R, C>: an owning row-major matrix that is also a
// (differentiable) variable of size R*C, plus the VALUE-ONLY product of any two matrix views.
// Contract of reference include/xyz_autodiff/dense_matrix.cuh:11-125.  (For a differentiable product
// use op::matmul from operations/binary/matmul_logic.cuh.)
#pragma once

#include "concept/matrix.cuh"
#include "detail/config.cuh"

namespace xyz_autodiff {

template <typename T, std::size_t Rows, std::size_t Cols>
class DenseMatrix {
public:
    using value_type = T;
    static constexpr std::size_t rows = Rows;
    static constexpr std::size_t cols = Cols;
    static constexpr std::size_t size = Rows * Cols;

    XYZ_HD constexpr DenseMatrix() {
        for (std::size_t i = 0; i < size; ++i) {
            v_[i] = T{};
            g_[i] = T{};
        }
    }

    // variable interface (flat, row-major)
    XYZ_HD T* data() { return v_; }
    XYZ_HD const T* data() const { return v_; }
    XYZ_HD T* grad() { return g_; }
    XYZ_HD const T* grad() const { return g_; }
    XYZ_HD T& operator[](std::size_t i) { return v_[i]; }
    XYZ_HD const T& operator[](std::size_t i) const { return v_[i]; }
    XYZ_HD const T& grad(std::size_t i) const { return g_[i]; }
    XYZ_HD void add_grad(std::size_t i, T value) { g_[i] += value; }
    XYZ_HD void zero_grad() {
        for (std::size_t i = 0; i < size; ++i) g_[i] = T{};
    }

    // matrix interface
    XYZ_HD T& operator()(std::size_t r, std::size_t c) { return v_[r * Cols + c]; }
    XYZ_HD const T& operator()(std::size_t r, std::size_t c) const { return v_[r * Cols + c]; }

    // a transposed COPY (values only; adjoints of the copy start at zero)
    XYZ_HD DenseMatrix<T, Cols, Rows> transpose() const {
        DenseMatrix<T, Cols, Rows> t;
        for (std::size_t r = 0; r < Rows; ++r)
            for (std::size_t c = 0; c < Cols; ++c) t(c, r) = (*this)(r, c);
        return t;
    }

private:
    T v_[Rows * Cols];
    T g_[Rows * Cols];
};

// value-only product of two matrix views; fixed trip counts, fully unrolled by the compiler.
// Summation order per entry: k = K-1 innermost first, i.e. a(i,0)b(0,j) + (a(i,1)b(1,j) + (... + 0)),
// the association of the reference's recursive template (dense_matrix.cuh:86-96).
template <typename A, typename B>
    requires MatrixViewConcept<A> && MatrixViewConcept<B> && (A::cols == B::rows)
XYZ_HD constexpr DenseMatrix<typename A::value_type, A::rows, B::cols> operator*(const A& a, const B& b) {
    using V = typename A::value_type;
    DenseMatrix<V, A::rows, B::cols> out;
    for (std::size_t i = 0; i < A::rows; ++i)
        for (std::size_t j = 0; j < B::cols; ++j) {
            V acc = V{0};
            for (std::size_t k = A::cols; k-- > 0;) acc = a(i, k) * b(k, j) + acc;
            out[i * B::cols + j] = acc;
        }
    return out;
}

}  // namespace xyz_autodiff
