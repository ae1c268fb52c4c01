// xyz_autodiff/diagonal_matrix_view.cuh -- an N-vector variable seen as diag(v): N stored values,
// N x N logical shape.  Contract of reference include/xyz_autodiff/diagonal_matrix_view.cuh:9-101.
#pragma once

#include "detail/config.cuh"
#include "variable.cuh"

namespace xyz_autodiff {

template <typename T, std::size_t N, typename VariableType = VariableRef<N, T>>
class DiagonalMatrixView {
public:
    using value_type = T;
    static constexpr std::size_t rows = N;
    static constexpr std::size_t cols = N;
    static constexpr std::size_t size = N;  // stored (diagonal) entries

    XYZ_HD DiagonalMatrixView(VariableType& diagonal) : diag_(diagonal) {}
    DiagonalMatrixView(const DiagonalMatrixView&) = default;
    DiagonalMatrixView& operator=(const DiagonalMatrixView&) = delete;  // a reference cannot be re-seated
    DiagonalMatrixView& operator=(DiagonalMatrixView&&) = delete;

    // variable interface over the diagonal
    XYZ_HD T* data() const { return diag_.data(); }
    XYZ_HD T* grad() const { return diag_.grad(); }
    XYZ_HD T& operator[](std::size_t i) const { return diag_[i]; }
    XYZ_HD const T& grad(std::size_t i) const { return diag_.grad(i); }
    XYZ_HD void add_grad(std::size_t i, T value) const { diag_.add_grad(i, value); }
    XYZ_HD void zero_grad() const { diag_.zero_grad(); }

    // matrix interface: by value, zero off the diagonal
    XYZ_HD T operator()(std::size_t r, std::size_t c) { return r == c ? diag_[r] : T{0}; }
    XYZ_HD constexpr T operator()(std::size_t r, std::size_t c) const { return r == c ? diag_[r] : T{0}; }
    XYZ_HD DiagonalMatrixView transpose() const { return *this; }

    XYZ_HD VariableType& underlying_variable() { return diag_; }
    XYZ_HD const VariableType& underlying_variable() const { return diag_; }

private:
    VariableType& diag_;
};

template <std::size_t N, typename T>
XYZ_HD auto make_diagonal_matrix_view(VariableRef<N, T>& v) {
    return DiagonalMatrixView<T, N, VariableRef<N, T>>(v);
}
template <std::size_t N, typename T>
XYZ_HD auto make_diagonal_matrix_view(Variable<N, T>& v) {
    return DiagonalMatrixView<T, N, Variable<N, T>>(v);
}

}  // namespace xyz_autodiff
