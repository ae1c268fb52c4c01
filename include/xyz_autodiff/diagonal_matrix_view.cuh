// xyz_autodiff/diagonal_matrix_view.cuh -- an N-vector variable seen as diag(v): N stored values,
// N x N logical shape.  Contract of reference include/xyz_autodiff/diagonal_matrix_view.cuh:9-101.
#pragma once

#include "detail/config.cuh"
#include "detail/variable_facade.cuh"
#include "variable.cuh"

namespace xyz_autodiff {

template <typename T, std::size_t N, typename VariableType = VariableRef<N, T>>
class DiagonalMatrixView : public detail::VariableFacade<DiagonalMatrixView<T, N, VariableType>> {
public:
    using value_type = T;
    static constexpr std::size_t rows = N, cols = N;
    static constexpr std::size_t size = N;  // stored entries: the diagonal (the variable interface is inherited)

    XYZ_HD DiagonalMatrixView(VariableType& diagonal_entries) : entries_(&diagonal_entries) {}

    // matrix interface: entries by value, exact zeros off the diagonal; diag(v) is its own transpose
    XYZ_HD constexpr T operator()(std::size_t r, std::size_t c) const { return r != c ? T{0} : (*entries_)[r]; }
    XYZ_HD DiagonalMatrixView transpose() const { return *this; }

    XYZ_HD VariableType& underlying_variable() { return *entries_; }
    XYZ_HD const VariableType& underlying_variable() const { return *entries_; }
    XYZ_HD VariableType& stored() const { return *entries_; }

    DiagonalMatrixView(const DiagonalMatrixView&) = default;
    DiagonalMatrixView& operator=(const DiagonalMatrixView&) = delete;  // a view is bound once, like a reference
    DiagonalMatrixView& operator=(DiagonalMatrixView&&) = delete;

private:
    VariableType* const entries_;
};

template <std::size_t N, typename T>
XYZ_HD auto make_diagonal_matrix_view(VariableRef<N, T>& diagonal_entries) {
    return DiagonalMatrixView<T, N, VariableRef<N, T>>(diagonal_entries);
}
template <std::size_t N, typename T>
XYZ_HD auto make_diagonal_matrix_view(Variable<N, T>& diagonal_entries) {
    return DiagonalMatrixView<T, N, Variable<N, T>>(diagonal_entries);
}

}  // namespace xyz_autodiff
