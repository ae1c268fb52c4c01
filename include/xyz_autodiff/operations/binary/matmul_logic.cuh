// xyz_autodiff/operations/binary/matmul_logic.cuh -- small dense matrix product C(a x c) = A(a x b) B(b x c),
// row-major, fully unrolled into registers (no tensor cores: 2x2..4x4 chains are not dense contractions).
// Contract of reference include/xyz_autodiff/operations/binary/matmul_logic.cuh:13-123.  The product is walked
// output element by output element (e = i c + j); per element the adjoint terms go first to row i of A (k ascending),
// then to column j of B (k ascending) -- the order in which the reference issues them.
#pragma once

#include "../../detail/pointwise.cuh"
#include "../operation.cuh"

namespace xyz_autodiff {
namespace detail {
// flat index of (row, column) in a row-major matrix with `Columns` columns
template <std::size_t Columns>
XYZ_HD constexpr std::size_t row_major(std::size_t row, std::size_t column) {
    return row * Columns + column;
}
// two operands of one scalar type holding exactly LeftCount and RightCount components
template <typename Left, std::size_t LeftCount, typename Right, std::size_t RightCount>
concept OperandExtents = BinaryLogicParameterConcept<Left, Right> && (Left::size == LeftCount) && (Right::size == RightCount);
}  // namespace detail

namespace op {

template <std::size_t a, std::size_t b, std::size_t c, typename Input1, typename Input2>
    requires detail::OperandExtents<Input1, a * b, Input2, b * c>
struct MatMulLogic {
    using T = typename Input1::value_type;
    static constexpr std::size_t rows_A = a;
    static constexpr std::size_t cols_A_rows_B = b;
    static constexpr std::size_t cols_B = c;
    static constexpr std::size_t output_size = a * c;
    static constexpr std::size_t outputDim = output_size;
    using Output = Variable<output_size, T>;

    XYZ_HD void forward(Output& product, const Input1& left, const Input2& right) const {
#pragma unroll
        for (std::size_t e = 0; e < output_size; ++e) {
            const std::size_t i = e / c, j = e % c;
            T dot = T(0);
#pragma unroll
            for (std::size_t k = 0; k < b; ++k) dot += left[detail::row_major<b>(i, k)] * right[detail::row_major<c>(k, j)];
            product[e] = dot;
        }
    }

    XYZ_HD void backward(const Output& product, Input1& left, Input2& right) const {
#pragma unroll
        for (std::size_t e = 0; e < output_size; ++e) {
            const std::size_t i = e / c, j = e % c;
            const T upstream = product.grad(e);
#pragma unroll
            for (std::size_t k = 0; k < b; ++k)
                left.add_grad(detail::row_major<b>(i, k), upstream * right[detail::row_major<c>(k, j)]);
#pragma unroll
            for (std::size_t k = 0; k < b; ++k)
                right.add_grad(detail::row_major<c>(k, j), upstream * left[detail::row_major<b>(i, k)]);
        }
    }
};

template <std::size_t a, std::size_t b, std::size_t c, typename Input1, typename Input2>
    requires detail::OperandExtents<Input1, a * b, Input2, b * c>
XYZ_HD auto matmul(Input1& A, Input2& B) {
    return detail::make_binary_node<MatMulLogic<a, b, c, Input1, Input2>>(A, B);
}

// The reference's named shapes, all instances of matmul<rows, inner, columns>: NAME(left, right) with the operand
// lengths LEFT and RIGHT; matvec / vecmat take their shape as template arguments <m, n>.
#define XYZ_AUTODIFF_MATMUL_ALIAS(NAME, SHAPE_PARAMS, LEFT, RIGHT, ...)                   \
    template <SHAPE_PARAMS typename Input1, typename Input2>                              \
        requires detail::OperandExtents<Input1, LEFT, Input2, RIGHT>                      \
    XYZ_HD auto NAME(Input1& left, Input2& right) {                                       \
        return matmul<__VA_ARGS__>(left, right);                                          \
    }
#define XYZ_AUTODIFF_NO_SHAPE
#define XYZ_AUTODIFF_MN std::size_t m, std::size_t n,
XYZ_AUTODIFF_MATMUL_ALIAS(matmul_2x2, XYZ_AUTODIFF_NO_SHAPE, 4, 4, 2, 2, 2)
XYZ_AUTODIFF_MATMUL_ALIAS(matmul_3x3, XYZ_AUTODIFF_NO_SHAPE, 9, 9, 3, 3, 3)
XYZ_AUTODIFF_MATMUL_ALIAS(matmul_4x4, XYZ_AUTODIFF_NO_SHAPE, 16, 16, 4, 4, 4)
XYZ_AUTODIFF_MATMUL_ALIAS(matvec, XYZ_AUTODIFF_MN, m * n, n, m, n, 1)  // A (m x n) times a column vector
XYZ_AUTODIFF_MATMUL_ALIAS(vecmat, XYZ_AUTODIFF_MN, m, m * n, 1, m, n)  // a row vector times A (m x n)
#undef XYZ_AUTODIFF_MN
#undef XYZ_AUTODIFF_NO_SHAPE
#undef XYZ_AUTODIFF_MATMUL_ALIAS

}  // namespace op
}  // namespace xyz_autodiff
