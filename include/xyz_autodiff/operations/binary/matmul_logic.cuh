// xyz_autodiff/operations/binary/matmul_logic.cuh -- small dense matrix product C(a x c) = A(a x b) B(b x c),
// row-major, fully unrolled into registers (no tensor cores: 2x2..4x4 chains are not dense
// contractions).  Contract of reference include/xyz_autodiff/operations/binary/matmul_logic.cuh:13-123;
// adjoint terms are issued in the reference's order: for each (i, j): all k into A, then all k into B.
#pragma once

#include "../operation.cuh"

namespace xyz_autodiff {
namespace op {

template <std::size_t a, std::size_t b, std::size_t c, typename Input1, typename Input2>
    requires BinaryLogicParameterConcept<Input1, Input2> && (Input1::size == a * b) && (Input2::size == b * c)
struct MatMulLogic {
    using T = typename Input1::value_type;
    static constexpr std::size_t rows_A = a;
    static constexpr std::size_t cols_A_rows_B = b;
    static constexpr std::size_t cols_B = c;
    static constexpr std::size_t output_size = a * c;
    static constexpr std::size_t outputDim = output_size;
    using Output = Variable<output_size, T>;

    XYZ_HD void forward(Output& C, const Input1& A, const Input2& B) const {
#pragma unroll
        for (std::size_t i = 0; i < a; ++i) {
#pragma unroll
            for (std::size_t j = 0; j < c; ++j) {
                T acc = T(0);
#pragma unroll
                for (std::size_t k = 0; k < b; ++k) acc += A[i * b + k] * B[k * c + j];
                C[i * c + j] = acc;
            }
        }
    }

    XYZ_HD void backward(const Output& C, Input1& A, Input2& B) const {
#pragma unroll
        for (std::size_t i = 0; i < a; ++i) {
#pragma unroll
            for (std::size_t j = 0; j < c; ++j) {
                const T g = C.grad(i * c + j);
#pragma unroll
                for (std::size_t k = 0; k < b; ++k) A.add_grad(i * b + k, g * B[k * c + j]);
#pragma unroll
                for (std::size_t k = 0; k < b; ++k) B.add_grad(k * c + j, g * A[i * b + k]);
            }
        }
    }
};

template <std::size_t a, std::size_t b, std::size_t c, typename Input1, typename Input2>
    requires BinaryLogicParameterConcept<Input1, Input2> && (Input1::size == a * b) && (Input2::size == b * c)
XYZ_HD auto matmul(Input1& A, Input2& B) {
    using Logic = MatMulLogic<a, b, c, Input1, Input2>;
    return BinaryOperation<Logic::outputDim, Logic, Input1, Input2>(Logic{}, A, B);
}

template <typename Input1, typename Input2>
    requires BinaryLogicParameterConcept<Input1, Input2> && (Input1::size == 4) && (Input2::size == 4)
XYZ_HD auto matmul_2x2(Input1& A, Input2& B) {
    return matmul<2, 2, 2>(A, B);
}

template <typename Input1, typename Input2>
    requires BinaryLogicParameterConcept<Input1, Input2> && (Input1::size == 9) && (Input2::size == 9)
XYZ_HD auto matmul_3x3(Input1& A, Input2& B) {
    return matmul<3, 3, 3>(A, B);
}

template <typename Input1, typename Input2>
    requires BinaryLogicParameterConcept<Input1, Input2> && (Input1::size == 16) && (Input2::size == 16)
XYZ_HD auto matmul_4x4(Input1& A, Input2& B) {
    return matmul<4, 4, 4>(A, B);
}

// A (m x n) times a column vector
template <std::size_t m, std::size_t n, typename Input1, typename Input2>
    requires BinaryLogicParameterConcept<Input1, Input2> && (Input1::size == m * n) && (Input2::size == n)
XYZ_HD auto matvec(Input1& A, Input2& x) {
    return matmul<m, n, 1>(A, x);
}

// a row vector times A (m x n)
template <std::size_t m, std::size_t n, typename Input1, typename Input2>
    requires BinaryLogicParameterConcept<Input1, Input2> && (Input1::size == m) && (Input2::size == m * n)
XYZ_HD auto vecmat(Input1& x, Input2& A) {
    return matmul<1, m, n>(x, A);
}

}  // namespace op
}  // namespace xyz_autodiff
