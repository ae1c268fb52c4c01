// xyz_autodiff/operations/binary/sub_logic.cuh -- element-wise difference of two equally sized operands.
// Contract of reference include/xyz_autodiff/operations/binary/sub_logic.cuh:11-50.
#pragma once

#include "../operation.cuh"

namespace xyz_autodiff {
namespace op {

template <typename Input1, typename Input2>
    requires BinaryLogicParameterConcept<Input1, Input2> && (Input1::size == Input2::size)
struct SubLogic {
    using T = typename Input1::value_type;
    static constexpr std::size_t Dim = Input1::size;
    static constexpr std::size_t outputDim = Dim;
    using Output = Variable<Dim, T>;

    XYZ_HD void forward(Output& y, const Input1& a, const Input2& b) const {
#pragma unroll
        for (std::size_t i = 0; i < Dim; ++i) y[i] = a[i] - b[i];
    }

    XYZ_HD void backward(const Output& y, Input1& a, Input2& b) const {
#pragma unroll
        for (std::size_t i = 0; i < Dim; ++i) {
            const T g = y.grad(i);
            a.add_grad(i, g);
            b.add_grad(i, -g);
        }
    }
};

template <typename Input1, typename Input2>
    requires BinaryLogicParameterConcept<Input1, Input2>
XYZ_HD auto sub(Input1& a, Input2& b) {
    using Logic = SubLogic<Input1, Input2>;
    return BinaryOperation<Logic::outputDim, Logic, Input1, Input2>(Logic{}, a, b);
}

}  // namespace op
}  // namespace xyz_autodiff
