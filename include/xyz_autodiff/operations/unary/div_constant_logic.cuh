// xyz_autodiff/operations/unary/div_constant_logic.cuh -- scaling (division) by a scalar constant held in the Logic.
// Contract of reference include/xyz_autodiff/operations/unary/div_constant_logic.cuh:11-48.
#pragma once

#include "../operation.cuh"

namespace xyz_autodiff {
namespace op {

template <typename Input>
    requires UnaryLogicParameterConcept<Input>
struct DivConstantLogic {
    using T = typename Input::value_type;
    static constexpr std::size_t Dim = Input::size;
    static constexpr std::size_t outputDim = Dim;
    using Output = Variable<Dim, T>;

    T constant_c;

    XYZ_HD explicit DivConstantLogic(T c) : constant_c(c) {}

    XYZ_HD void forward(Output& y, const Input& x) const {
#pragma unroll
        for (std::size_t i = 0; i < Dim; ++i) y[i] = x[i] / constant_c;
    }

    XYZ_HD void backward(const Output& y, Input& x) const {
#pragma unroll
        for (std::size_t i = 0; i < Dim; ++i) x.add_grad(i, y.grad(i) / constant_c);
    }
};

template <typename Input>
    requires UnaryLogicParameterConcept<Input>
XYZ_HD auto div_constant(Input& x, typename Input::value_type constant) {
    using Logic = DivConstantLogic<Input>;
    return UnaryOperation<Logic::outputDim, Logic, Input>(Logic(constant), x);
}

}  // namespace op
}  // namespace xyz_autodiff
