// xyz_autodiff/operations/unary/const_array_add_logic.cuh -- y[i] = x[i] + c[i] with c a constant array
// held BY REFERENCE (the array must outlive the node).
// Contract of reference include/xyz_autodiff/operations/unary/const_array_add_logic.cuh:14-48.
#pragma once

#include "../operation.cuh"
#include "const_array_concepts.cuh"

namespace xyz_autodiff {
namespace op {

template <std::size_t Dim, typename ConstantArray>
    requires ArrayLikeConcept<ConstantArray>
struct ConstArrayAddLogic {
    static constexpr std::size_t outputDim = Dim;

    const ConstantArray& constant_array;

    XYZ_HD explicit ConstArrayAddLogic(const ConstantArray& c) : constant_array(c) {}

    template <typename Output, typename Input>
    XYZ_HD void forward(Output& y, const Input& x) const {
#pragma unroll
        for (std::size_t i = 0; i < Dim; ++i) y[i] = x[i] + constant_array[i];
    }

    template <typename Output, typename Input>
    XYZ_HD void backward(const Output& y, Input& x) const {
#pragma unroll
        for (std::size_t i = 0; i < Dim; ++i) x.add_grad(i, y.grad(i));
    }
};

template <DifferentiableVariableConcept Input, typename ConstantArray>
    requires ArrayLikeConcept<ConstantArray>
XYZ_HD auto const_add(Input& x, const ConstantArray& c) {
    using Logic = ConstArrayAddLogic<Input::size, ConstantArray>;
    return UnaryOperation<Input::size, Logic, Input>(Logic(c), x);
}

}  // namespace op
}  // namespace xyz_autodiff
