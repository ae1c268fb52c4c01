// xyz_autodiff/operations/unary/neg_logic.cuh -- element-wise negation.
// Contract of reference include/xyz_autodiff/operations/unary/neg_logic.cuh:13-49.
#pragma once

#include "../math.cuh"
#include "../operation.cuh"

namespace xyz_autodiff {
namespace op {

template <std::size_t Dim>
struct NegLogic {
    static constexpr std::size_t outputDim = Dim;

    template <typename Output, typename Input>
    XYZ_HD void forward(Output& y, const Input& x) const {
        using T = typename Input::value_type;
#pragma unroll
        for (std::size_t i = 0; i < Dim; ++i) {
            const T v = x[i];
            y[i] = -v;
        }
    }

    template <typename Output, typename Input>
    XYZ_HD void backward(const Output& y, Input& x) const {
        using T = typename Input::value_type;
#pragma unroll
        for (std::size_t i = 0; i < Dim; ++i) {
            const T g = y.grad(i);
            x.add_grad(i, -g);
        }
    }
};

template <std::size_t Dim, DifferentiableVariableConcept Input>
    requires(Input::size == Dim)
XYZ_HD auto neg(Input& x) {
    return UnaryOperation<Dim, NegLogic<Dim>, Input>(NegLogic<Dim>{}, x);
}

template <DifferentiableVariableConcept Input>
XYZ_HD auto neg(Input& x) {
    return neg<Input::size>(x);
}

}  // namespace op
}  // namespace xyz_autodiff
