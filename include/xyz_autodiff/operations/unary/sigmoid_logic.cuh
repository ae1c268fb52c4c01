// xyz_autodiff/operations/unary/sigmoid_logic.cuh -- element-wise logistic sigmoid 1 / (1 + exp(-x)).
// Contract of reference include/xyz_autodiff/operations/unary/sigmoid_logic.cuh:13-52.
#pragma once

#include "../math.cuh"
#include "../operation.cuh"

namespace xyz_autodiff {
namespace op {

template <std::size_t Dim>
struct SigmoidLogic {
    static constexpr std::size_t outputDim = Dim;

    template <typename Output, typename Input>
    XYZ_HD void forward(Output& y, const Input& x) const {
        using T = typename Input::value_type;
#pragma unroll
        for (std::size_t i = 0; i < Dim; ++i) {
            const T v = x[i];
            y[i] = math::sigmoid(v);
        }
    }

    // the local derivative is recomputed from the INPUT (nothing is cached between the passes)
    template <typename Output, typename Input>
    XYZ_HD void backward(const Output& y, Input& x) const {
        using T = typename Input::value_type;
#pragma unroll
        for (std::size_t i = 0; i < Dim; ++i) {
            const T v = x[i];
            const T g = y.grad(i);
            const T s = math::sigmoid(v);
            x.add_grad(i, g * (s * (T(1) - s)));
        }
    }
};

template <std::size_t Dim, DifferentiableVariableConcept Input>
    requires(Input::size == Dim)
XYZ_HD auto sigmoid(Input& x) {
    return UnaryOperation<Dim, SigmoidLogic<Dim>, Input>(SigmoidLogic<Dim>{}, x);
}

template <DifferentiableVariableConcept Input>
XYZ_HD auto sigmoid(Input& x) {
    return sigmoid<Input::size>(x);
}

}  // namespace op
}  // namespace xyz_autodiff
