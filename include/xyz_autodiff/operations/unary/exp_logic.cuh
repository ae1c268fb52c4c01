// xyz_autodiff/operations/unary/exp_logic.cuh -- element-wise exponential.
// Contract of reference include/xyz_autodiff/operations/unary/exp_logic.cuh:13-51.
#pragma once

#include "../math.cuh"
#include "../operation.cuh"

namespace xyz_autodiff {
namespace op {

template <std::size_t Dim>
struct ExpLogic {
    static constexpr std::size_t outputDim = Dim;

    template <typename Output, typename Input>
    XYZ_HD void forward(Output& y, const Input& x) const {
        using T = typename Input::value_type;
#pragma unroll
        for (std::size_t i = 0; i < Dim; ++i) {
            const T v = x[i];
            y[i] = math::exp(v);
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
            x.add_grad(i, g * math::exp(v));
        }
    }
};

template <std::size_t Dim, DifferentiableVariableConcept Input>
    requires(Input::size == Dim)
XYZ_HD auto exp(Input& x) {
    return UnaryOperation<Dim, ExpLogic<Dim>, Input>(ExpLogic<Dim>{}, x);
}

template <DifferentiableVariableConcept Input>
XYZ_HD auto exp(Input& x) {
    return exp<Input::size>(x);
}

}  // namespace op
}  // namespace xyz_autodiff
