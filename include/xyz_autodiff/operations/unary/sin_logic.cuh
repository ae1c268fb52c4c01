// xyz_autodiff/operations/unary/sin_logic.cuh -- element-wise sine.
// Contract of reference include/xyz_autodiff/operations/unary/sin_logic.cuh:13-51.
#pragma once

#include "../math.cuh"
#include "../operation.cuh"

namespace xyz_autodiff {
namespace op {

template <std::size_t Dim>
struct SinLogic {
    static constexpr std::size_t outputDim = Dim;

    template <typename Output, typename Input>
    XYZ_HD void forward(Output& y, const Input& x) const {
        using T = typename Input::value_type;
#pragma unroll
        for (std::size_t i = 0; i < Dim; ++i) {
            const T v = x[i];
            y[i] = math::sin(v);
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
            x.add_grad(i, g * math::cos(v));
        }
    }
};

template <std::size_t Dim, DifferentiableVariableConcept Input>
    requires(Input::size == Dim)
XYZ_HD auto sin(Input& x) {
    return UnaryOperation<Dim, SinLogic<Dim>, Input>(SinLogic<Dim>{}, x);
}

template <DifferentiableVariableConcept Input>
XYZ_HD auto sin(Input& x) {
    return sin<Input::size>(x);
}

}  // namespace op
}  // namespace xyz_autodiff
