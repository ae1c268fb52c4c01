// xyz_autodiff/operations/unary/cos_logic.cuh -- element-wise cosine.
// Contract of reference include/xyz_autodiff/operations/unary/cos_logic.cuh:13-51.
#pragma once

#include "../math.cuh"
#include "../operation.cuh"

namespace xyz_autodiff {
namespace op {

template <std::size_t Dim>
struct CosLogic {
    static constexpr std::size_t outputDim = Dim;

    template <typename Output, typename Input>
    XYZ_HD void forward(Output& y, const Input& x) const {
        using T = typename Input::value_type;
#pragma unroll
        for (std::size_t i = 0; i < Dim; ++i) {
            const T v = x[i];
            y[i] = math::cos(v);
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
            x.add_grad(i, g * (-math::sin(v)));
        }
    }
};

template <std::size_t Dim, DifferentiableVariableConcept Input>
    requires(Input::size == Dim)
XYZ_HD auto cos(Input& x) {
    return UnaryOperation<Dim, CosLogic<Dim>, Input>(CosLogic<Dim>{}, x);
}

template <DifferentiableVariableConcept Input>
XYZ_HD auto cos(Input& x) {
    return cos<Input::size>(x);
}

}  // namespace op
}  // namespace xyz_autodiff
