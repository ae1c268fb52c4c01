// xyz_autodiff/operations/unary/squared_logic.cuh -- element-wise square.
// Contract of reference include/xyz_autodiff/operations/unary/squared_logic.cuh:11-45, including its
// `g * 2.0 * x` backward: the double literal promotes the product, so an fp32 graph rounds once from
// the exact double product.
#pragma once

#include "../operation.cuh"

namespace xyz_autodiff {
namespace op {

template <std::size_t Dim>
struct SquaredLogic {
    static constexpr std::size_t outputDim = Dim;

    template <typename Input>
    XYZ_HD void forward(Variable<Dim, typename Input::value_type>& y, const Input& x) const {
#pragma unroll
        for (std::size_t i = 0; i < Dim; ++i) {
            const typename Input::value_type v = x[i];
            y[i] = v * v;
        }
    }

    template <typename Input>
    XYZ_HD void backward(const Variable<Dim, typename Input::value_type>& y, Input& x) const {
        using T = typename Input::value_type;
#pragma unroll
        for (std::size_t i = 0; i < Dim; ++i) x.add_grad(i, static_cast<T>(y.grad(i) * 2.0 * x[i]));
    }
};

template <typename Input>
    requires UnaryLogicParameterConcept<Input>
XYZ_HD auto squared(Input& x) {
    using Logic = SquaredLogic<Input::size>;
    return UnaryOperation<Input::size, Logic, Input>(Logic{}, x);
}

}  // namespace op
}  // namespace xyz_autodiff
