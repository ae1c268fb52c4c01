// xyz_autodiff/operations/unary/sum_logic.cuh -- sum of an InputDim-vector (scalar output).
// Contract of reference include/xyz_autodiff/operations/unary/sum_logic.cuh:13-53.
#pragma once

#include "../math.cuh"
#include "../operation.cuh"

namespace xyz_autodiff {
namespace op {

template <std::size_t InputDim>
struct SumLogic {
    static constexpr std::size_t outputDim = 1;

    template <typename Output, typename Input>
    XYZ_HD void forward(Output& y, const Input& x) const {
        using T = typename Input::value_type;
        T acc = T(0);
#pragma unroll
        for (std::size_t i = 0; i < InputDim; ++i) acc += x[i];
        y[0] = acc;
    }

    template <typename Output, typename Input>
    XYZ_HD void backward(const Output& y, Input& x) const {
        using T = typename Input::value_type;
        const T g = y.grad(0);
#pragma unroll
        for (std::size_t i = 0; i < InputDim; ++i) x.add_grad(i, g);
    }
};

template <std::size_t Dim, DifferentiableVariableConcept Input>
    requires(Input::size == Dim)
XYZ_HD auto sum(Input& x) {
    return UnaryOperation<1, SumLogic<Dim>, Input>(SumLogic<Dim>{}, x);
}

template <DifferentiableVariableConcept Input>
XYZ_HD auto sum(Input& x) {
    return sum<Input::size>(x);
}

}  // namespace op
}  // namespace xyz_autodiff
