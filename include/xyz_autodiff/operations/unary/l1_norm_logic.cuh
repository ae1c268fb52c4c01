// xyz_autodiff/operations/unary/l1_norm_logic.cuh -- sum of absolute values of an InputDim-vector (scalar output).
// Contract of reference include/xyz_autodiff/operations/unary/l1_norm_logic.cuh:13-54.
#pragma once

#include "../math.cuh"
#include "../operation.cuh"

namespace xyz_autodiff {
namespace op {

template <std::size_t InputDim>
struct L1NormLogic {
    static constexpr std::size_t outputDim = 1;

    template <typename Output, typename Input>
    XYZ_HD void forward(Output& y, const Input& x) const {
        using T = typename Input::value_type;
        T acc = T(0);
#pragma unroll
        for (std::size_t i = 0; i < InputDim; ++i) acc += math::abs(x[i]);
        y[0] = acc;
    }

    template <typename Output, typename Input>
    XYZ_HD void backward(const Output& y, Input& x) const {
        using T = typename Input::value_type;
        const T g = y.grad(0);
#pragma unroll
        for (std::size_t i = 0; i < InputDim; ++i) {
            const T v = x[i];
            const T s = v > T(0) ? T(1) : (v < T(0) ? T(-1) : T(0));  // subgradient 0 at the kink
            x.add_grad(i, g * s);
        }
    }
};

template <std::size_t Dim, DifferentiableVariableConcept Input>
    requires(Input::size == Dim)
XYZ_HD auto l1_norm(Input& x) {
    return UnaryOperation<1, L1NormLogic<Dim>, Input>(L1NormLogic<Dim>{}, x);
}

template <DifferentiableVariableConcept Input>
XYZ_HD auto l1_norm(Input& x) {
    return l1_norm<Input::size>(x);
}

}  // namespace op
}  // namespace xyz_autodiff
