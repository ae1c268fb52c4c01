// xyz_autodiff/operations/unary/l2_norm_logic.cuh -- Euclidean norm of an InputDim-vector (scalar output).
// Contract of reference include/xyz_autodiff/operations/unary/l2_norm_logic.cuh:13-57.
#pragma once

#include "../math.cuh"
#include "../operation.cuh"

namespace xyz_autodiff {
namespace op {

template <std::size_t InputDim>
struct L2NormLogic {
    static constexpr std::size_t outputDim = 1;

    template <typename Output, typename Input>
    XYZ_HD void forward(Output& y, const Input& x) const {
        using T = typename Input::value_type;
        T acc = T(0);
#pragma unroll
        for (std::size_t i = 0; i < InputDim; ++i) acc += x[i] * x[i];
        y[0] = math::sqrt(acc);
    }

    template <typename Output, typename Input>
    XYZ_HD void backward(const Output& y, Input& x) const {
        using T = typename Input::value_type;
        const T g = y.grad(0);
        const T norm = y[0];
        if (norm > T(1e-8)) {  // no adjoint at (numerically) zero vectors, like the reference
#pragma unroll
            for (std::size_t i = 0; i < InputDim; ++i) x.add_grad(i, g * x[i] / norm);
        }
    }
};

template <std::size_t Dim, DifferentiableVariableConcept Input>
    requires(Input::size == Dim)
XYZ_HD auto l2_norm(Input& x) {
    return UnaryOperation<1, L2NormLogic<Dim>, Input>(L2NormLogic<Dim>{}, x);
}

template <DifferentiableVariableConcept Input>
XYZ_HD auto l2_norm(Input& x) {
    return l2_norm<Input::size>(x);
}

}  // namespace op
}  // namespace xyz_autodiff
