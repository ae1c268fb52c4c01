// xyz_autodiff/operations/unary/to_rotation_matrix_logic.cuh -- quaternion (x, y, z, w) -> row-major 3x3
// rotation matrix (the quaternion is NOT normalised here), with the explicit 4 x 9 reverse rule.
// Contract of reference include/xyz_autodiff/operations/unary/to_rotation_matrix_logic.cuh:12-124.
#pragma once

#include "../operation.cuh"

namespace xyz_autodiff {
namespace op {

template <std::size_t InputDim>
    requires(InputDim == 4)
struct QuaternionToRotationMatrixLogic {
    static constexpr std::size_t outputDim = 9;

    template <typename Output, typename Input>
    XYZ_HD void forward(Output& R, const Input& q) const {
        using T = typename Input::value_type;
        const T x = q[0], y = q[1], z = q[2], w = q[3];
        R[0] = T(1) - T(2) * (y * y + z * z);
        R[1] = T(2) * (x * y - z * w);
        R[2] = T(2) * (x * z + y * w);
        R[3] = T(2) * (x * y + z * w);
        R[4] = T(1) - T(2) * (x * x + z * z);
        R[5] = T(2) * (y * z - x * w);
        R[6] = T(2) * (x * z - y * w);
        R[7] = T(2) * (y * z + x * w);
        R[8] = T(1) - T(2) * (x * x + y * y);
    }

    // dR_k/d(x,y,z,w) as four 9-entry coefficient rows; each adjoint is the dot product with dL/dR,
    // accumulated k = 0..8 like the reference does.
    template <typename Output, typename Input>
    XYZ_HD void backward(const Output& R, Input& q) const {
        using T = typename Input::value_type;
        const T x = q[0], y = q[1], z = q[2], w = q[3];
        const T z0 = T(0);
        const T jac[4][9] = {
            {z0, T(2) * y, T(2) * z, T(2) * y, T(-4) * x, T(-2) * w, T(2) * z, T(2) * w, T(-4) * x},
            {T(-4) * y, T(2) * x, T(2) * w, T(2) * x, z0, T(2) * z, T(-2) * w, T(2) * z, T(-4) * y},
            {T(-4) * z, T(-2) * w, T(2) * x, T(2) * w, T(-4) * z, T(2) * y, T(2) * x, T(2) * y, z0},
            {z0, T(-2) * z, T(2) * y, T(2) * z, z0, T(-2) * x, T(-2) * y, T(2) * x, z0},
        };
#pragma unroll
        for (std::size_t c = 0; c < 4; ++c) {
            T acc = T(0);
#pragma unroll
            for (std::size_t k = 0; k < 9; ++k) acc += R.grad(k) * jac[c][k];
            q.add_grad(c, acc);
        }
    }
};

template <DifferentiableVariableConcept Input>
    requires(Input::size == 4)
XYZ_HD auto quaternion_to_rotation_matrix(Input& quaternion) {
    using Logic = QuaternionToRotationMatrixLogic<4>;
    return UnaryOperation<9, Logic, Input>(Logic{}, quaternion);
}

}  // namespace op
}  // namespace xyz_autodiff
