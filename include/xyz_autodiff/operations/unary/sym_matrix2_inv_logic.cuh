// xyz_autodiff/operations/unary/sym_matrix2_inv_logic.cuh -- inverse of a symmetric 2x2 matrix stored as
// (a, b, c) = [[a, b], [b, c]]:  (c, -b, a) / det,  det = a c - b^2.
// Contract of reference include/xyz_autodiff/operations/unary/sym_matrix2_inv_logic.cuh:15-89, including the
// regularisation |det| < 1e-8 -> det = +1e-8 in both passes.
#pragma once

#include "../math.cuh"
#include "../operation.cuh"

namespace xyz_autodiff {
namespace op {

template <std::size_t InputDim>
    requires(InputDim == 3)
struct SymMatrix2InvLogic {
    static constexpr std::size_t outputDim = 3;

    template <typename T>
    XYZ_HD static T guarded_det(T a, T b, T c) {
        T det = a * c - b * b;
        if (math::abs(det) < T(1e-8)) det = T(1e-8);
        return det;
    }

    template <typename Output, typename Input>
    XYZ_HD void forward(Output& y, const Input& x) const {
        using T = typename Input::value_type;
        const T a = x[0], b = x[1], c = x[2];
        const T r = T(1) / guarded_det(a, b, c);
        y[0] = c * r;
        y[1] = -b * r;
        y[2] = a * r;
    }

    // full 3x3 Jacobian of (c, -b, a)/det w.r.t. (a, b, c); rows = outputs
    template <typename Output, typename Input>
    XYZ_HD void backward(const Output& y, Input& x) const {
        using T = typename Input::value_type;
        const T a = x[0], b = x[1], c = x[2];
        const T r = T(1) / guarded_det(a, b, c);
        const T r2 = r * r;
        const T g0 = y.grad(0), g1 = y.grad(1), g2 = y.grad(2);
        const T d0_da = -c * c * r2, d0_db = T(2) * c * b * r2, d0_dc = r - a * c * r2;
        const T d1_da = b * c * r2, d1_db = -r - T(2) * b * b * r2, d1_dc = a * b * r2;
        const T d2_da = r - a * c * r2, d2_db = T(2) * a * b * r2, d2_dc = -a * a * r2;
        x.add_grad(0, g0 * d0_da + g1 * d1_da + g2 * d2_da);
        x.add_grad(1, g0 * d0_db + g1 * d1_db + g2 * d2_db);
        x.add_grad(2, g0 * d0_dc + g1 * d1_dc + g2 * d2_dc);
    }
};

template <DifferentiableVariableConcept Input>
    requires(Input::size == 3)
XYZ_HD auto sym_matrix2_inv(Input& x) {
    return UnaryOperation<3, SymMatrix2InvLogic<3>, Input>(SymMatrix2InvLogic<3>{}, x);
}

}  // namespace op
}  // namespace xyz_autodiff
