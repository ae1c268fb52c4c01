// examples/mini-gaussian-splatting/operations/covariance_generation.cuh -- user-extension-style custom
// Logics of the splat example: 2-D covariance from (scale, rotation).
//   M = R(theta) diag(sx, sy)      CovarianceMatrixGenerationLogic   (2 + 1 -> 4, row-major M)
//   Sigma = M M^T, packed          MatrixToCovariance3ParamLogic     (4 -> 3: S00, S01, S11)
//   both fused                     ScaleRotationToCovariance3ParamLogic (2 + 1 -> 3; the one the kernel uses)
// Contract of reference examples/mini-gaussian-splatting/operations/covariance_generation.cuh:15-239.
#pragma once

#include <xyz_autodiff/operations/math.cuh>
#include <xyz_autodiff/operations/operation.cuh>

namespace xyz_autodiff {
namespace op {

namespace detail_cov {
// M = R(theta) diag(sx, sy) and its reverse rule, shared by the fused and unfused Logics
template <typename T>
struct RotScale {
    T c, s, m00, m01, m10, m11;
    XYZ_HD RotScale(T sx, T sy, T theta) : c(math::cos(theta)), s(math::sin(theta)) {
        m00 = sx * c;
        m01 = -sy * s;
        m10 = sx * s;
        m11 = sy * c;
    }
    // adjoints of (sx, sy, theta) given the adjoint of M
    XYZ_HD void pull(T sx, T sy, T g00, T g01, T g10, T g11, T& gsx, T& gsy, T& gtheta) const {
        gsx = g00 * c + g10 * s;
        gsy = g01 * (-s) + g11 * c;
        gtheta = g00 * (-sx * s) + g01 * (-sy * c) + g10 * (sx * c) + g11 * (-sy * s);
    }
};
}  // namespace detail_cov

template <typename Input1, typename Input2>
    requires BinaryLogicParameterConcept<Input1, Input2> && (Input1::size == 2) && (Input2::size == 1)
struct CovarianceMatrixGenerationLogic {
    using T = typename Input1::value_type;
    static constexpr std::size_t Dim = 4;
    static constexpr std::size_t outputDim = Dim;
    using Output = Variable<Dim, T>;

    XYZ_HD void forward(Output& M, const Input1& scale, const Input2& rotation) const {
        const detail_cov::RotScale<T> rs(scale[0], scale[1], rotation[0]);
        M[0] = rs.m00;
        M[1] = rs.m01;
        M[2] = rs.m10;
        M[3] = rs.m11;
    }

    XYZ_HD void backward(const Output& M, Input1& scale, Input2& rotation) const {
        const detail_cov::RotScale<T> rs(scale[0], scale[1], rotation[0]);
        T gsx, gsy, gtheta;
        rs.pull(scale[0], scale[1], M.grad(0), M.grad(1), M.grad(2), M.grad(3), gsx, gsy, gtheta);
        scale.add_grad(0, gsx);
        scale.add_grad(1, gsy);
        rotation.add_grad(0, gtheta);
    }
};

template <typename Input>
    requires UnaryLogicParameterConcept<Input> && (Input::size == 4)
struct MatrixToCovariance3ParamLogic {
    using T = typename Input::value_type;
    static constexpr std::size_t Dim = 3;
    static constexpr std::size_t outputDim = Dim;
    using Output = Variable<Dim, T>;

    XYZ_HD void forward(Output& S, const Input& M) const {
        S[0] = M[0] * M[0] + M[1] * M[1];
        S[1] = M[0] * M[2] + M[1] * M[3];
        S[2] = M[2] * M[2] + M[3] * M[3];
    }

    XYZ_HD void backward(const Output& S, Input& M) const {
        const T g00 = S.grad(0), g01 = S.grad(1), g11 = S.grad(2);
        M.add_grad(0, g00 * T(2) * M[0] + g01 * M[2]);
        M.add_grad(1, g00 * T(2) * M[1] + g01 * M[3]);
        M.add_grad(2, g01 * M[0] + g11 * T(2) * M[2]);
        M.add_grad(3, g01 * M[1] + g11 * T(2) * M[3]);
    }
};

template <typename Input1, typename Input2>
    requires BinaryLogicParameterConcept<Input1, Input2> && (Input1::size == 2) && (Input2::size == 1)
struct ScaleRotationToCovariance3ParamLogic {
    using T = typename Input1::value_type;
    static constexpr std::size_t Dim = 3;
    static constexpr std::size_t outputDim = Dim;
    using Output = Variable<Dim, T>;

    XYZ_HD void forward(Output& S, const Input1& scale, const Input2& rotation) const {
        const detail_cov::RotScale<T> rs(scale[0], scale[1], rotation[0]);
        S[0] = rs.m00 * rs.m00 + rs.m01 * rs.m01;
        S[1] = rs.m00 * rs.m10 + rs.m01 * rs.m11;
        S[2] = rs.m10 * rs.m10 + rs.m11 * rs.m11;
    }

    XYZ_HD void backward(const Output& S, Input1& scale, Input2& rotation) const {
        const detail_cov::RotScale<T> rs(scale[0], scale[1], rotation[0]);
        const T g00 = S.grad(0), g01 = S.grad(1), g11 = S.grad(2);
        const T gm00 = g00 * T(2) * rs.m00 + g01 * rs.m10;
        const T gm01 = g00 * T(2) * rs.m01 + g01 * rs.m11;
        const T gm10 = g01 * rs.m00 + g11 * T(2) * rs.m10;
        const T gm11 = g01 * rs.m01 + g11 * T(2) * rs.m11;
        T gsx, gsy, gtheta;
        rs.pull(scale[0], scale[1], gm00, gm01, gm10, gm11, gsx, gsy, gtheta);
        scale.add_grad(0, gsx);
        scale.add_grad(1, gsy);
        rotation.add_grad(0, gtheta);
    }
};

template <typename Input1, typename Input2>
    requires BinaryLogicParameterConcept<Input1, Input2> && (Input1::size == 2) && (Input2::size == 1)
XYZ_HD auto generate_covariance_matrix(Input1& scale, Input2& rotation) {
    using Logic = CovarianceMatrixGenerationLogic<Input1, Input2>;
    return BinaryOperation<Logic::outputDim, Logic, Input1, Input2>(Logic{}, scale, rotation);
}

template <typename Input>
    requires UnaryLogicParameterConcept<Input> && (Input::size == 4)
XYZ_HD auto matrix_to_covariance_3param(Input& M) {
    using Logic = MatrixToCovariance3ParamLogic<Input>;
    return UnaryOperation<Logic::outputDim, Logic, Input>(Logic{}, M);
}

template <typename Input1, typename Input2>
    requires BinaryLogicParameterConcept<Input1, Input2> && (Input1::size == 2) && (Input2::size == 1)
XYZ_HD auto scale_rotation_to_covariance_3param(Input1& scale, Input2& rotation) {
    using Logic = ScaleRotationToCovariance3ParamLogic<Input1, Input2>;
    return BinaryOperation<Logic::outputDim, Logic, Input1, Input2>(Logic{}, scale, rotation);
}

}  // namespace op
}  // namespace xyz_autodiff
