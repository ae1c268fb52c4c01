// examples/mini-gaussian-splatting/operations/mahalanobis_distance.cuh -- squared Mahalanobis distance
// d^T Q d for a packed symmetric 2x2 Q = (a, b, c):  a dx^2 + 2 b dx dy + c dy^2.
//   MahalanobisDistanceLogic            operands: the difference vector d, and Q
//   MahalanobisDistanceWithCenterLogic  the query point is a constant of the Logic, operands: center, Q
// Contract of reference examples/mini-gaussian-splatting/operations/mahalanobis_distance.cuh:18-146.
#pragma once

#include <xyz_autodiff/operations/operation.cuh>

namespace xyz_autodiff {
namespace op {

namespace detail_maha {
template <typename T>
XYZ_HD T quadratic_form(T a, T b, T c, T dx, T dy) {
    return a * dx * dx + T(2) * b * dx * dy + c * dy * dy;
}
}  // namespace detail_maha

template <typename Input1, typename Input2>
    requires BinaryLogicParameterConcept<Input1, Input2> && (Input1::size == 2) && (Input2::size == 3)
struct MahalanobisDistanceLogic {
    using T = typename Input1::value_type;
    static constexpr std::size_t Dim = 1;
    static constexpr std::size_t outputDim = Dim;
    using Output = Variable<Dim, T>;

    XYZ_HD void forward(Output& d2, const Input1& diff, const Input2& Q) const {
        d2[0] = detail_maha::quadratic_form<T>(Q[0], Q[1], Q[2], diff[0], diff[1]);
    }

    XYZ_HD void backward(const Output& d2, Input1& diff, Input2& Q) const {
        const T g = d2.grad(0), dx = diff[0], dy = diff[1];
        diff.add_grad(0, g * (T(2) * Q[0] * dx + T(2) * Q[1] * dy));
        diff.add_grad(1, g * (T(2) * Q[1] * dx + T(2) * Q[2] * dy));
        Q.add_grad(0, g * dx * dx);
        Q.add_grad(1, g * T(2) * dx * dy);
        Q.add_grad(2, g * dy * dy);
    }
};

template <typename Input1, typename Input2>
    requires BinaryLogicParameterConcept<Input1, Input2> && (Input1::size == 2) && (Input2::size == 3)
struct MahalanobisDistanceWithCenterLogic {
    using T = typename Input1::value_type;
    static constexpr std::size_t Dim = 1;
    static constexpr std::size_t outputDim = Dim;
    using Output = Variable<Dim, T>;

    T point_x, point_y;  // the query point (a pixel), not differentiated

    XYZ_HD MahalanobisDistanceWithCenterLogic(T px, T py) : point_x(px), point_y(py) {}

    XYZ_HD void forward(Output& d2, const Input1& center, const Input2& Q) const {
        d2[0] = detail_maha::quadratic_form<T>(Q[0], Q[1], Q[2], point_x - center[0], point_y - center[1]);
    }

    XYZ_HD void backward(const Output& d2, Input1& center, Input2& Q) const {
        const T g = d2.grad(0), dx = point_x - center[0], dy = point_y - center[1];
        // d(dx)/d(center_x) = -1
        center.add_grad(0, -(g * (T(2) * Q[0] * dx + T(2) * Q[1] * dy)));
        center.add_grad(1, -(g * (T(2) * Q[1] * dx + T(2) * Q[2] * dy)));
        Q.add_grad(0, g * dx * dx);
        Q.add_grad(1, g * T(2) * dx * dy);
        Q.add_grad(2, g * dy * dy);
    }
};

template <typename Input1, typename Input2>
    requires BinaryLogicParameterConcept<Input1, Input2> && (Input1::size == 2) && (Input2::size == 3)
XYZ_HD auto mahalanobis_distance(Input1& diff, Input2& Q) {
    using Logic = MahalanobisDistanceLogic<Input1, Input2>;
    return BinaryOperation<Logic::outputDim, Logic, Input1, Input2>(Logic{}, diff, Q);
}

template <typename Input1, typename Input2>
    requires BinaryLogicParameterConcept<Input1, Input2> && (Input1::size == 2) && (Input2::size == 3)
XYZ_HD auto mahalanobis_distance_with_center(typename Input1::value_type point_x, typename Input1::value_type point_y,
                                             Input1& center, Input2& Q) {
    using Logic = MahalanobisDistanceWithCenterLogic<Input1, Input2>;
    return BinaryOperation<Logic::outputDim, Logic, Input1, Input2>(Logic(point_x, point_y), center, Q);
}

}  // namespace op
}  // namespace xyz_autodiff
