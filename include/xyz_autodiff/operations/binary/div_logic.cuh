// xyz_autodiff/operations/binary/div_logic.cuh -- element-wise quotient of two equally sized operands.
// Contract of reference include/xyz_autodiff/operations/binary/div_logic.cuh:11-50.
#pragma once

#include "../../detail/pointwise.cuh"
#include "../operation.cuh"

namespace xyz_autodiff {
namespace detail::rule {
struct Quotient {
    template <typename S>
    XYZ_HD static S value(S a, S b) {
        return a / b;
    }
    template <typename S>
    XYZ_HD static void pullback(S a, S b, S g, S& to_a, S& to_b) {
        to_a = g / b;
        to_b = -g * a / (b * b);
    }
};
}  // namespace detail::rule

namespace op {

template <typename Input1, typename Input2>
    requires BinaryLogicParameterConcept<Input1, Input2> && (Input1::size == Input2::size)
struct DivLogic : detail::PointwisePair<Input1, Input2, detail::rule::Quotient> {};

template <typename Input1, typename Input2>
    requires BinaryLogicParameterConcept<Input1, Input2>
XYZ_HD auto div(Input1& a, Input2& b) {
    return detail::make_binary_node<DivLogic<Input1, Input2>>(a, b);
}

}  // namespace op
}  // namespace xyz_autodiff
