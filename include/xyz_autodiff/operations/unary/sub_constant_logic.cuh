// xyz_autodiff/operations/unary/sub_constant_logic.cuh -- shift down by a scalar constant held in the Logic.
// Contract of reference include/xyz_autodiff/operations/unary/sub_constant_logic.cuh:11-48.
#pragma once

#include "../../detail/pointwise.cuh"
#include "../operation.cuh"

namespace xyz_autodiff {
namespace detail::rule {
struct ShiftDown {
    template <typename S>
    XYZ_HD static S value(S x, S c) {
        return x - c;
    }
    template <typename S>
    XYZ_HD static S pullback(S g, S c) {
        (void)c;
        return g;
    }
};
}  // namespace detail::rule

namespace op {

template <typename Input>
    requires UnaryLogicParameterConcept<Input>
struct SubConstantLogic : detail::PointwiseWithScalar<Input, detail::rule::ShiftDown> {
    using detail::PointwiseWithScalar<Input, detail::rule::ShiftDown>::PointwiseWithScalar;
};

template <typename Input>
    requires UnaryLogicParameterConcept<Input>
XYZ_HD auto sub_constant(Input& x, typename Input::value_type constant) {
    return detail::make_unary_node<SubConstantLogic<Input>>(x, constant);
}

}  // namespace op
}  // namespace xyz_autodiff
