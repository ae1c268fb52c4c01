// xyz_autodiff/operations/unary/add_constant_logic.cuh -- shift by a scalar constant held in the Logic.
// Contract of reference include/xyz_autodiff/operations/unary/add_constant_logic.cuh:11-48.
#pragma once

#include "../../detail/pointwise.cuh"
#include "../operation.cuh"

namespace xyz_autodiff {
namespace detail::rule {
struct ShiftUp {
    template <typename S>
    XYZ_HD static S value(S x, S c) {
        return x + c;
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
struct AddConstantLogic : detail::PointwiseWithScalar<Input, detail::rule::ShiftUp> {
    using detail::PointwiseWithScalar<Input, detail::rule::ShiftUp>::PointwiseWithScalar;
};

template <typename Input>
    requires UnaryLogicParameterConcept<Input>
XYZ_HD auto add_constant(Input& x, typename Input::value_type constant) {
    return detail::make_unary_node<AddConstantLogic<Input>>(x, constant);
}

}  // namespace op
}  // namespace xyz_autodiff
