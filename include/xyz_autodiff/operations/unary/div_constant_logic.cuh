// xyz_autodiff/operations/unary/div_constant_logic.cuh -- division by a scalar constant held in the Logic.
// Contract of reference include/xyz_autodiff/operations/unary/div_constant_logic.cuh:11-48.
#pragma once

#include "../../detail/pointwise.cuh"
#include "../operation.cuh"

namespace xyz_autodiff {
namespace detail::rule {
struct DivideBy {
    template <typename S>
    XYZ_HD static S value(S x, S c) {
        return x / c;
    }
    template <typename S>
    XYZ_HD static S pullback(S g, S c) {
        return g / c;
    }
};
}  // namespace detail::rule

namespace op {

template <typename Input>
    requires UnaryLogicParameterConcept<Input>
struct DivConstantLogic : detail::PointwiseWithScalar<Input, detail::rule::DivideBy> {
    using detail::PointwiseWithScalar<Input, detail::rule::DivideBy>::PointwiseWithScalar;
};

template <typename Input>
    requires UnaryLogicParameterConcept<Input>
XYZ_HD auto div_constant(Input& x, typename Input::value_type constant) {
    return detail::make_unary_node<DivConstantLogic<Input>>(x, constant);
}

}  // namespace op
}  // namespace xyz_autodiff
