// xyz_autodiff/operations/unary/mul_constant_logic.cuh -- scaling by a scalar constant held in the Logic.
// Contract of reference include/xyz_autodiff/operations/unary/mul_constant_logic.cuh:11-48.
#pragma once

#include "../../detail/pointwise.cuh"
#include "../operation.cuh"

namespace xyz_autodiff {
namespace detail::rule {
struct ScaleBy {
    template <typename S>
    XYZ_HD static S value(S x, S c) {
        return x * c;
    }
    template <typename S>
    XYZ_HD static S pullback(S g, S c) {
        return g * c;
    }
};
}  // namespace detail::rule

namespace op {

template <typename Input>
    requires UnaryLogicParameterConcept<Input>
struct MulConstantLogic : detail::PointwiseWithScalar<Input, detail::rule::ScaleBy> {
    using detail::PointwiseWithScalar<Input, detail::rule::ScaleBy>::PointwiseWithScalar;
};

template <typename Input>
    requires UnaryLogicParameterConcept<Input>
XYZ_HD auto mul_constant(Input& x, typename Input::value_type constant) {
    return detail::make_unary_node<MulConstantLogic<Input>>(x, constant);
}

}  // namespace op
}  // namespace xyz_autodiff
