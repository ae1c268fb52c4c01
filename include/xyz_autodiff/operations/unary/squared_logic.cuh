// xyz_autodiff/operations/unary/squared_logic.cuh -- element-wise square.
// Contract of reference include/xyz_autodiff/operations/unary/squared_logic.cuh:11-45.
// The pullback keeps the reference's `g * 2.0 * x`: the double literal promotes the product, so an fp32 graph rounds
// once from the exact double product.
#pragma once

#include "../../detail/pointwise.cuh"
#include "../math.cuh"
#include "../operation.cuh"

namespace xyz_autodiff {
namespace detail::rule {
struct Square {
    template <typename S>
    XYZ_HD static S value(S x) {
        return x * x;
    }
    template <typename S>
    XYZ_HD static S pullback(S x, S g) {
        return static_cast<S>(g * 2.0 * x);
    }
};
}  // namespace detail::rule

namespace op {

template <std::size_t Dim>
struct SquaredLogic : detail::PointwiseMap<Dim, detail::rule::Square> {};

template <typename Input>
    requires UnaryLogicParameterConcept<Input>
XYZ_HD auto squared(Input& x) {
    return detail::make_unary_node<SquaredLogic<Input::size>>(x);
}

}  // namespace op
}  // namespace xyz_autodiff
