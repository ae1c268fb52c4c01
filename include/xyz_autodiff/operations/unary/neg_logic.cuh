// xyz_autodiff/operations/unary/neg_logic.cuh -- element-wise negation.
// Contract of reference include/xyz_autodiff/operations/unary/neg_logic.cuh:13-49.
#pragma once

#include "../../detail/pointwise.cuh"
#include "../math.cuh"
#include "../operation.cuh"

namespace xyz_autodiff {
namespace detail::rule {
struct Negation {
    template <typename S>
    XYZ_HD static S value(S x) {
        return -x;
    }
    template <typename S>
    XYZ_HD static S pullback(S x, S g) {
        (void)x;
        return -g;
    }
};
}  // namespace detail::rule

namespace op {

template <std::size_t Dim>
struct NegLogic : detail::PointwiseMap<Dim, detail::rule::Negation> {};

template <std::size_t Dim, DifferentiableVariableConcept Input>
    requires(Input::size == Dim)
XYZ_HD auto neg(Input& x) {
    return detail::make_unary_node<NegLogic<Dim>>(x);
}

template <DifferentiableVariableConcept Input>
XYZ_HD auto neg(Input& x) {
    return neg<Input::size>(x);
}

}  // namespace op
}  // namespace xyz_autodiff
