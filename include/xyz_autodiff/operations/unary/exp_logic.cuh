// xyz_autodiff/operations/unary/exp_logic.cuh -- element-wise exponential.
// Contract of reference include/xyz_autodiff/operations/unary/exp_logic.cuh:13-51.
#pragma once

#include "../../detail/pointwise.cuh"
#include "../math.cuh"
#include "../operation.cuh"

namespace xyz_autodiff {
namespace detail::rule {
struct Exponential {
    template <typename S>
    XYZ_HD static S value(S x) {
        return math::exp(x);
    }
    template <typename S>
    XYZ_HD static S pullback(S x, S g) {
        return g * math::exp(x);
    }
};
}  // namespace detail::rule

namespace op {

template <std::size_t Dim>
struct ExpLogic : detail::PointwiseMap<Dim, detail::rule::Exponential> {};

template <std::size_t Dim, DifferentiableVariableConcept Input>
    requires(Input::size == Dim)
XYZ_HD auto exp(Input& x) {
    return detail::make_unary_node<ExpLogic<Dim>>(x);
}

template <DifferentiableVariableConcept Input>
XYZ_HD auto exp(Input& x) {
    return exp<Input::size>(x);
}

}  // namespace op
}  // namespace xyz_autodiff
