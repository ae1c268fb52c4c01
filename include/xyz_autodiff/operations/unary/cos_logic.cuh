// xyz_autodiff/operations/unary/cos_logic.cuh -- element-wise cosine.
// Contract of reference include/xyz_autodiff/operations/unary/cos_logic.cuh:13-51.
#pragma once

#include "../../detail/pointwise.cuh"
#include "../math.cuh"
#include "../operation.cuh"

namespace xyz_autodiff {
namespace detail::rule {
struct Cosine {
    template <typename S>
    XYZ_HD static S value(S x) {
        return math::cos(x);
    }
    template <typename S>
    XYZ_HD static S pullback(S x, S g) {
        return g * (-math::sin(x));
    }
};
}  // namespace detail::rule

namespace op {

template <std::size_t Dim>
struct CosLogic : detail::PointwiseMap<Dim, detail::rule::Cosine> {};

template <std::size_t Dim, DifferentiableVariableConcept Input>
    requires(Input::size == Dim)
XYZ_HD auto cos(Input& x) {
    return detail::make_unary_node<CosLogic<Dim>>(x);
}

template <DifferentiableVariableConcept Input>
XYZ_HD auto cos(Input& x) {
    return cos<Input::size>(x);
}

}  // namespace op
}  // namespace xyz_autodiff
