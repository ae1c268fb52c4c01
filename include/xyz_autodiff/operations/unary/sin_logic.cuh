// xyz_autodiff/operations/unary/sin_logic.cuh -- element-wise sine.
// Contract of reference include/xyz_autodiff/operations/unary/sin_logic.cuh:13-51.
#pragma once

#include "../../detail/pointwise.cuh"
#include "../math.cuh"
#include "../operation.cuh"

namespace xyz_autodiff {
namespace detail::rule {
struct Sine {
    template <typename S>
    XYZ_HD static S value(S x) {
        return math::sin(x);
    }
    template <typename S>
    XYZ_HD static S pullback(S x, S g) {
        return g * math::cos(x);
    }
};
}  // namespace detail::rule

namespace op {

template <std::size_t Dim>
struct SinLogic : detail::PointwiseMap<Dim, detail::rule::Sine> {};

template <std::size_t Dim, DifferentiableVariableConcept Input>
    requires(Input::size == Dim)
XYZ_HD auto sin(Input& x) {
    return detail::make_unary_node<SinLogic<Dim>>(x);
}

template <DifferentiableVariableConcept Input>
XYZ_HD auto sin(Input& x) {
    return sin<Input::size>(x);
}

}  // namespace op
}  // namespace xyz_autodiff
