// xyz_autodiff/operations/unary/sigmoid_logic.cuh -- element-wise logistic sigmoid 1 / (1 + exp(-x)).
// Contract of reference include/xyz_autodiff/operations/unary/sigmoid_logic.cuh:13-52.
#pragma once

#include "../../detail/pointwise.cuh"
#include "../math.cuh"
#include "../operation.cuh"

namespace xyz_autodiff {
namespace detail::rule {
struct Logistic {
    template <typename S>
    XYZ_HD static S value(S x) {
        return math::sigmoid(x);
    }
    template <typename S>
    XYZ_HD static S pullback(S x, S g) {
        const S s = math::sigmoid(x);
        return g * (s * (S(1) - s));
    }
};
}  // namespace detail::rule

namespace op {

template <std::size_t Dim>
struct SigmoidLogic : detail::PointwiseMap<Dim, detail::rule::Logistic> {};

template <std::size_t Dim, DifferentiableVariableConcept Input>
    requires(Input::size == Dim)
XYZ_HD auto sigmoid(Input& x) {
    return detail::make_unary_node<SigmoidLogic<Dim>>(x);
}

template <DifferentiableVariableConcept Input>
XYZ_HD auto sigmoid(Input& x) {
    return sigmoid<Input::size>(x);
}

}  // namespace op
}  // namespace xyz_autodiff
