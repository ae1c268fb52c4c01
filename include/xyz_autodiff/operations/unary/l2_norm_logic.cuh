// xyz_autodiff/operations/unary/l2_norm_logic.cuh -- Euclidean norm of an InputDim-vector (scalar output).
// Contract of reference include/xyz_autodiff/operations/unary/l2_norm_logic.cuh:13-57.
#pragma once

#include "../../detail/pointwise.cuh"
#include "../math.cuh"
#include "../operation.cuh"

namespace xyz_autodiff {
namespace detail::rule {
struct EuclideanLength {
    template <typename S>
    XYZ_HD static S term(S x) {
        return x * x;
    }
    template <typename S>
    XYZ_HD static S finish(S total) {
        return math::sqrt(total);
    }
    template <typename S>
    XYZ_HD static bool has_adjoint(S result) {
        return result > S(1e-8);  // no adjoint at (numerically) zero vectors, like the reference
    }
    template <typename S>
    XYZ_HD static S pullback(S x, S result, S g) {
        return g * x / result;
    }
};
}  // namespace detail::rule

namespace op {

template <std::size_t InputDim>
struct L2NormLogic : detail::FoldToScalar<InputDim, detail::rule::EuclideanLength> {};

template <std::size_t Dim, DifferentiableVariableConcept Input>
    requires(Input::size == Dim)
XYZ_HD auto l2_norm(Input& x) {
    return detail::make_unary_node<L2NormLogic<Dim>>(x);
}

template <DifferentiableVariableConcept Input>
XYZ_HD auto l2_norm(Input& x) {
    return l2_norm<Input::size>(x);
}

}  // namespace op
}  // namespace xyz_autodiff
