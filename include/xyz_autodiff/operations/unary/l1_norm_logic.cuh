// xyz_autodiff/operations/unary/l1_norm_logic.cuh -- sum of absolute values of an InputDim-vector (scalar output).
// Contract of reference include/xyz_autodiff/operations/unary/l1_norm_logic.cuh:13-54.
#pragma once

#include "../../detail/pointwise.cuh"
#include "../math.cuh"
#include "../operation.cuh"

namespace xyz_autodiff {
namespace detail::rule {
struct AbsoluteSum {
    template <typename S>
    XYZ_HD static S term(S x) {
        return math::abs(x);
    }
    template <typename S>
    XYZ_HD static S finish(S total) {
        return total;
    }
    template <typename S>
    XYZ_HD static bool has_adjoint(S result) {
        (void)result;
        return true;
    }
    template <typename S>
    XYZ_HD static S pullback(S x, S result, S g) {
        (void)result;
        const S sign = x > S(0) ? S(1) : (x < S(0) ? S(-1) : S(0));  // subgradient 0 at the kink
        return g * sign;
    }
};
}  // namespace detail::rule

namespace op {

template <std::size_t InputDim>
struct L1NormLogic : detail::FoldToScalar<InputDim, detail::rule::AbsoluteSum> {};

template <std::size_t Dim, DifferentiableVariableConcept Input>
    requires(Input::size == Dim)
XYZ_HD auto l1_norm(Input& x) {
    return detail::make_unary_node<L1NormLogic<Dim>>(x);
}

template <DifferentiableVariableConcept Input>
XYZ_HD auto l1_norm(Input& x) {
    return l1_norm<Input::size>(x);
}

}  // namespace op
}  // namespace xyz_autodiff
