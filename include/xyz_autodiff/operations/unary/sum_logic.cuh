// xyz_autodiff/operations/unary/sum_logic.cuh -- sum of an InputDim-vector (scalar output).
// Contract of reference include/xyz_autodiff/operations/unary/sum_logic.cuh:13-53.
#pragma once

#include "../../detail/pointwise.cuh"
#include "../math.cuh"
#include "../operation.cuh"

namespace xyz_autodiff {
namespace detail::rule {
struct PlainSum {
    template <typename S>
    XYZ_HD static S term(S x) {
        return x;
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
        (void)x;
        (void)result;
        return g;
    }
};
}  // namespace detail::rule

namespace op {

template <std::size_t InputDim>
struct SumLogic : detail::FoldToScalar<InputDim, detail::rule::PlainSum> {};

template <std::size_t Dim, DifferentiableVariableConcept Input>
    requires(Input::size == Dim)
XYZ_HD auto sum(Input& x) {
    return detail::make_unary_node<SumLogic<Dim>>(x);
}

template <DifferentiableVariableConcept Input>
XYZ_HD auto sum(Input& x) {
    return sum<Input::size>(x);
}

}  // namespace op
}  // namespace xyz_autodiff
