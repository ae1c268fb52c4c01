// xyz_autodiff/operations/unary/const_array_sub_logic.cuh -- y[i] = x[i] - c[i] with c a constant array held BY REFERENCE
// (the array must outlive the node).
// Contract of reference include/xyz_autodiff/operations/unary/const_array_sub_logic.cuh:14-48.
#pragma once

#include "../../detail/pointwise.cuh"
#include "../operation.cuh"
#include "const_array_concepts.cuh"

namespace xyz_autodiff {
namespace detail::rule {
struct MinusEntry {
    template <typename S, typename C>
    XYZ_HD static auto value(const S& x, const C& c) {
        return x - c;
    }
};
}  // namespace detail::rule

namespace op {

template <std::size_t Dim, typename ConstantArray>
    requires ArrayLikeConcept<ConstantArray>
struct ConstArraySubLogic : detail::PointwiseWithArray<Dim, ConstantArray, detail::rule::MinusEntry> {
    using detail::PointwiseWithArray<Dim, ConstantArray, detail::rule::MinusEntry>::PointwiseWithArray;
};

template <DifferentiableVariableConcept Input, typename ConstantArray>
    requires ArrayLikeConcept<ConstantArray>
XYZ_HD auto const_sub(Input& x, const ConstantArray& c) {
    return detail::make_unary_node<ConstArraySubLogic<Input::size, ConstantArray>>(x, c);
}

}  // namespace op
}  // namespace xyz_autodiff
