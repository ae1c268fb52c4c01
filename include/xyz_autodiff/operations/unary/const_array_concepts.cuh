// xyz_autodiff/operations/unary/const_array_concepts.cuh
// Contract of reference include/xyz_autodiff/operations/unary/const_array_concepts.cuh:9-12.
#pragma once

#include <concepts>
#include <cstddef>
#include <type_traits>

namespace xyz_autodiff {
namespace op {

// indexable with a value_type: ConstArray, Variable, VariableRef, operation nodes, ...
template <typename A>
concept ArrayLikeConcept = requires(A a, std::size_t i) {
    { a[i] } -> std::convertible_to<typename std::remove_reference_t<A>::value_type>;
};

}  // namespace op
}  // namespace xyz_autodiff
