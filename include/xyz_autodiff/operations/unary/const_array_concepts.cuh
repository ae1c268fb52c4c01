// xyz_autodiff/operations/unary/const_array_concepts.cuh
// Contract of reference include/xyz_autodiff/operations/unary/const_array_concepts.cuh:9-12.
#pragma once

#include <concepts>
#include <cstddef>
#include <type_traits>
#include <utility>

namespace xyz_autodiff {
namespace op {

// Anything that names its element type and can be subscripted: ConstArray, Variable, VariableRef, operation nodes.
template <typename A>
concept ArrayLikeConcept =
    std::convertible_to<decltype(std::declval<A&>()[std::size_t{}]), typename std::remove_reference_t<A>::value_type>;

}  // namespace op
}  // namespace xyz_autodiff
