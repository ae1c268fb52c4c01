// xyz_autodiff/concept/variable.cuh -- what a graph value must offer.
// Contract of reference include/xyz_autodiff/concept/variable.cuh:9-36.
#pragma once

#include <concepts>
#include <cstddef>
#include <type_traits>

namespace xyz_autodiff {

template <typename T>
concept FloatingPointConcept = std::same_as<T, float> || std::same_as<T, double>;

// Forward side: typed, fixed-size, indexable, resettable.
template <typename V>
concept VariableConcept = requires(V v) {
    typename V::value_type;
    requires FloatingPointConcept<typename V::value_type>;
    { V::size } -> std::convertible_to<std::size_t>;
    { v[std::size_t{}] } -> std::convertible_to<typename V::value_type&>;
    { v[std::size_t{}] } -> std::convertible_to<const typename V::value_type&>;
    { v.zero_grad() } -> std::same_as<void>;
};

// Reverse side: readable adjoint, thread-safe accumulation.
template <typename V>
concept DifferentiableVariableConcept = VariableConcept<V> && requires(V v, typename V::value_type x) {
    { v.grad(std::size_t{}) } -> std::convertible_to<const typename V::value_type&>;
    { v.add_grad(std::size_t{}, x) } -> std::same_as<void>;
};

}  // namespace xyz_autodiff
