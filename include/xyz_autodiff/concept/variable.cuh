// xyz_autodiff/concept/variable.cuh -- what a graph value must offer.
// Contract of reference include/xyz_autodiff/concept/variable.cuh:9-36, assembled from single-purpose pieces so that
// a failed constraint names the missing capability.
#pragma once

#include <concepts>
#include <cstddef>
#include <type_traits>

namespace xyz_autodiff {

namespace detail {

template <typename S>
inline constexpr bool is_graph_scalar_v = std::is_same_v<S, float> || std::is_same_v<S, double>;

// a compile-time length and a scalar type the graph can differentiate
template <typename V>
concept StaticallyShaped = requires {
    typename V::value_type;
    { V::size } -> std::convertible_to<std::size_t>;
} && is_graph_scalar_v<typename V::value_type>;

// component access, usable as an lvalue and as a read
template <typename V>
concept ComponentAccess = requires(V value, std::size_t i) {
    { value[i] } -> std::convertible_to<typename V::value_type&>;
    { value[i] } -> std::convertible_to<const typename V::value_type&>;
};

template <typename V>
concept AdjointResettable = requires(V value) {
    { value.zero_grad() } -> std::same_as<void>;
};

// the reverse-mode port: read one adjoint component, accumulate into one
template <typename V>
concept AdjointPort = requires(V value, std::size_t i, typename V::value_type increment) {
    { value.grad(i) } -> std::convertible_to<const typename V::value_type&>;
    { value.add_grad(i, increment) } -> std::same_as<void>;
};

}  // namespace detail

template <typename T>
concept FloatingPointConcept = detail::is_graph_scalar_v<T>;

// Forward side: typed, fixed-size, indexable, resettable.
template <typename V>
concept VariableConcept = detail::StaticallyShaped<V> && detail::ComponentAccess<V> && detail::AdjointResettable<V>;

// Reverse side: readable adjoint, thread-safe accumulation.
template <typename V>
concept DifferentiableVariableConcept = VariableConcept<V> && detail::AdjointPort<V>;

}  // namespace xyz_autodiff
