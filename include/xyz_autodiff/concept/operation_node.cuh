// xyz_autodiff/concept/operation_node.cuh -- interior nodes of the expression DAG.
// Contract of reference include/xyz_autodiff/concept/operation_node.cuh:10-26.
#pragma once

#include <concepts>
#include <cstddef>

namespace xyz_autodiff {

namespace detail {

// the two sweeps every node takes part in
template <typename N>
concept SweepsBothWays = requires(N node) {
    { node.forward() } -> std::same_as<void>;
    { node.backward() } -> std::same_as<void>;
};

// ... and the finite-difference variant of the reverse sweep, with a step of the node's scalar type
template <typename N>
concept SweepsNumerically = requires(N node, typename N::value_type step) {
    { node.backward_numerical(step) } -> std::same_as<void>;
};

}  // namespace detail

template <typename N>
concept OperationNode = requires(N node) {
    typename N::value_type;
    { N::size } -> std::convertible_to<std::size_t>;
    { node.zero_grad() } -> std::same_as<void>;
} && detail::SweepsBothWays<N> && detail::SweepsNumerically<N>;

}  // namespace xyz_autodiff
