// xyz_autodiff/concept/operation_node.cuh -- interior nodes of the expression DAG.
// Contract of reference include/xyz_autodiff/concept/operation_node.cuh:10-26.
#pragma once

#include <concepts>
#include <cstddef>

namespace xyz_autodiff {

template <typename N>
concept OperationNode = requires(N node) {
    typename N::value_type;
    { N::size } -> std::convertible_to<std::size_t>;
    { node.forward() } -> std::same_as<void>;
    { node.zero_grad() } -> std::same_as<void>;
    { node.backward() } -> std::same_as<void>;
    { node.backward_numerical(typename N::value_type{}) } -> std::same_as<void>;
};

}  // namespace xyz_autodiff
