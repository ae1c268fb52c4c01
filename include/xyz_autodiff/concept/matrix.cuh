// xyz_autodiff/concept/matrix.cuh -- 2-D views over Variables.
// Contract of reference include/xyz_autodiff/concept/matrix.cuh:10-28.
#pragma once

#include <concepts>
#include <cstddef>
#include <type_traits>
#include <utility>

namespace xyz_autodiff {

namespace detail {

template <typename M>
concept HasStaticExtents = requires {
    typename M::value_type;
    { M::rows } -> std::convertible_to<std::size_t>;
    { M::cols } -> std::convertible_to<std::size_t>;
};

// (row, column) access on a mutable and on a const view
template <typename M>
concept ElementAt = requires(M view, const M const_view, std::size_t r, std::size_t c) {
    { view(r, c) } -> std::convertible_to<typename M::value_type>;
    { const_view(r, c) } -> std::convertible_to<typename M::value_type>;
};

}  // namespace detail

template <typename M>
concept MatrixViewConcept = detail::HasStaticExtents<M> && detail::ElementAt<M> && requires(M view) {
    { view.data() } -> std::convertible_to<const typename M::value_type*>;
    { view.transpose() };
};

}  // namespace xyz_autodiff
