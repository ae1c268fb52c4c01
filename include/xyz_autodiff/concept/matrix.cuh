// xyz_autodiff/concept/matrix.cuh -- 2-D views over Variables.
// Contract of reference include/xyz_autodiff/concept/matrix.cuh:10-28.
#pragma once

#include <concepts>
#include <cstddef>
#include <type_traits>
#include <utility>

namespace xyz_autodiff {

template <typename M>
concept MatrixViewConcept = requires(M m) {
    typename M::value_type;
    { M::rows } -> std::convertible_to<std::size_t>;
    { M::cols } -> std::convertible_to<std::size_t>;
    { m(std::size_t{}, std::size_t{}) } -> std::convertible_to<typename M::value_type>;
    { std::as_const(m)(std::size_t{}, std::size_t{}) } -> std::convertible_to<typename M::value_type>;
    { m.data() } -> std::convertible_to<const typename M::value_type*>;
    { m.transpose() };
};

}  // namespace xyz_autodiff
