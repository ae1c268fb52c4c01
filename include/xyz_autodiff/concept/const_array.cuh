// xyz_autodiff/concept/const_array.cuh -- anything indexable with a static size and a value_type.
// Contract of reference include/xyz_autodiff/concept/const_array.cuh:9-27.
#pragma once

#include <concepts>
#include <cstddef>
#include <type_traits>

namespace xyz_autodiff {

template <typename A>
concept ConstArrayLike = requires(A a, std::size_t i) {
    typename A::value_type;
    { A::size } -> std::convertible_to<std::size_t>;
    { a[i] } -> std::convertible_to<typename A::value_type>;
};

template <typename A, typename B>
concept ConstArrayCompatible = ConstArrayLike<A> && ConstArrayLike<B> &&
    std::same_as<typename A::value_type, typename B::value_type>;

template <typename A, typename B>
concept ConstArraySameSize = ConstArrayLike<A> && ConstArrayLike<B> && (A::size == B::size);

}  // namespace xyz_autodiff
