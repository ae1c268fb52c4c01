// xyz_autodiff/concept/const_array.cuh -- anything indexable with a static size and a value_type.
// Contract of reference include/xyz_autodiff/concept/const_array.cuh:9-27.
#pragma once

#include <concepts>
#include <cstddef>
#include <type_traits>
#include <utility>

namespace xyz_autodiff {

namespace detail {

template <typename A>
concept NamesElementAndLength = requires {
    typename A::value_type;
    { A::size } -> std::convertible_to<std::size_t>;
};

template <typename A>
concept ReadableByIndex =
    std::convertible_to<decltype(std::declval<A&>()[std::declval<std::size_t>()]), typename A::value_type>;

}  // namespace detail

template <typename A>
concept ConstArrayLike = detail::NamesElementAndLength<A> && detail::ReadableByIndex<A>;

// two arrays of the same element type / of the same length
template <typename A, typename B>
concept ConstArrayCompatible =
    ConstArrayLike<A> && ConstArrayLike<B> && std::is_same_v<typename A::value_type, typename B::value_type>;

template <typename A, typename B>
concept ConstArraySameSize = ConstArrayLike<A> && ConstArrayLike<B> && (static_cast<std::size_t>(A::size) == static_cast<std::size_t>(B::size));

}  // namespace xyz_autodiff
