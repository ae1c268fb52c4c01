// xyz_autodiff/concept/core_logic.cuh -- the Logic contract: a stateless-or-small functor with
// `outputDim`, `forward(out, in...)` and `backward(out, in...)`.
// Contract of reference include/xyz_autodiff/concept/core_logic.cuh:10-75.
#pragma once

#include <concepts>
#include <type_traits>

#include "variable.cuh"

namespace xyz_autodiff {

namespace detail {
template <typename L>
concept HasOutputDim = requires {
    { L::outputDim } -> std::convertible_to<std::size_t>;
};
}  // namespace detail

template <typename L, typename Input, typename Output>
concept UnaryLogicConcept = VariableConcept<Input> && VariableConcept<Output> && detail::HasOutputDim<L> &&
    requires(L logic, Output& out, const Output& cout, const Input& cin, Input& in) {
        { logic.forward(out, cin) } -> std::same_as<void>;
        { logic.backward(cout, in) } -> std::same_as<void>;
    };

template <typename L, typename Input1, typename Input2, typename Output>
concept BinaryLogicConcept = VariableConcept<Input1> && VariableConcept<Input2> && VariableConcept<Output> &&
    detail::HasOutputDim<L> &&
    requires(L logic, Output& out, const Output& cout, const Input1& c1, const Input2& c2, Input1& i1, Input2& i2) {
        { logic.forward(out, c1, c2) } -> std::same_as<void>;
        { logic.backward(cout, i1, i2) } -> std::same_as<void>;
    };

template <typename L, typename Input1, typename Input2, typename Input3, typename Output>
concept TernaryLogicConcept = VariableConcept<Input1> && VariableConcept<Input2> && VariableConcept<Input3> &&
    VariableConcept<Output> && detail::HasOutputDim<L> &&
    requires(L logic, Output& out, const Output& cout, const Input1& c1, const Input2& c2, const Input3& c3, Input1& i1,
             Input2& i2, Input3& i3) {
        { logic.forward(out, c1, c2, c3) } -> std::same_as<void>;
        { logic.backward(cout, i1, i2, i3) } -> std::same_as<void>;
    };

// Constraints on the operands handed to factories.
template <typename Input>
concept UnaryLogicParameterConcept = DifferentiableVariableConcept<Input>;

template <typename Input1, typename Input2>
concept BinaryLogicParameterConcept = DifferentiableVariableConcept<Input1> && DifferentiableVariableConcept<Input2> &&
    std::is_same_v<typename Input1::value_type, typename Input2::value_type>;

template <typename Input1, typename Input2, typename Input3>
concept TernaryLogicParameterConcept = BinaryLogicParameterConcept<Input1, Input2> &&
    DifferentiableVariableConcept<Input3> && std::is_same_v<typename Input1::value_type, typename Input3::value_type>;

}  // namespace xyz_autodiff
