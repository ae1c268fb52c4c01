// xyz_autodiff/concept/core_logic.cuh -- the Logic contract: a stateless-or-small functor with
// `outputDim`, `forward(out, in...)` and `backward(out, in...)`.
// Contract of reference include/xyz_autodiff/concept/core_logic.cuh:10-75.  One variadic definition serves every
// arity; the reference's three named concepts (and the three operand constraints used by the factories) are its
// instances.
#pragma once

#include <concepts>
#include <cstddef>
#include <type_traits>
#include <utility>

#include "variable.cuh"

namespace xyz_autodiff {

namespace detail {

template <typename L>
concept HasOutputDim = requires {
    { L::outputDim } -> std::convertible_to<std::size_t>;
};

// logic.forward(result, const operands...) and logic.backward(const result, operands...) both return void
template <typename L, typename Result, typename... Operands>
concept LogicOver = HasOutputDim<L> && VariableConcept<Result> && (VariableConcept<Operands> && ...) &&
    requires(L logic, Result& result, const Result& frozen_result) {
        { logic.forward(result, std::declval<const Operands&>()...) } -> std::same_as<void>;
        { logic.backward(frozen_result, std::declval<Operands&>()...) } -> std::same_as<void>;
    };

// operands a factory accepts: differentiable, and all of one scalar type
template <typename First, typename... Others>
concept OperandsOfOneScalar = DifferentiableVariableConcept<First> && (DifferentiableVariableConcept<Others> && ...) &&
    (std::is_same_v<typename First::value_type, typename Others::value_type> && ...);

}  // namespace detail

template <typename L, typename Input, typename Output>
concept UnaryLogicConcept = detail::LogicOver<L, Output, Input>;

template <typename L, typename Input1, typename Input2, typename Output>
concept BinaryLogicConcept = detail::LogicOver<L, Output, Input1, Input2>;

template <typename L, typename Input1, typename Input2, typename Input3, typename Output>
concept TernaryLogicConcept = detail::LogicOver<L, Output, Input1, Input2, Input3>;

template <typename Input>
concept UnaryLogicParameterConcept = detail::OperandsOfOneScalar<Input>;

template <typename Input1, typename Input2>
concept BinaryLogicParameterConcept = detail::OperandsOfOneScalar<Input1, Input2>;

template <typename Input1, typename Input2, typename Input3>
concept TernaryLogicParameterConcept = detail::OperandsOfOneScalar<Input1, Input2, Input3>;

}  // namespace xyz_autodiff
