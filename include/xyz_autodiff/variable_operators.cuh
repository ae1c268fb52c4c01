// xyz_autodiff/variable_operators.cuh -- operator sugar over the op:: factories.
// Contract of reference include/xyz_autodiff/variable_operators.cuh:23-79.  Operands are taken by
// non-const lvalue reference: a node keeps a reference to them, so temporaries are rejected at
// compile time (name every intermediate).
#pragma once

#include "concept/variable.cuh"
#include "operations/binary/add_logic.cuh"
#include "operations/binary/div_logic.cuh"
#include "operations/binary/mul_logic.cuh"
#include "operations/binary/sub_logic.cuh"
#include "operations/unary/add_constant_logic.cuh"
#include "operations/unary/const_array_add_logic.cuh"
#include "operations/unary/const_array_sub_logic.cuh"
#include "operations/unary/div_constant_logic.cuh"
#include "operations/unary/mul_constant_logic.cuh"
#include "operations/unary/sub_constant_logic.cuh"

namespace xyz_autodiff {

// variable (op) scalar constant
template <DifferentiableVariableConcept Var>
XYZ_HD auto operator+(Var& v, const typename Var::value_type& c) { return op::add_constant(v, c); }
template <DifferentiableVariableConcept Var>
XYZ_HD auto operator-(Var& v, const typename Var::value_type& c) { return op::sub_constant(v, c); }
template <DifferentiableVariableConcept Var>
XYZ_HD auto operator*(Var& v, const typename Var::value_type& c) { return op::mul_constant(v, c); }
template <DifferentiableVariableConcept Var>
XYZ_HD auto operator/(Var& v, const typename Var::value_type& c) { return op::div_constant(v, c); }

// variable (op) variable, element-wise
#define XYZ_VAR_VAR_OPERATOR(SYMBOL, FACTORY)                                                                  \
    template <DifferentiableVariableConcept Var1, DifferentiableVariableConcept Var2>                          \
        requires(Var1::size == Var2::size) && std::same_as<typename Var1::value_type, typename Var2::value_type> \
    XYZ_HD auto operator SYMBOL(Var1& a, Var2& b) {                                                            \
        return op::FACTORY(a, b);                                                                              \
    }
XYZ_VAR_VAR_OPERATOR(+, add)
XYZ_VAR_VAR_OPERATOR(-, sub)
XYZ_VAR_VAR_OPERATOR(*, mul)
XYZ_VAR_VAR_OPERATOR(/, div)
#undef XYZ_VAR_VAR_OPERATOR

// variable +/- constant array (array held by reference)
template <DifferentiableVariableConcept Var, typename ConstArray>
    requires op::ArrayLikeConcept<ConstArray>
XYZ_HD auto operator+(Var& v, const ConstArray& c) { return op::const_add(v, c); }
template <DifferentiableVariableConcept Var, typename ConstArray>
    requires op::ArrayLikeConcept<ConstArray>
XYZ_HD auto operator-(Var& v, const ConstArray& c) { return op::const_sub(v, c); }

}  // namespace xyz_autodiff
