// xyz_autodiff/operations/operation.cuh -- the graph runtime: interior nodes of a statically typed
// expression DAG that lives entirely in one thread's registers.
//
// Same protocol as the reference (include/xyz_autodiff/operations/operation.cuh:14-623; semantics
// catalogued in SURVEY.md Appendix A/B.1), one implementation for every arity:
//
//   forward()   recurse into operands that are themselves nodes (bumping their reference count once
//               per consumer), then logic.forward(output, operands...)
//   backward()  logic.backward(output, operands...) -- which calls operand.add_grad(i, v) -- then, operand by
//               operand in declaration order, continue into a node operand once its last consumer has
//               arrived (reference count reaches zero)
//   run()       forward(); zero this node's adjoint; seed EVERY output component with 1; backward()
//   *_numerical the same traversal with each node's local Jacobian estimated by central differences
//               of logic.forward (operands perturbed in place and restored)
//
// Nodes hold REFERENCES to their operands: every intermediate must be a named lvalue that outlives
// its consumers.  Values and adjoints are a Variable<OutputSize, T> member, i.e. registers once the
// fixed-trip loops are unrolled; there is no tape, no heap and no virtual dispatch.  Everything is
// __host__ __device__.
#pragma once

#include <cstdint>
#include <tuple>
#include <type_traits>
#include <utility>

#include "../concept/core_logic.cuh"
#include "../concept/operation_node.cuh"
#include "../concept/variable.cuh"
#include "../detail/config.cuh"
#include "../detail/traversal.cuh"
#include "../variable.cuh"

namespace xyz_autodiff {
namespace detail {

template <std::size_t I, typename Operand>
struct OperandSlot {
    Operand& ref;
    XYZ_HD explicit OperandSlot(Operand& r) : ref(r) {}
};

template <typename First, typename...>
struct first_type {
    using type = First;
};

template <typename Indices, std::size_t OutputSize, typename Logic, typename... Operands>
class GraphNode;

template <std::size_t... Is, std::size_t OutputSize, typename Logic, typename... Operands>
class GraphNode<std::index_sequence<Is...>, OutputSize, Logic, Operands...> : private OperandSlot<Is, Operands>... {
public:
    using value_type = typename first_type<Operands...>::type::value_type;
    using output_type = Variable<OutputSize, value_type>;
    static constexpr std::size_t output_size = OutputSize;
    static constexpr std::size_t size = OutputSize;

    GraphNode() = delete;
    GraphNode(const GraphNode&) = delete;
    GraphNode& operator=(const GraphNode&) = delete;
    GraphNode& operator=(GraphNode&&) = delete;

    // Moving re-seats the operand references and carries the output; the consumer count restarts
    // at zero (a moved node has not been forwarded by anybody yet).
    XYZ_HD GraphNode(GraphNode&& other) noexcept
        : OperandSlot<Is, Operands>(static_cast<OperandSlot<Is, Operands>&>(other).ref)...,
          logic_(other.logic_),
          output_(other.output_),
          consumers_() {}

    // Construction does not evaluate anything: the output starts at zero until forward() runs.
    XYZ_HD GraphNode(const Logic& logic, Operands&... operands)
        : OperandSlot<Is, Operands>(operands)..., logic_(logic), output_(), consumers_() {}

    XYZ_HD void forward() {
        (sweep_down(operand<Is>()), ...);
        logic_.forward(output_, operand<Is>()...);
    }

    XYZ_HD void backward() {
        logic_.backward(output_, operand<Is>()...);
        (sweep_up(operand<Is>()), ...);
    }

    XYZ_HD void backward_numerical(const value_type delta = value_type(1e-5)) {
        const output_type saved = output_;
        (numerical_operand(operand<Is>(), saved, delta), ...);
        (sweep_up_numerically(operand<Is>(), delta), ...);
    }

    XYZ_HD void run() {
        forward();
        seed_ones();
        backward();
    }

    XYZ_HD void run_numerical(const value_type delta = value_type(1e-5)) {
        forward();
        seed_ones();
        backward_numerical(delta);
    }

    // DAG bookkeeping: one increment per consumer's forward(), one decrement per consumer's backward().
    XYZ_HD void increment_ref_count() const { consumers_.check_in(); }
    XYZ_HD bool decrement_ref_count_and_check() const { return consumers_.check_out_was_last(); }

    // The node is itself a (differentiable) variable: consumers read its output and add to its adjoint.
    XYZ_HD output_type& output() { return output_; }
    XYZ_HD const output_type& output() const { return output_; }
    XYZ_HD value_type& operator[](std::size_t i) { return output_[i]; }
    XYZ_HD const value_type& operator[](std::size_t i) const { return output_[i]; }
    XYZ_HD const value_type& grad(std::size_t i) const { return output_.grad(i); }
    XYZ_HD void add_grad(std::size_t i, value_type v) { output_.add_grad(i, v); }
    XYZ_HD value_type* data() { return output_.data(); }
    XYZ_HD const value_type* data() const { return output_.data(); }
    XYZ_HD value_type* grad() { return output_.grad(); }
    XYZ_HD const value_type* grad() const { return output_.grad(); }
    XYZ_HD void zero_grad() { output_.zero_grad(); }  // this node only; leaves are the caller's business

protected:
    template <std::size_t I>
    XYZ_HD auto& operand() {
        using Slot = OperandSlot<I, std::tuple_element_t<I, std::tuple<Operands...>>>;
        return static_cast<Slot&>(*this).ref;
    }

private:
    XYZ_HD void seed_ones() {
        output_.zero_grad();
#pragma unroll
        for (std::size_t j = 0; j < OutputSize; ++j) output_.add_grad(j, value_type(1));
    }

    // central differences of logic.forward w.r.t. every component of one operand
    template <typename Operand>
    XYZ_HD void numerical_operand(Operand& x, const output_type& saved, value_type delta) {
        for (std::size_t i = 0; i < Operand::size; ++i) {
            const value_type original = x[i];
            x[i] = original + delta;
            logic_.forward(output_, operand<Is>()...);
            const output_type plus = output_;
            x[i] = original - delta;
            logic_.forward(output_, operand<Is>()...);
            const output_type minus = output_;
            x[i] = original;
            output_ = saved;
            for (std::size_t j = 0; j < OutputSize; ++j) {
                const value_type dj_di = (plus[j] - minus[j]) / (value_type(2) * delta);
                x.add_grad(i, output_.grad(j) * dj_di);
            }
        }
    }

    Logic logic_;
    output_type output_;
    ConsumerLedger consumers_;
};

}  // namespace detail

// 1 operand -> OutputSize values.  reference: operation.cuh:14-174
template <std::size_t OutputSize, typename Logic, typename Input>
    requires UnaryLogicConcept<Logic, Input, Variable<OutputSize, typename Input::value_type>>
class UnaryOperation : public detail::GraphNode<std::index_sequence<0>, OutputSize, Logic, Input> {
    using Base = detail::GraphNode<std::index_sequence<0>, OutputSize, Logic, Input>;

public:
    using input_type = Input;
    using Base::Base;
    XYZ_HD UnaryOperation(UnaryOperation&& other) noexcept : Base(static_cast<Base&&>(other)) {}
};

// 2 operands.  reference: operation.cuh:177-379
template <std::size_t OutputSize, typename Logic, typename Input1, typename Input2>
    requires BinaryLogicConcept<Logic, Input1, Input2, Variable<OutputSize, typename Input1::value_type>>
class BinaryOperation : public detail::GraphNode<std::index_sequence<0, 1>, OutputSize, Logic, Input1, Input2> {
    using Base = detail::GraphNode<std::index_sequence<0, 1>, OutputSize, Logic, Input1, Input2>;

public:
    using input1_type = Input1;
    using input2_type = Input2;
    using Base::Base;
    XYZ_HD BinaryOperation(BinaryOperation&& other) noexcept : Base(static_cast<Base&&>(other)) {}
};

// 3 operands.  reference: operation.cuh:382-623
template <std::size_t OutputSize, typename Logic, typename Input1, typename Input2, typename Input3>
    requires TernaryLogicConcept<Logic, Input1, Input2, Input3, Variable<OutputSize, typename Input1::value_type>>
class TernaryOperation
    : public detail::GraphNode<std::index_sequence<0, 1, 2>, OutputSize, Logic, Input1, Input2, Input3> {
    using Base = detail::GraphNode<std::index_sequence<0, 1, 2>, OutputSize, Logic, Input1, Input2, Input3>;

public:
    using input1_type = Input1;
    using input2_type = Input2;
    using input3_type = Input3;
    using Base::Base;
    XYZ_HD TernaryOperation(TernaryOperation&& other) noexcept : Base(static_cast<Base&&>(other)) {}
};

}  // namespace xyz_autodiff
