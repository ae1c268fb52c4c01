// xyz_autodiff/operations/unary/broadcast.cuh -- size-1 -> size-N broadcast as a zero-storage VIEW:
// every index reads the operand's single value and every add_grad lands directly on the operand's
// single adjoint (so the N upstream adjoints are summed by the accumulation itself).
// Contract of reference include/xyz_autodiff/operations/unary/broadcast.cuh:10-109, including:
// zero_grad() clears the OPERAND's adjoint and run() seeds nothing.
#pragma once

#include "../../concept/variable.cuh"
#include "../../detail/config.cuh"
#include "../../detail/traversal.cuh"

namespace xyz_autodiff::detail {
// a one-component differentiable operand fanned out to more than one component
template <typename Operand, std::size_t FanOut>
concept ScalarFanOut = DifferentiableVariableConcept<Operand> && (Operand::size == 1) && (FanOut > 1);
}  // namespace xyz_autodiff::detail

namespace xyz_autodiff::op {

template <typename Input, std::size_t OutputSize>
    requires detail::ScalarFanOut<Input, OutputSize>
class BroadcastOperator {
    using Scalar = typename Input::value_type;

public:
    using value_type = Scalar;
    static constexpr std::size_t size = OutputSize;

    XYZ_HD explicit BroadcastOperator(Input& scalar_operand) : operand_(scalar_operand) {}

    // ---- the view: N aliases of component 0 of the operand
    XYZ_HD Scalar& operator[](std::size_t) { return const_cast<Scalar&>(operand_[0]); }
    XYZ_HD const Scalar& operator[](std::size_t) const { return operand_[0]; }
    XYZ_HD const Scalar& grad(std::size_t) const { return operand_.grad(0); }
    XYZ_HD void add_grad(std::size_t, Scalar term) { operand_.add_grad(0, term); }
    XYZ_HD void zero_grad() { operand_.zero_grad(); }

    // ---- the node: no Logic of its own, it only relays the two sweeps
    XYZ_HD void forward() { detail::sweep_down(operand_); }
    XYZ_HD void backward() { detail::sweep_up(operand_); }
    XYZ_HD void backward_numerical(Scalar delta = Scalar(1e-5)) { detail::sweep_up_numerically(operand_, delta); }
    XYZ_HD void run() {
        forward();
        backward();
    }
    XYZ_HD void run_numerical(Scalar delta = Scalar(1e-5)) {
        forward();
        backward_numerical(delta);
    }
    XYZ_HD void increment_ref_count() const { consumers_.check_in(); }
    XYZ_HD bool decrement_ref_count_and_check() const { return consumers_.check_out_was_last(); }

private:
    Input& operand_;
    detail::ConsumerLedger consumers_;
};

template <std::size_t OutputSize, typename Input>
    requires detail::ScalarFanOut<Input, OutputSize>
XYZ_HD auto broadcast(Input& scalar_operand) {
    return BroadcastOperator<Input, OutputSize>(scalar_operand);
}

}  // namespace xyz_autodiff::op
