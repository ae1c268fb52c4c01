// xyz_autodiff/operations/unary/broadcast.cuh -- size-1 -> size-N broadcast as a zero-storage VIEW:
// every index reads the operand's single value and every add_grad lands directly on the operand's
// single adjoint (so the N upstream adjoints are summed by the accumulation itself).
// Contract of reference include/xyz_autodiff/operations/unary/broadcast.cuh:10-109, including:
// zero_grad() clears the OPERAND's adjoint and run() seeds nothing.
#pragma once

#include <cstdint>

#include "../../concept/operation_node.cuh"
#include "../../concept/variable.cuh"
#include "../../detail/config.cuh"

namespace xyz_autodiff::op {

template <typename Input, std::size_t OutputSize>
    requires DifferentiableVariableConcept<Input> && (Input::size == 1) && (OutputSize > 1)
class BroadcastOperator {
public:
    using value_type = typename Input::value_type;
    static constexpr std::size_t size = OutputSize;

    XYZ_HD explicit BroadcastOperator(Input& source) : source_(source) {}

    XYZ_HD const value_type& operator[](std::size_t) const { return source_[0]; }
    XYZ_HD value_type& operator[](std::size_t) { return const_cast<value_type&>(source_[0]); }
    XYZ_HD const value_type& grad(std::size_t) const { return source_.grad(0); }
    XYZ_HD void add_grad(std::size_t, value_type v) { source_.add_grad(0, v); }
    XYZ_HD void zero_grad() { source_.zero_grad(); }

    XYZ_HD void forward() {
        if constexpr (OperationNode<Input>) {
            source_.forward();
            source_.increment_ref_count();
        }
    }
    XYZ_HD void backward() {
        if constexpr (OperationNode<Input>) {
            if (source_.decrement_ref_count_and_check()) source_.backward();
        }
    }
    XYZ_HD void backward_numerical(value_type delta = value_type(1e-5)) {
        if constexpr (OperationNode<Input>) {
            if (source_.decrement_ref_count_and_check()) source_.backward_numerical(delta);
        }
    }
    XYZ_HD void run() {
        forward();
        backward();
    }
    XYZ_HD void run_numerical(value_type delta = value_type(1e-5)) {
        forward();
        backward_numerical(delta);
    }

    XYZ_HD void increment_ref_count() const { ++pending_consumers_; }
    XYZ_HD bool decrement_ref_count_and_check() const { return --pending_consumers_ == 0; }

private:
    Input& source_;
    mutable std::uint8_t pending_consumers_ = 0;
};

template <std::size_t OutputSize, typename Input>
    requires DifferentiableVariableConcept<Input> && (Input::size == 1) && (OutputSize > 1)
XYZ_HD auto broadcast(Input& source) {
    return BroadcastOperator<Input, OutputSize>(source);
}

}  // namespace xyz_autodiff::op
