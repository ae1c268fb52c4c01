// xyz_autodiff/detail/traversal.cuh -- how a consumer walks into one operand during the two sweeps, shared by the
// interior nodes (operations/operation.cuh) and the broadcast view (operations/unary/broadcast.cuh).
//
// The DAG protocol (reference include/xyz_autodiff/operations/operation.cuh:53-83, unary/broadcast.cuh:55-101; SURVEY
// Appendix A, Q3): a consumer's forward() forwards every operand that is itself a node and registers as one of its
// consumers; a consumer's backward() first pushes its adjoint terms into the operands, then deregisters, and the
// operand continues the reverse sweep when its LAST consumer has deregistered.  Leaves take no part in either.
#pragma once

#include <cstdint>

#include "../concept/operation_node.cuh"
#include "config.cuh"

namespace xyz_autodiff::detail {

// number of consumers that have forwarded a node and not yet come back (8 bits, like the reference)
class ConsumerLedger {
public:
    XYZ_HD void check_in() const { ++outstanding_; }
    XYZ_HD bool check_out_was_last() const { return --outstanding_ == 0; }

private:
    mutable std::uint8_t outstanding_ = 0;
};

template <typename Operand>
XYZ_HD void sweep_down(Operand& x) {
    if constexpr (OperationNode<Operand>) {
        x.forward();
        x.increment_ref_count();
    }
}

template <typename Operand>
XYZ_HD void sweep_up(Operand& x) {
    if constexpr (OperationNode<Operand>) {
        if (x.decrement_ref_count_and_check()) x.backward();
    }
}

template <typename Operand, typename Step>
XYZ_HD void sweep_up_numerically(Operand& x, Step delta) {
    if constexpr (OperationNode<Operand>) {
        if (x.decrement_ref_count_and_check()) x.backward_numerical(delta);
    }
}

}  // namespace xyz_autodiff::detail
