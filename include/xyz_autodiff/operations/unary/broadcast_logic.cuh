// xyz_autodiff/operations/unary/broadcast_logic.cuh -- the MATERIALISING broadcast (copies the value
// into N outputs; backward sums the N adjoints into the operand).  Lives in namespace xyz_autodiff, not
// xyz_autodiff::op, like reference include/xyz_autodiff/operations/unary/broadcast_logic.cuh:11-46.
#pragma once

#include "../operation.cuh"

namespace xyz_autodiff {

template <std::size_t OutputDim>
struct BroadcastLogic {
    static constexpr std::size_t outputDim = OutputDim;

    template <typename Output, typename Input>
    XYZ_HD void forward(Output& y, const Input& x) const {
        static_assert(Input::size == 1, "broadcast takes a size-1 operand");
#pragma unroll
        for (std::size_t i = 0; i < OutputDim; ++i) y[i] = x[0];
    }

    template <typename Output, typename Input>
    XYZ_HD void backward(const Output& y, Input& x) const {
        static_assert(Input::size == 1, "broadcast takes a size-1 operand");
        typename Input::value_type total = 0;
#pragma unroll
        for (std::size_t i = 0; i < OutputDim; ++i) total += y.grad(i);
        x.add_grad(0, total);
    }
};

template <std::size_t OutputDim, DifferentiableVariableConcept Input>
    requires(Input::size == 1)
XYZ_HD auto broadcast(Input& x) {
    return UnaryOperation<OutputDim, BroadcastLogic<OutputDim>, Input>(BroadcastLogic<OutputDim>{}, x);
}

}  // namespace xyz_autodiff
