// xyz_autodiff/operations/unary/broadcast_logic.cuh -- the MATERIALISING broadcast (copies the value
// into N outputs; backward sums the N adjoints into the operand).  Lives in namespace xyz_autodiff, not
// xyz_autodiff::op, like reference include/xyz_autodiff/operations/unary/broadcast_logic.cuh:11-46.
#pragma once

#include "../../detail/pointwise.cuh"
#include "../operation.cuh"

namespace xyz_autodiff {

template <std::size_t OutputDim>
struct BroadcastLogic {
    static constexpr std::size_t outputDim = OutputDim;

    template <typename Copies, typename Scalar1>
        requires(Scalar1::size == 1)
    XYZ_HD void forward(Copies& y, const Scalar1& x) const {
        const typename Scalar1::value_type value = x[0];
#pragma unroll
        for (std::size_t n = 0; n < OutputDim; ++n) y[n] = value;
    }

    template <typename Copies, typename Scalar1>
        requires(Scalar1::size == 1)
    XYZ_HD void backward(const Copies& y, Scalar1& x) const {
        typename Scalar1::value_type gathered = 0;
#pragma unroll
        for (std::size_t n = 0; n < OutputDim; ++n) gathered += y.grad(n);
        x.add_grad(0, gathered);  // ONE accumulation into the operand, whatever N is
    }
};

template <std::size_t OutputDim, DifferentiableVariableConcept Input>
    requires(Input::size == 1)
XYZ_HD auto broadcast(Input& x) {
    return detail::make_unary_node<BroadcastLogic<OutputDim>>(x);
}

}  // namespace xyz_autodiff
