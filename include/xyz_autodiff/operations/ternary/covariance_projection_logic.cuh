// xyz_autodiff/operations/ternary/covariance_projection_logic.cuh -- the covariance projection
//     S' = (J W) S (J W)^T        J: 2x3, W: 3x3 (both row-major), S: symmetric 3x3, S': symmetric 2x2
// as ONE node with three operands.  Not in the reference, which ships TernaryOperation
// (include/xyz_autodiff/operations/operation.cuh:382-623) without any Logic that uses it; composing the product from its
// op::matmul nodes (operations/binary/matmul_logic.cuh:13-81) needs W^T and J^T as separate leaves because there is no
// differentiable transpose and a shared sub-graph would be re-forwarded per consumer.  Here J, W and
// S are each ONE operand and receive their complete adjoint.
//
// Storage: S and S' are packed upper triangles (symmetric_matrix_view.cuh:24-29): S = {00 01 02 11 12 22},
// S' = {00 01 11}; the upstream adjoint of S' is the adjoint of those three stored numbers (S'_10 is not stored, so it
// carries no adjoint of its own).  Arithmetic and its order are those of the batched kernel
// xyz_covproj_fwd_bwd_f32 (csrc/covproj_kernels.cu), whose results this node reproduces.
#pragma once

#include "../operation.cuh"

namespace xyz_autodiff {
namespace op {

template <typename InputJ, typename InputW, typename InputS>
    requires TernaryLogicParameterConcept<InputJ, InputW, InputS> && (InputJ::size == 6) && (InputW::size == 9) &&
             (InputS::size == 6)
struct CovarianceProjectionLogic {
    using T = typename InputJ::value_type;
    static constexpr std::size_t outputDim = 3;
    using Output = Variable<3, T>;

    XYZ_HD void forward(Output& projected, const InputJ& J, const InputW& W, const InputS& S) const {
        T JW[6], JWS[6];
        products(J, W, S, JW, JWS);
        T p00 = T(0), p01 = T(0), p11 = T(0);
#pragma unroll
        for (std::size_t k = 0; k < 3; ++k) {
            p00 += JWS[k] * JW[k];
            p01 += JWS[k] * JW[3 + k];
            p11 += JWS[3 + k] * JW[3 + k];
        }
        projected[0] = p00;
        projected[1] = p01;
        projected[2] = p11;
    }

    XYZ_HD void backward(const Output& projected, InputJ& J, InputW& W, InputS& S) const {
        T JW[6], JWS[6];
        products(J, W, S, JW, JWS);  // nothing is cached between the sweeps
        const T Sf[9] = {S[0], S[1], S[2], S[1], S[3], S[4], S[2], S[4], S[5]};
        // adjoint of the 2x2 product P = (JWS)(JW)^T given the adjoints of its stored entries
        const T G[4] = {projected.grad(0), projected.grad(1), T(0), projected.grad(2)};
        T dU[6], dT_left[6], dT_right[6];
#pragma unroll
        for (std::size_t i = 0; i < 2; ++i)
#pragma unroll
            for (std::size_t k = 0; k < 3; ++k) {
                T via_u = T(0), via_t = T(0);
#pragma unroll
                for (std::size_t j = 0; j < 2; ++j) {
                    via_u += G[i * 2 + j] * JW[j * 3 + k];   // dU = G (JW)
                    via_t += G[j * 2 + i] * JWS[j * 3 + k];  // the (JW)^T factor: G^T (JWS)
                }
                dU[i * 3 + k] = via_u;
                dT_right[i * 3 + k] = via_t;
            }
#pragma unroll
        for (std::size_t i = 0; i < 2; ++i)
#pragma unroll
            for (std::size_t k = 0; k < 3; ++k) {
                T acc = T(0);
#pragma unroll
                for (std::size_t j = 0; j < 3; ++j) acc += dU[i * 3 + j] * Sf[k * 3 + j];  // through U = (JW) S
                dT_left[i * 3 + k] = acc;
            }
        // S: dS_full = (JW)^T dU, folded onto the packed upper triangle
        T dSf[9];
#pragma unroll
        for (std::size_t k = 0; k < 3; ++k)
#pragma unroll
            for (std::size_t j = 0; j < 3; ++j) {
                T acc = T(0);
#pragma unroll
                for (std::size_t i = 0; i < 2; ++i) acc += JW[i * 3 + k] * dU[i * 3 + j];
                dSf[k * 3 + j] = acc;
            }
        // J: (dT_left + dT_right) W^T ; W: J^T (dT_left + dT_right) -- the two paths are summed separately first
#pragma unroll
        for (std::size_t i = 0; i < 2; ++i)
#pragma unroll
            for (std::size_t k = 0; k < 3; ++k) {
                T left = T(0), right = T(0);
#pragma unroll
                for (std::size_t j = 0; j < 3; ++j) {
                    left += dT_left[i * 3 + j] * W[k * 3 + j];
                    right += dT_right[i * 3 + j] * W[k * 3 + j];
                }
                J.add_grad(i * 3 + k, left + right);
            }
#pragma unroll
        for (std::size_t k = 0; k < 3; ++k)
#pragma unroll
            for (std::size_t j = 0; j < 3; ++j) {
                T left = T(0), right = T(0);
#pragma unroll
                for (std::size_t i = 0; i < 2; ++i) {
                    left += J[i * 3 + k] * dT_left[i * 3 + j];
                    right += J[i * 3 + k] * dT_right[i * 3 + j];
                }
                W.add_grad(k * 3 + j, left + right);
            }
        S.add_grad(0, dSf[0]);
        S.add_grad(1, dSf[1] + dSf[3]);
        S.add_grad(2, dSf[2] + dSf[6]);
        S.add_grad(3, dSf[4]);
        S.add_grad(4, dSf[5] + dSf[7]);
        S.add_grad(5, dSf[8]);
    }

private:
    // JW = J W (2x3) and JWS = (J W) S_full (2x3)
    XYZ_HD static void products(const InputJ& J, const InputW& W, const InputS& S, T (&JW)[6], T (&JWS)[6]) {
        const T Sf[9] = {S[0], S[1], S[2], S[1], S[3], S[4], S[2], S[4], S[5]};
#pragma unroll
        for (std::size_t i = 0; i < 2; ++i)
#pragma unroll
            for (std::size_t j = 0; j < 3; ++j) {
                T acc = T(0);
#pragma unroll
                for (std::size_t k = 0; k < 3; ++k) acc += J[i * 3 + k] * W[k * 3 + j];
                JW[i * 3 + j] = acc;
            }
#pragma unroll
        for (std::size_t i = 0; i < 2; ++i)
#pragma unroll
            for (std::size_t j = 0; j < 3; ++j) {
                T acc = T(0);
#pragma unroll
                for (std::size_t k = 0; k < 3; ++k) acc += JW[i * 3 + k] * Sf[k * 3 + j];
                JWS[i * 3 + j] = acc;
            }
    }
};

template <typename InputJ, typename InputW, typename InputS>
    requires TernaryLogicParameterConcept<InputJ, InputW, InputS> && (InputJ::size == 6) && (InputW::size == 9) &&
             (InputS::size == 6)
XYZ_HD auto covariance_projection(InputJ& J, InputW& W, InputS& S) {
    using Logic = CovarianceProjectionLogic<InputJ, InputW, InputS>;
    return TernaryOperation<Logic::outputDim, Logic, InputJ, InputW, InputS>(Logic{}, J, W, S);
}

}  // namespace op
}  // namespace xyz_autodiff
