// tests/csrc/batched_probe.cu -- TEST INFRASTRUCTURE: two user graphs written with the public op:: API and run
// through xyz_autodiff::batched::for_each (include/xyz_autodiff/batched.cuh) on the GPU:
//   * the least-squares graph of the reference's optimisation example (fp64, 4 shared parameters + the loss sum,
//     examples/optimization/tests/test_linear_regression_gradient.cu:44-78);
//   * the matrix chain S' = (J W) S (J W)^T with ONE shared W built from op::matmul nodes as a tree (fp32; W and
//     W^T are separate shared leaves, J and J^T separate per-element leaves -- the reference has no differentiable
//     transpose).
// Each entry point runs the kernel `reps` times and returns the average milliseconds (CUDA events), or -1.
#include <cuda_runtime.h>

#include <xyz_autodiff/xyz_autodiff.cuh>
#include <xyz_autodiff/accumulate.cuh>
#include <xyz_autodiff/batched.cuh>

using namespace xyz_autodiff;

struct LsqPoint {
    double x1, x2, y;
};

struct LsqGraph {
    // shared = {a, b, c, d, loss accumulator}
    __device__ void operator()(const LsqPoint& p, batched::NoOutput&, accum::RegisterLeaf<5, double>& shared) const {
        auto a = accum::slice<0, 1>(shared);
        auto b = accum::slice<1, 1>(shared);
        auto c = accum::slice<2, 1>(shared);
        auto d = accum::slice<3, 1>(shared);
        auto x1_minus_a = op::sub_constant(a, p.x1);
        auto x1_term = op::squared(x1_minus_a);
        auto x2_minus_c = op::sub_constant(c, p.x2);
        auto x2_squared = op::squared(x2_minus_c);
        auto x2_term = op::mul(b, x2_squared);
        auto combined = op::add(x1_term, x2_term);
        auto y_pred = op::add(combined, d);
        auto y_diff = op::sub_constant(y_pred, p.y);
        auto loss = op::squared(y_diff);
        loss.run();
        shared.add_grad(4, loss[0]);
    }
};

struct ChainIn {
    float J[6], S[6], g[3];
};
struct ChainOut {
    float out[3], gJ[6], gS[6];
};

struct ChainGraph {
    // shared = {W (9, row-major), W^T (9)}
    __device__ void operator()(const ChainIn& x, ChainOut& y, accum::RegisterLeaf<18, float>& shared) const {
        auto W = accum::slice<0, 9>(shared);
        auto Wt = accum::slice<9, 9>(shared);
        float Jt[6];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) Jt[j * 2 + i] = x.J[i * 3 + j];
        const float Sf[9] = {x.S[0], x.S[1], x.S[2], x.S[1], x.S[3], x.S[4], x.S[2], x.S[4], x.S[5]};
        Variable<6, float> vJ(x.J), vJt(Jt);
        Variable<9, float> vS(Sf);
        auto T = op::matmul<2, 3, 3>(vJ, W);
        auto U = op::matmul<2, 3, 3>(T, vS);
        auto Tt = op::matmul<3, 3, 2>(Wt, vJt);
        auto P = op::matmul<2, 3, 2>(U, Tt);
        P.forward();
        y.out[0] = P[0];
        y.out[1] = P[1];
        y.out[2] = P[3];
        P.zero_grad();
        P.add_grad(0, x.g[0]);
        P.add_grad(1, x.g[1]);
        P.add_grad(3, x.g[2]);
        P.backward();
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) y.gJ[i * 3 + j] = vJ.grad(i * 3 + j) + vJt.grad(j * 2 + i);
        y.gS[0] = vS.grad(0);
        y.gS[1] = vS.grad(1) + vS.grad(3);
        y.gS[2] = vS.grad(2) + vS.grad(6);
        y.gS[3] = vS.grad(4);
        y.gS[4] = vS.grad(5) + vS.grad(7);
        y.gS[5] = vS.grad(8);
    }
};

// the same chain as ONE ternary node (operations/ternary/covariance_projection_logic.cuh): shared = {W (9)}
struct ChainTernaryGraph {
    __device__ void operator()(const ChainIn& x, ChainOut& y, accum::RegisterLeaf<9, float>& shared) const {
        Variable<6, float> vJ(x.J), vS(x.S);
        auto P = op::covariance_projection(vJ, shared, vS);
        P.forward();
#pragma unroll
        for (int i = 0; i < 3; ++i) y.out[i] = P[i];
        P.zero_grad();
#pragma unroll
        for (int i = 0; i < 3; ++i) P.add_grad(i, x.g[i]);
        P.backward();
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            y.gJ[i] = vJ.grad(i);
            y.gS[i] = vS.grad(i);
        }
    }
};

namespace {
batched::Workspace g_ws;

template <class Fn>
float timed(Fn launch, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) {
        if (launch() != cudaSuccess) return -1.f;
    }
    cudaEventRecord(b);
    if (cudaEventSynchronize(b) != cudaSuccess) return -1.f;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    return ms / reps;
}
}  // namespace

extern "C" {

// data: n x {x1, x2, y}; values: {a, b, c, d, unused}; grads: {da, db, dc, dd, loss sum} accumulated `reps` times
float batched_lsq(const double* data, long long n, const double* values5, double* grads5, int reps) {
    return timed([&] {
        return batched::for_each<LsqPoint, batched::NoOutput, 5, double>(reinterpret_cast<const LsqPoint*>(data), nullptr, n,
                                                                         values5, grads5, LsqGraph{}, g_ws);
    }, reps);
}

// in: n x ChainIn (60 B rows), out: n x ChainOut; w18 = {W, W^T}; gw18 accumulated `reps` times
float batched_chain(const float* in, float* out, long long n, const float* w18, float* gw18, int reps) {
    return timed([&] {
        return batched::for_each<ChainIn, ChainOut, 18, float>(reinterpret_cast<const ChainIn*>(in),
                                                               reinterpret_cast<ChainOut*>(out), n, w18, gw18, ChainGraph{}, g_ws);
    }, reps);
}

// as batched_chain, with the chain as one op::covariance_projection node; w9 = W, gw9 accumulated `reps` times
float batched_chain_ternary(const float* in, float* out, long long n, const float* w9, float* gw9, int reps) {
    return timed([&] {
        return batched::for_each<ChainIn, ChainOut, 9, float>(reinterpret_cast<const ChainIn*>(in),
                                                              reinterpret_cast<ChainOut*>(out), n, w9, gw9, ChainTernaryGraph{}, g_ws);
    }, reps);
}

}  // extern "C"
