// tests/csrc/api_probe_cuda.cu -- TEST INFRASTRUCTURE: THIS repo's include/xyz_autodiff inside __global__
// kernels on sm_100a, driven by the same tests/csrc/api_eval.inc, plus the reference's concurrency
// known-answer tests (global and shared-memory accumulation) and the accumulate.cuh helpers.
// Exports cuda_* entry points taking HOST pointers.
#include "api_headers.inc"

#define API_FN __host__ __device__ inline
#include "api_eval.inc"

#include <cstdio>
#include <cstring>
#include <vector>

namespace {

struct EvalArgs {
    int op, aux, n1, n2, nout, rc;
    double cst;
    double in1[16], in2[16], gout[16], out[16], gin1[16], gin2[16];
};

template <class T>
__global__ void eval_kernel(EvalArgs* a) {
    T in1[16], in2[16], gout[16], out[16], gin1[16], gin2[16];
    for (int i = 0; i < 16; ++i) {
        in1[i] = T(a->in1[i]); in2[i] = T(a->in2[i]); gout[i] = T(a->gout[i]);
        out[i] = gin1[i] = gin2[i] = T(0);
    }
    int nout = 0;
    a->rc = api_eval::eval_op<T>(a->op, a->aux, in1, a->n1, in2, a->n2, T(a->cst), gout, out, &nout, gin1, gin2);
    a->nout = nout;
    for (int i = 0; i < 16; ++i) { a->out[i] = out[i]; a->gin1[i] = gin1[i]; a->gin2[i] = gin2[i]; }
}

template <class T>
int eval_on_device(int op, int aux, const T* in1, int n1, const T* in2, int n2, T cst, const T* gout, T* out, int* nout,
                   T* gin1, T* gin2) {
    EvalArgs h{};
    h.op = op; h.aux = aux; h.n1 = n1; h.n2 = n2; h.cst = cst;
    for (int i = 0; i < n1 && i < 16; ++i) h.in1[i] = in1[i];
    for (int i = 0; i < n2 && i < 16; ++i) h.in2[i] = in2[i];
    for (int i = 0; i < 16; ++i) h.gout[i] = gout[i];
    EvalArgs* d = nullptr;
    if (cudaMalloc(&d, sizeof(EvalArgs)) != cudaSuccess) return -100;
    cudaMemcpy(d, &h, sizeof(h), cudaMemcpyHostToDevice);
    eval_kernel<T><<<1, 1>>>(d);
    const cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return -101;
    *nout = h.nout;
    for (int i = 0; i < 16; ++i) { out[i] = T(h.out[i]); gin1[i] = T(h.gin1[i]); gin2[i] = T(h.gin2[i]); }
    return h.rc;
}

struct KatArgs {
    int which, n;
    double in[16];
    double res[64];
    float fres[64];
    double scratch[8];  // global memory for VariableRef leaves
};

__global__ void kat_kernel(KatArgs* a) {
    switch (a->which) {
        case 0: a->n = api_eval::kat_dag(a->res, a->scratch); break;
        case 1: a->n = api_eval::kat_shared_subgraph(a->res); break;
        case 2: a->n = api_eval::kat_broadcast(a->res); break;
        case 3: a->n = api_eval::kat_chain(a->in[0], a->in[1], a->in[2], a->in[3], a->res); break;
        case 4: a->n = api_eval::kat_operators(a->in, a->in + 2, a->in + 4, a->in + 6, a->res); break;
        case 5: a->n = api_eval::kat_lsq_point(a->in, a->in[4], a->in[5], a->in[6], a->in[7], a->res, a->scratch); break;
        case 6: a->n = api_eval::kat_splat_pair(a->in, a->res); break;
        case 7: a->n = api_eval::kat_math<double>(a->in[0], a->res); break;
        case 8: a->n = api_eval::kat_math<float>(float(a->in[0]), a->fres); break;
        case 9: a->n = api_eval::kat_matrices(a->fres); break;
        case 10: a->n = api_eval::kat_broadcast_logic(a->in[0], a->in + 1, a->in[5], a->in + 6, a->res); break;
        case 11: a->n = api_eval::kat_const_array(a->fres); break;
        case 12: a->n = api_eval::kat_variable(a->res, a->scratch); break;
        default: a->n = -1;
    }
}

// reference tests/test_parallel_gradient_accumulation.cu:25-49: every thread adds 1.0, 1.0 and
// (1 + tid/1000) + (2 + tid/1000) to three shared fp64 adjoints through VariableRef::add_grad
__global__ void global_accumulation_kernel(double* values, double* grads, std::size_t n) {
    const std::size_t tid = blockIdx.x * static_cast<std::size_t>(blockDim.x) + threadIdx.x;
    if (tid >= n) return;
    xyz_autodiff::VariableRef<1, double> x(values + 0, grads + 0), y(values + 1, grads + 1), r(values + 2, grads + 2);
    x.add_grad(0, 1.0);
    y.add_grad(0, 1.0);
    r.add_grad(0, (1.0 + tid * 0.001) + (2.0 + tid * 0.001));
}

// reference tests/test_shared_memory_atomic.cu:29-64 / :119-151: VariableRef over __shared__ memory;
// each thread adds +1 and +2, thread 0 publishes the block's totals
__global__ void shared_accumulation_kernel(float* out) {
    __shared__ float s_val[2];
    __shared__ float s_grad[2];
    if (threadIdx.x == 0) { s_val[0] = s_val[1] = 0.f; s_grad[0] = s_grad[1] = 0.f; }
    __syncthreads();
    xyz_autodiff::VariableRef<2, float> ref(s_val, s_grad);
    ref.add_grad(0, 1.0f);
    ref.add_grad(1, 2.0f);
    __syncthreads();
    if (threadIdx.x == 0) { out[2 * blockIdx.x] = s_grad[0]; out[2 * blockIdx.x + 1] = s_grad[1]; }
}

// accumulate.cuh: the least-squares graph per thread with RegisterLeaf parameters and one RED per CTA
__global__ void __launch_bounds__(256) lsq_register_leaf_kernel(const double* data, long long n, const double* values,
                                                                double* grads) {
    using namespace xyz_autodiff;
    __shared__ double scratch[8 * 4];
    accum::RegisterLeaf<1, double> a(values + 0), b(values + 1), c(values + 2), d(values + 3);
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
        const double x1 = data[3 * i], x2 = data[3 * i + 1], yt = data[3 * i + 2];
        auto u = op::sub_constant(a, x1);
        auto u2 = op::squared(u);
        auto v = op::sub_constant(c, x2);
        auto v2 = op::squared(v);
        auto bv2 = op::mul(b, v2);
        auto s = op::add(u2, bv2);
        auto yp = op::add(s, d);
        auto r = op::sub_constant(yp, yt);
        auto loss = op::squared(r);
        loss.run();
    }
    double g[4] = {a.grad(0), b.grad(0), c.grad(0), d.grad(0)};
    accum::block_accumulate(g, grads, scratch);
}

__global__ void __launch_bounds__(128) block_accumulate_f32_kernel(const float* vals, int n, float* target) {
    __shared__ float scratch[4 * 8];
    float v[8];
    for (int k = 0; k < 8; ++k) v[k] = 0.f;
    for (int i = blockIdx.x * 128 + threadIdx.x; i < n; i += gridDim.x * 128)
        for (int k = 0; k < 8; ++k) v[k] += vals[i] * float(k + 1);
    xyz_autodiff::accum::block_accumulate(v, target, scratch);
}

}  // namespace

extern "C" {
int cuda_eval_op_f64(int op, int aux, const double* in1, int n1, const double* in2, int n2, double cst,
                     const double* gout, double* out, int* nout, double* gin1, double* gin2) {
    return eval_on_device<double>(op, aux, in1, n1, in2, n2, cst, gout, out, nout, gin1, gin2);
}
int cuda_eval_op_f32(int op, int aux, const float* in1, int n1, const float* in2, int n2, float cst, const float* gout,
                     float* out, int* nout, float* gin1, float* gin2) {
    return eval_on_device<float>(op, aux, in1, n1, in2, n2, cst, gout, out, nout, gin1, gin2);
}
// which: 0 dag, 1 shared_subgraph, 2 broadcast, 3 chain, 4 operators, 5 lsq_point, 6 splat_pair, 7 math f64,
// 8 math f32 (res as float), 9 matrices (float).  in: up to 16 doubles.  Returns the count.
int cuda_kat(int which, const double* in, double* res, float* fres) {
    KatArgs h{};
    h.which = which;
    if (in) std::memcpy(h.in, in, sizeof(h.in));
    KatArgs* d = nullptr;
    if (cudaMalloc(&d, sizeof(KatArgs)) != cudaSuccess) return -100;
    cudaMemcpy(d, &h, sizeof(h), cudaMemcpyHostToDevice);
    kat_kernel<<<1, 1>>>(d);
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) fprintf(stderr, "kat_kernel(%d): %s\n", which, cudaGetErrorString(e));
    cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return -101;
    if (res) std::memcpy(res, h.res, sizeof(h.res));
    if (fres) std::memcpy(fres, h.fres, sizeof(h.fres));
    return h.n;
}
// grads_out[3]
int cuda_global_accumulation(long long n, int block, double* grads_out) {
    auto buf = makeCudaUniqueArray<double>(6);
    cudaMemset(buf.get(), 0, 6 * sizeof(double));
    global_accumulation_kernel<<<static_cast<unsigned>((n + block - 1) / block), block>>>(buf.get(), buf.get() + 3,
                                                                                          static_cast<std::size_t>(n));
    if (cudaDeviceSynchronize() != cudaSuccess) return -101;
    cudaMemcpy(grads_out, buf.get() + 3, 3 * sizeof(double), cudaMemcpyDeviceToHost);
    return 0;
}
// out[2 * blocks]
int cuda_shared_accumulation(int blocks, int threads, float* out) {
    auto buf = makeCudaUniqueArray<float>(2 * blocks);
    shared_accumulation_kernel<<<blocks, threads>>>(buf.get());
    if (cudaDeviceSynchronize() != cudaSuccess) return -101;
    cudaMemcpy(out, buf.get(), 2 * blocks * sizeof(float), cudaMemcpyDeviceToHost);
    return 0;
}
// data: n x 3 doubles (host), values[4] -> grads[4] (host)
int cuda_lsq_register_leaf(const double* data, long long n, const double* values, double* grads) {
    auto d = makeCudaUniqueArray<double>(3 * n);
    auto v = makeCudaUniqueArray<double>(8);
    cudaMemcpy(d.get(), data, 3 * n * sizeof(double), cudaMemcpyHostToDevice);
    cudaMemset(v.get(), 0, 8 * sizeof(double));
    cudaMemcpy(v.get(), values, 4 * sizeof(double), cudaMemcpyHostToDevice);
    lsq_register_leaf_kernel<<<148, 256>>>(d.get(), n, v.get(), v.get() + 4);
    if (cudaDeviceSynchronize() != cudaSuccess) return -101;
    cudaMemcpy(grads, v.get() + 4, 4 * sizeof(double), cudaMemcpyDeviceToHost);
    return 0;
}
// target[8] (host) receives sum(vals) * (k + 1)
int cuda_block_accumulate_f32(const float* vals, int n, float* target) {
    auto d = makeCudaUniqueArray<float>(n);
    auto t = makeCudaUniqueArray<float>(8);
    cudaMemcpy(d.get(), vals, n * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemset(t.get(), 0, 8 * sizeof(float));
    block_accumulate_f32_kernel<<<64, 128>>>(d.get(), n, t.get());
    if (cudaDeviceSynchronize() != cudaSuccess) return -101;
    cudaMemcpy(target, t.get(), 8 * sizeof(float), cudaMemcpyDeviceToHost);
    return 0;
}
}
