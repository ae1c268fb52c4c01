// dev/addgrad_lab.cu -- development harness (not shipped): VariableRef::add_grad from 2^24 threads in the reference's
// own pattern (tests/test_parallel_gradient_accumulation.cu:32-43), with and without the warp aggregation of
// include/xyz_autodiff/detail/config.cuh.  Build twice:
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++20 -O3 --expt-relaxed-constexpr -Iinclude dev/addgrad_lab.cu -o dev/_build/addgrad_agg
//   nvcc ... -DXYZ_AUTODIFF_PLAIN_ATOMICS dev/addgrad_lab.cu -o dev/_build/addgrad_plain
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <xyz_autodiff/xyz_autodiff.cuh>
using namespace xyz_autodiff;

// pattern 0: every thread adds to the SAME 3 parameters (the reference's test); 1: parameter = tid mod K (warp-distinct);
// 2: parameter = hashed id (random, K = 1024); 3: the least-squares graph on VariableRef leaves of one parameter block
template <typename T>
__global__ void addgrad_kernel(T* values, T* grads, long long n, int k, int pattern, const int* ids) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    if (pattern == 0) {
        VariableRef<3, T> p(values, grads);
        p.add_grad(0, T(1));
        p.add_grad(1, T(1));
        p.add_grad(2, T(3) + T(0.002) * static_cast<T>(i & 1023));
    } else if (pattern == 1) {
        const int id = static_cast<int>(i % k);
        VariableRef<1, T> p(values + id, grads + id);
        p.add_grad(0, T(1));
    } else if (pattern == 2) {
        const int id = ids[i];
        VariableRef<1, T> p(values + id, grads + id);
        p.add_grad(0, T(1));
    } else {
        VariableRef<1, T> a(values + 0, grads + 0), b(values + 1, grads + 1), c(values + 2, grads + 2), d(values + 3, grads + 3);
        const T x1 = T(0.001) * static_cast<T>(i & 4095) - T(2), x2 = T(0.002) * static_cast<T>((i >> 3) & 2047) - T(2), yt = T(1.5);
        auto u = op::sub_constant(a, x1);
        auto u2 = op::squared(u);
        auto v = op::sub_constant(c, x2);
        auto v2 = op::squared(v);
        auto t = op::mul(b, v2);
        auto s = op::add(u2, t);
        auto pred = op::add(s, d);
        auto r = op::sub_constant(pred, yt);
        auto loss = op::squared(r);
        loss.run();
    }
}

template <typename T>
void run(const char* tname) {
    const long long n = 1LL << 24;
    const int k = 1024;
    T *values, *grads;
    int* ids;
    cudaMalloc(&values, k * sizeof(T));
    cudaMalloc(&grads, k * sizeof(T));
    cudaMalloc(&ids, n * sizeof(int));
    std::vector<int> h(n);
    unsigned long long s = 88172645463325252ull;
    for (long long i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = static_cast<int>(s % k); }
    cudaMemcpy(ids, h.data(), n * sizeof(int), cudaMemcpyHostToDevice);
    std::vector<T> hv(k, T(0.5));
    cudaMemcpy(values, hv.data(), k * sizeof(T), cudaMemcpyHostToDevice);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const char* names[4] = {"same 3 parameters", "tid mod 1024", "random of 1024", "least-squares graph, 4 shared leaves"};
    for (int pattern = 0; pattern < 4; ++pattern) {
        float best = 1e30f;
        std::vector<T> g(k);
        for (int it = 0; it < 5; ++it) {
            cudaMemset(grads, 0, k * sizeof(T));
            cudaEventRecord(a);
            addgrad_kernel<T><<<static_cast<unsigned>((n + 255) / 256), 256>>>(values, grads, n, k, pattern, ids);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            if (ms < best) best = ms;
        }
        cudaMemcpy(g.data(), grads, k * sizeof(T), cudaMemcpyDeviceToHost);
        printf("%-6s %-40s %9.3f ms   grads[0..2] = %.6g %.6g %.6g\n", tname, names[pattern], best, double(g[0]), double(g[1]), double(g[2]));
    }
}

int main() {
#ifdef XYZ_AUTODIFF_PLAIN_ATOMICS
    printf("one atomic per thread (XYZ_AUTODIFF_PLAIN_ATOMICS, the reference's behaviour)\n");
#else
    printf("warp-aggregated add_grad (default)\n");
#endif
    run<float>("float");
    run<double>("double");
    return 0;
}
