// dev/pipe_lab.cu -- micro-benchmarks: FFMA vs FFMA2 (fma.rn.f32x2) issue rate, MUFU.EX2 rate, LDS.128 broadcast.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;
constexpr int ILP = 8;

__global__ void k_ffma(float* out, float a, float b) {
    float x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = fmaf(x[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__global__ void k_ffma2(float* out, float a, float b) {
    unsigned long long x[ILP];
    float2 av = make_float2(a, a), bv = make_float2(b, b);
    unsigned long long aa = *reinterpret_cast<unsigned long long*>(&av), bb = *reinterpret_cast<unsigned long long*>(&bv);
#pragma unroll
    for (int i = 0; i < ILP; ++i) { float2 t = make_float2(threadIdx.x * 1e-3f + i, i); x[i] = *reinterpret_cast<unsigned long long*>(&t); }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = ffma2(x[i], aa, bb);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { float2 t = *reinterpret_cast<float2*>(&x[i]); s += t.x + t.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: FFMA2 + LOP3 interleaved (does FFMA2 leave issue slots for the ALU pipe?)
__global__ void k_mix2(float* out, float a, float b) {
    unsigned long long x[ILP];
    unsigned int y[ILP];
    float2 av = make_float2(a, a), bv = make_float2(b, b);
    unsigned long long aa = *reinterpret_cast<unsigned long long*>(&av), bb = *reinterpret_cast<unsigned long long*>(&bv);
#pragma unroll
    for (int i = 0; i < ILP; ++i) { float2 t = make_float2(threadIdx.x * 1e-3f + i, i); x[i] = *reinterpret_cast<unsigned long long*>(&t); y[i] = threadIdx.x + i; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) { x[i] = ffma2(x[i], aa, bb); asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(y[i]) : "r"(it), "r"(0x80000000u)); }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { float2 t = *reinterpret_cast<float2*>(&x[i]); s += t.x + t.y + y[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mix1(float* out, float a, float b) {
    float x[ILP];
    unsigned int y[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = threadIdx.x * 1e-3f + i; y[i] = threadIdx.x + i; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) { x[i] = fmaf(x[i], a, b); asm volatile("lop3.b32 %0, %0, %1, %2, 0x78;" : "+r"(y[i]) : "r"(it), "r"(0x80000000u)); }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// NP packed FFMA2 chains + NS scalar FFMA chains per thread: do the scalar ones run on the "lite" FMA pipe beside the packed ones?
template <int NP, int NS>
__global__ void k_mixps(float* out, float a, float b) {
    unsigned long long x[NP > 0 ? NP : 1];
    float y[NS > 0 ? NS : 1];
    float2 av = make_float2(a, a), bv = make_float2(b, b);
    unsigned long long aa = *reinterpret_cast<unsigned long long*>(&av), bb = *reinterpret_cast<unsigned long long*>(&bv);
#pragma unroll
    for (int i = 0; i < NP; ++i) { float2 t = make_float2(threadIdx.x * 1e-3f + i, i); x[i] = *reinterpret_cast<unsigned long long*>(&t); }
#pragma unroll
    for (int i = 0; i < NS; ++i) y[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < (NP > NS ? NP : NS); ++i) {
            if (i < NP) x[i] = ffma2(x[i], aa, bb);
            if (i < NS) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(y[i]) : "f"(a), "f"(b));
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NP; ++i) { float2 t = *reinterpret_cast<float2*>(&x[i]); s += t.x + t.y; }
#pragma unroll
    for (int i = 0; i < NS; ++i) s += y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_mufu(float* out, float a) {
    float x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// MUFU + FFMA interleaved 1:4
__global__ void k_mufu_ffma(float* out, float a, float b) {
    float x[ILP], y[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = threadIdx.x * 1e-3f + i; y[i] = i; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if ((i & 3) == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            y[i] = fmaf(y[i], a, b);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i] + y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int W>
__global__ void k_lds(float* out) {
    __shared__ float4 sm[256];
    sm[threadIdx.x & 255] = make_float4(threadIdx.x, 1, 2, 3);
    __syncthreads();
    float s = 0;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            const int j = (it + i) & 255;
            if (W == 4) { float4 v = sm[j]; s += v.x + v.w; }
            if (W == 2) { float2 v = reinterpret_cast<float2*>(sm)[j]; s += v.x + v.y; }
            if (W == 1) { float v = reinterpret_cast<float*>(sm)[j]; s += v; }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static float time_it(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const int threads = 256, blocks = sms * 8;   // 64 warps per SM
    float* out; CK(cudaMalloc(&out, sizeof(float) * threads * blocks));
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double warp_instr = double(blocks) * (threads / 32) * ITERS * ILP;  // per kernel
    auto report = [&](const char* name, float ms, double instr_per_iter_elem) {
        const double wi = warp_instr * instr_per_iter_elem;
        printf("%-28s %8.3f ms  %6.2f warp-instr/clk/SM (at %d MHz)\n", name, ms, wi / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
    };
    report("FFMA", time_it([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 1);
    report("FFMA2", time_it([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 1);
    report("FFMA + LOP3 (2 instr)", time_it([&] { k_mix1<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 2);
    report("FFMA2 + LOP3 (2 instr)", time_it([&] { k_mix2<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 2);
    report("MUFU.EX2", time_it([&] { k_mufu<<<blocks, threads>>>(out, 1.0f); }), 1);
    report("MUFU:FFMA 1:4 (1.25 instr)", time_it([&] { k_mufu_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 1.25);
    report("LDS.32 broadcast", time_it([&] { k_lds<1><<<blocks, threads>>>(out); }), 1);
    report("LDS.64 broadcast", time_it([&] { k_lds<2><<<blocks, threads>>>(out); }), 1);
    report("LDS.128 broadcast", time_it([&] { k_lds<4><<<blocks, threads>>>(out); }), 1);
    {
        auto fm = [&](const char* name, float ms, int np, int ns) {
            const double groups = double(blocks) * (threads / 32) * ITERS;
            const double fma_lane_ops = groups * (2.0 * np + ns);
            printf("%-28s %8.3f ms  %6.2f FMA-warp-equivalents/clk/SM  (%d packed + %d scalar per group)\n", name, ms,
                   fma_lane_ops / (ms * 1e-3) / sms / (clk * 1e3), np, ns);
        };
        fm("mix 8 FFMA2 + 0 FFMA", time_it([&] { k_mixps<8, 0><<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 8, 0);
        fm("mix 0 FFMA2 + 8 FFMA", time_it([&] { k_mixps<0, 8><<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 0, 8);
        fm("mix 8 FFMA2 + 4 FFMA", time_it([&] { k_mixps<8, 4><<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 8, 4);
        fm("mix 8 FFMA2 + 8 FFMA", time_it([&] { k_mixps<8, 8><<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 8, 8);
        fm("mix 4 FFMA2 + 8 FFMA", time_it([&] { k_mixps<4, 8><<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 4, 8);
        fm("mix 6 FFMA2 + 6 FFMA", time_it([&] { k_mixps<6, 6><<<blocks, threads>>>(out, 1.0001f, 0.5f); }), 6, 6);
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
