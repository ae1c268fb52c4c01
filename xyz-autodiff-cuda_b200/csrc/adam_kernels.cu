// csrc/adam_kernels.cu -- the two streaming passes either side of the splat kernel:
// zero_gradients_kernel (reference examples/mini-gaussian-splatting/gaussian_parameters.cu:227-241)
// and adam_step_individual_kernel / adam_step_kernel (:260-320 / :173-224; host wrappers :322-386).
//
// The reference runs one thread per Gaussian over 36-byte / 72-byte AoS records (stride-9 and
// stride-18 scalar accesses).  Here both passes run one thread per FLOAT over the flat arrays, so
// params/grads are read and written fully coalesced; the Adam moments of a warp's 32 consecutive
// components sit in at most 5 consecutive 72-byte records, every sector of which is consumed.
// 252 algorithmic bytes per Gaussian (read 36+36+72, write 36+72) -> HBM bound.
// Like the reference's GPU kernels (and unlike its host adam_step) nothing is clamped.
#include "common.cuh"

#include <cmath>

namespace xyzb {
namespace {

__global__ void zero_kernel(float4* __restrict__ p4, long long n4, float* __restrict__ tail, int ntail) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i < n4) p4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < ntail) tail[i] = 0.f;
}

struct AdamLr {
    float lr_corrected[5];  // center, scale, rotation, color, opacity
};

// component j of GaussianParams -> (lr group, offset of m inside AdamState, group length)
__device__ __forceinline__ void adam_slot(int j, int& group, int& m_off, int& v_off) {
    // AdamState (gaussian_parameters.h:21-32): m_center[2] v_center[2] m_scale[2] v_scale[2]
    //                                          m_rot v_rot m_color[3] v_color[3] m_op v_op
    if (j < 2) { group = 0; m_off = j; v_off = 2 + j; }
    else if (j < 4) { group = 1; m_off = 4 + (j - 2); v_off = 6 + (j - 2); }
    else if (j == 4) { group = 2; m_off = 8; v_off = 9; }
    else if (j < 8) { group = 3; m_off = 10 + (j - 5); v_off = 13 + (j - 5); }
    else { group = 4; m_off = 16; v_off = 17; }
}

__global__ void __launch_bounds__(256)
    adam_kernel(float* __restrict__ params, const float* grads, float* __restrict__ adam, long long n_comp, AdamLr lr,
                float beta1, float beta2, float eps, float* grads_to_zero /* == grads or nullptr */) {
    const long long f = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (f >= n_comp) return;
    const long long g = f / 9;
    const int j = static_cast<int>(f - g * 9);
    int group, m_off, v_off;
    adam_slot(j, group, m_off, v_off);
    float* st = adam + g * 18;
    const float gr = grads[f];
    const float m = beta1 * st[m_off] + (1.0f - beta1) * gr;
    const float v = beta2 * st[v_off] + (1.0f - beta2) * gr * gr;
    st[m_off] = m;
    st[v_off] = v;
    params[f] -= lr.lr_corrected[group] * m / (sqrtf(v) + eps);
    // fused zero_gradients_kernel (gaussian_parameters.cu:227-241): the gradient has been consumed, the next
    // iteration's kernels accumulate into zeros without a separate pass
    if (grads_to_zero) grads_to_zero[f] = 0.f;
}

int adam_launch(xyz_gaussian_params* params, const xyz_gaussian_grads* grads, xyz_adam_state* adam, int n,
                const float lr[5], float beta1, float beta2, float eps, int iteration, void* stream,
                xyz_gaussian_grads* grads_to_zero = nullptr) {
    if (n < 0 || iteration < 1) return XYZ_ERR_INVALID_ARGUMENT;
    if (n == 0) return 0;
    if (!params || !grads || !adam) return XYZ_ERR_INVALID_ARGUMENT;
    // host wrapper, gaussian_parameters.cu:357-358 + kernel :279-283
    const float b1t = std::pow(beta1, static_cast<float>(iteration));
    const float b2t = std::pow(beta2, static_cast<float>(iteration));
    AdamLr l;
    for (int k = 0; k < 5; ++k) l.lr_corrected[k] = lr[k] * std::sqrt(1.0f - b2t) / (1.0f - b1t);
    const long long n_comp = static_cast<long long>(n) * 9;
    const int grid = static_cast<int>((n_comp + 255) / 256);
    adam_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<float*>(params),
                                                                     reinterpret_cast<const float*>(grads),
                                                                     reinterpret_cast<float*>(adam), n_comp, l, beta1,
                                                                     beta2, eps, reinterpret_cast<float*>(grads_to_zero));
    count_launch();
    return last_error();
}

}  // namespace

// adam_step_individual on the Gaussians [g_begin, g_end) only (sharded optimiser step, csrc/comm.cu)
int adam_launch_range(xyz_gaussian_params* params, const xyz_gaussian_grads* grads, xyz_adam_state* adam, int g_begin,
                      int g_end, const float lr[5], float beta1, float beta2, float eps, int iteration, void* stream) {
    if (g_begin < 0 || g_end < g_begin) return XYZ_ERR_INVALID_ARGUMENT;
    return adam_launch(params + g_begin, grads + g_begin, adam + g_begin, g_end - g_begin, lr, beta1, beta2, eps, iteration,
                       stream);
}

}  // namespace xyzb

extern "C" int xyz_zero_gradients(xyz_gaussian_grads* gradients, int num_gaussians, void* stream) {
    using namespace xyzb;
    if (num_gaussians < 0) return XYZ_ERR_INVALID_ARGUMENT;
    if (num_gaussians == 0) return 0;
    if (!gradients) return XYZ_ERR_INVALID_ARGUMENT;
    float* p = reinterpret_cast<float*>(gradients);
    const long long n = static_cast<long long>(num_gaussians) * 9;
    long long n4 = 0;
    int ntail = static_cast<int>(n);
    float* tail = p;
    if (aligned16(p)) {
        n4 = n / 4;
        ntail = static_cast<int>(n - n4 * 4);
        tail = p + n4 * 4;
    } else if (n > 1 << 20) {  // misaligned and large: a memset is still one coalesced pass
        cudaError_t e = cudaMemsetAsync(p, 0, n * 4, static_cast<cudaStream_t>(stream));
        return static_cast<int>(e);
    }
    const long long work = n4 > ntail ? n4 : ntail;
    const int grid = static_cast<int>((work + 255) / 256);
    zero_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<float4*>(p), n4, tail, ntail);
    count_launch();
    return last_error();
}

extern "C" int xyz_adam_step_individual(xyz_gaussian_params* params, const xyz_gaussian_grads* grads,
                                        xyz_adam_state* adam, int num_gaussians, const float lr_host[5], float beta1,
                                        float beta2, float epsilon, int iteration, void* stream) {
    if (!lr_host) return XYZ_ERR_INVALID_ARGUMENT;
    return xyzb::adam_launch(params, grads, adam, num_gaussians, lr_host, beta1, beta2, epsilon, iteration, stream);
}

extern "C" int xyz_adam_step(xyz_gaussian_params* params, const xyz_gaussian_grads* grads, xyz_adam_state* adam,
                             int num_gaussians, float learning_rate, float beta1, float beta2, float epsilon,
                             int iteration, void* stream) {
    const float lr[5] = {learning_rate, learning_rate, learning_rate, learning_rate, learning_rate};
    return xyzb::adam_launch(params, grads, adam, num_gaussians, lr, beta1, beta2, epsilon, iteration, stream);
}

extern "C" int xyz_adam_step_individual_zero_grads(xyz_gaussian_params* params, xyz_gaussian_grads* grads,
                                                   xyz_adam_state* adam, int num_gaussians, const float lr_host[5],
                                                   float beta1, float beta2, float epsilon, int iteration, void* stream) {
    if (!lr_host) return XYZ_ERR_INVALID_ARGUMENT;
    return xyzb::adam_launch(params, grads, adam, num_gaussians, lr_host, beta1, beta2, epsilon, iteration, stream, grads);
}
