// csrc/splat_kernels.cuh -- forward rasterise and backward kernels of the splat pipeline.
// Included by splat_fast.cu (nvcc -use_fast_math: ex2.approx.ftz / rcp.approx / sin.approx, the
// reference training app's flags, examples/mini-gaussian-splatting/CMakeLists.txt:22-31) and by
// splat_precise.cu (IEEE expf/sinf/cosf/div, the reference's test builds); XYZ_SPLAT_FLAVOR names
// the exported launchers.  Formula order per pair follows gaussian_splatting_kernel.cu:44-61 /
// :84-110 and the Logic structs cited inline.
#pragma once

#include "splat_common.cuh"

#ifndef XYZ_SPLAT_FLAVOR
#error "define XYZ_SPLAT_FLAVOR (fast|precise) before including splat_kernels.cuh"
#endif
#define XYZ_CAT2(a, b) a##b
#define XYZ_CAT(a, b) XYZ_CAT2(a, b)

namespace xyzb {
namespace {

struct PairRec {  // per Gaussian, staged in shared memory for the forward pass
    float4 a;     // cx, cy, ia, 2*ib
    float4 b;     // ic, sigmoid(opacity), r, g
    float c;      // b
};

// weight of one pair: mahalanobis_distance.cuh:88-97, then *0.5f, neg, exp, * sigmoid(opacity)
__device__ __forceinline__ float pair_weight(float px, float py, float cx, float cy, float ia, float ib2, float ic,
                                             float so, float& dx, float& dy, float& e) {
    dx = px - cx;
    dy = py - cy;
    const float d2 = ia * dx * dx + ib2 * dx * dy + ic * dy * dy;
    e = expf(-(d2 * 0.5f));
    return e * so;
}

// ---- forward: one CTA per tile, one pixel per thread -------------------------------------------
__global__ void __launch_bounds__(kTilePixels)
    splat_forward_kernel(SplatView v, const float4* __restrict__ records, const int* __restrict__ sorted_gid,
                         const int2* __restrict__ tile_ranges, const float* __restrict__ target,
                         float* __restrict__ output, float* __restrict__ tile_loss, int tile_y0) {
    __shared__ float4 s_a[kTilePixels];
    __shared__ float4 s_b[kTilePixels];
    __shared__ float s_c[kTilePixels];
    __shared__ float s_red[kTilePixels / 32];

    const int tid = threadIdx.x;
    const int tile_x = blockIdx.x, tile_y = tile_y0 + blockIdx.y;
    const int tile = tile_y * v.tiles_x + tile_x;
    const int pxi = tile_x * kTile + (tid & (kTile - 1));
    const int pyi = tile_y * kTile + (tid >> 4);
    const bool active = pxi < v.width && pyi >= v.row_begin && pyi < v.row_end;
    const float px = static_cast<float>(pxi), py = static_cast<float>(pyi);

    const int2 range = tile_ranges[tile];
    float o0 = 0.f, o1 = 0.f, o2 = 0.f;
    for (int base = range.x; base < range.y; base += kTilePixels) {
        const int n = min(kTilePixels, range.y - base);
        __syncthreads();
        if (tid < n) {
            const int g = sorted_gid[base + tid];
            const float4 r0 = __ldg(records + 3 * g), r1 = __ldg(records + 3 * g + 1), r2 = __ldg(records + 3 * g + 2);
            s_a[tid] = make_float4(r0.x, r0.y, r0.z, 2.0f * r0.w);
            s_b[tid] = make_float4(r1.x, r1.y, r1.z, r1.w);
            s_c[tid] = r2.x;
        }
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            const float4 a = s_a[j];
            const float4 b = s_b[j];
            const float cb = s_c[j];
            float dx, dy, e;
            const float w = pair_weight(px, py, a.x, a.y, a.z, a.w, b.x, b.y, dx, dy, e);
            o0 += b.z * w;  // color * broadcast(weighted_gauss), ascending Gaussian index
            o1 += b.w * w;
            o2 += cb * w;
        }
    }
    float l = 0.f;
    if (active) {
        const size_t p = static_cast<size_t>(pyi) * v.width + pxi;
        output[3 * p] = o0;
        output[3 * p + 1] = o1;
        output[3 * p + 2] = o2;
        // gaussian_splatting_kernel.cu:68-70
        l = fabsf(o0 - __ldg(target + 3 * p)) + fabsf(o1 - __ldg(target + 3 * p + 1)) +
            fabsf(o2 - __ldg(target + 3 * p + 2));
    }
    l = warp_sum(l);
    if ((tid & 31) == 0) s_red[tid >> 5] = l;
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kTilePixels / 32; ++w) s += s_red[w];
        tile_loss[tile] = s;
    }
}

// ---- backward: one thread per (tile, Gaussian) list entry ----------------------------------------
// Per pair (gaussian_splatting_kernel.cu:84-110, run() of the l1_norm root):
//   cd_i = wc_i - (tgt_i - out_i)          (rest_sum += un-forwarded weighted_color adds 0: Q2)
//   s_i  = sign(cd_i)                      l1_norm_logic.cuh:29-37
//   d color_i += s_i * w ; g_w = sum s_i * color_i          mul_logic.cuh:33-41
//   g_e = g_w * so ; d so += g_w * e
//   g_d2 = -(g_e * e) * 0.5                exp / neg / mul_constant backward
//   d center -= g_d2 * (2 ia dx + 2 ib dy , 2 ib dx + 2 ic dy)     mahalanobis_distance.cuh:100-113
//   d inv += g_d2 * (dx^2, 2 dx dy, dy^2)                           :116-118
// Everything after this is linear with per-Gaussian coefficients and is applied once per entry.
__global__ void __launch_bounds__(kTilePixels)
    splat_backward_kernel(SplatView v, const float4* __restrict__ records, const unsigned int* __restrict__ keys_sorted,
                          const int* __restrict__ sorted_gid, const unsigned int* __restrict__ sorted_orig,
                          const xyz_gaussian_params* __restrict__ params, xyz_gaussian_grads* grads,
                          const float* __restrict__ target, const float* __restrict__ output, long long entries,
                          float* __restrict__ entry_grads) {
    __shared__ float4 s_rest[kTilePixels];  // tgt - out per pixel of the current tile; w < 0: pixel inactive
    __shared__ int s_tiles[kTilePixels];
    __shared__ int s_ntiles;

    const int tid = threadIdx.x;
    const long long i = blockIdx.x * static_cast<long long>(kTilePixels) + tid;
    const bool valid = i < entries;
    const int my_tile = valid ? static_cast<int>(keys_sorted[i]) : -1;
    if (tid == 0) s_ntiles = 0;
    __syncthreads();
    {
        const int prev = (tid > 0 && valid) ? static_cast<int>(keys_sorted[i - 1]) : -2;
        if (valid && (tid == 0 || prev != my_tile)) s_tiles[atomicAdd(&s_ntiles, 1)] = my_tile;
    }

    float cx = 0.f, cy = 0.f, ia = 0.f, ib = 0.f, ic = 0.f, so = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
    int g = 0;
    if (valid) {
        g = sorted_gid[i];
        const float4 r0 = __ldg(records + 3 * g), r1 = __ldg(records + 3 * g + 1), r2 = __ldg(records + 3 * g + 2);
        cx = r0.x; cy = r0.y; ia = r0.z; ib = r0.w;
        ic = r1.x; so = r1.y; c0 = r1.z; c1 = r1.w; c2 = r2.x;
    }
    const float ib2 = 2.0f * ib;
    float a_c0 = 0.f, a_c1 = 0.f, a_c2 = 0.f, a_so = 0.f, a_cx = 0.f, a_cy = 0.f, a_ia = 0.f, a_ib = 0.f, a_ic = 0.f;
    __syncthreads();
    const int ntiles = s_ntiles;

    for (int k = 0; k < ntiles; ++k) {
        const int tile = s_tiles[k];
        const int tile_x = tile % v.tiles_x, tile_y = tile / v.tiles_x;
        __syncthreads();
        {
            const int pxi = tile_x * kTile + (tid & (kTile - 1));
            const int pyi = tile_y * kTile + (tid >> 4);
            float4 r = make_float4(0.f, 0.f, 0.f, -1.f);
            if (pxi < v.width && pyi >= v.row_begin && pyi < v.row_end) {
                const size_t p = static_cast<size_t>(pyi) * v.width + pxi;
                r.x = __ldg(target + 3 * p) - __ldg(output + 3 * p);  // rest_sum = target_color - pixel_out
                r.y = __ldg(target + 3 * p + 1) - __ldg(output + 3 * p + 1);
                r.z = __ldg(target + 3 * p + 2) - __ldg(output + 3 * p + 2);
                r.w = 1.f;
            }
            s_rest[tid] = r;
        }
        __syncthreads();
        if (my_tile != tile) continue;
        const float px0 = static_cast<float>(tile_x * kTile), py0 = static_cast<float>(tile_y * kTile);
#pragma unroll 2
        for (int p = 0; p < kTilePixels; ++p) {
            const float4 rest = s_rest[p];
            if (rest.w < 0.f) continue;  // uniform across the CTA
            const float px = px0 + static_cast<float>(p & (kTile - 1));
            const float py = py0 + static_cast<float>(p >> 4);
            float dx, dy, e;
            const float w = pair_weight(px, py, cx, cy, ia, ib2, ic, so, dx, dy, e);
            const float cd0 = c0 * w - rest.x, cd1 = c1 * w - rest.y, cd2 = c2 * w - rest.z;
            const float s0 = cd0 > 0.f ? 1.f : (cd0 < 0.f ? -1.f : 0.f);
            const float s1 = cd1 > 0.f ? 1.f : (cd1 < 0.f ? -1.f : 0.f);
            const float s2 = cd2 > 0.f ? 1.f : (cd2 < 0.f ? -1.f : 0.f);
            a_c0 += s0 * w;
            a_c1 += s1 * w;
            a_c2 += s2 * w;
            const float g_w = s0 * c0 + s1 * c1 + s2 * c2;
            a_so += g_w * e;
            const float g_d2 = -((g_w * so) * e) * 0.5f;
            a_cx -= g_d2 * (2.0f * ia * dx + ib2 * dy);
            a_cy -= g_d2 * (ib2 * dx + 2.0f * ic * dy);
            a_ia += g_d2 * dx * dx;
            a_ib += g_d2 * 2.0f * dx * dy;
            a_ic += g_d2 * dy * dy;
        }
    }
    if (!valid) return;

    // Per-Gaussian chain rule, once per entry (linear in the sums above).
    const xyz_gaussian_params gp = params[g];
    // sym_matrix2_inv backward (sym_matrix2_inv_logic.cuh:43-77) needs Sigma = (A, B, C):
    const float es0 = expf(gp.scale[0]), es1 = expf(gp.scale[1]);  // exp_logic.cuh:17-36
    const float ct = cosf(gp.rotation[0]), sn = sinf(gp.rotation[0]);
    const float m00 = es0 * ct, m01 = -es1 * sn, m10 = es0 * sn, m11 = es1 * ct;  // covariance_generation.cuh:154-172
    const float A = m00 * m00 + m01 * m01, B = m00 * m10 + m01 * m11, C = m10 * m10 + m11 * m11;
    float det = A * C - B * B;
    if (fabsf(det) < 1e-8f) det = 1e-8f;
    const float inv_det = 1.0f / det;
    const float inv_det2 = inv_det * inv_det;
    const float g_A = a_ia * (-C * C * inv_det2) + a_ib * (B * C * inv_det2) + a_ic * (inv_det - A * C * inv_det2);
    const float g_B = a_ia * (2.0f * C * B * inv_det2) + a_ib * (-inv_det - 2.0f * B * B * inv_det2) +
                      a_ic * (2.0f * A * B * inv_det2);
    const float g_C = a_ia * (inv_det - A * C * inv_det2) + a_ib * (A * B * inv_det2) + a_ic * (-A * A * inv_det2);
    // scale_rotation_to_covariance_3param backward (covariance_generation.cuh:175-211)
    const float g00 = g_A * 2.0f * m00 + g_B * m10;
    const float g01 = g_A * 2.0f * m01 + g_B * m11;
    const float g10 = g_B * m00 + g_C * 2.0f * m10;
    const float g11 = g_B * m01 + g_C * 2.0f * m11;
    const float g_es0 = g00 * ct + g10 * sn;
    const float g_es1 = g01 * (-sn) + g11 * ct;
    const float g_theta = g00 * (-es0 * sn) + g01 * (-es1 * ct) + g10 * (es0 * ct) + g11 * (-es1 * sn);
    float out9[9];
    out9[0] = a_cx;
    out9[1] = a_cy;
    out9[2] = g_es0 * es0;  // exp backward recomputes exp(scale)
    out9[3] = g_es1 * es1;
    out9[4] = g_theta;
    out9[5] = a_c0;
    out9[6] = a_c1;
    out9[7] = a_c2;
    out9[8] = a_so * (so * (1.0f - so));  // sigmoid_logic.cuh:27-36
    if (entry_grads) {
        float* row = entry_grads + static_cast<size_t>(sorted_orig[i]) * 9;
#pragma unroll
        for (int k = 0; k < 9; ++k) row[k] = out9[k];
    } else {
        float* gg = reinterpret_cast<float*>(grads + g);
#pragma unroll
        for (int k = 0; k < 9; ++k) atomicAdd(gg + k, out9[k]);  // VariableRef::add_grad, variable.cuh:48-50
    }
}

}  // namespace

int XYZ_CAT(splat_forward_launch_, XYZ_SPLAT_FLAVOR)(const SplatView& v, const SplatBuffers& b, const float* target,
                                                     float* output, cudaStream_t st) {
    const int ty0 = v.row_begin / kTile, ty1 = (v.row_end + kTile - 1) / kTile;
    if (ty1 <= ty0) return 0;
    dim3 grid(v.tiles_x, ty1 - ty0);
    splat_forward_kernel<<<grid, kTilePixels, 0, st>>>(v, b.records, b.sorted_gid, b.tile_ranges, target, output,
                                                       b.tile_loss, ty0);
    count_launch();
    return last_error();
}

int XYZ_CAT(splat_backward_launch_, XYZ_SPLAT_FLAVOR)(const SplatView& v, const SplatBuffers& b,
                                                      const xyz_gaussian_params* params, xyz_gaussian_grads* grads,
                                                      const float* target, const float* output, long long entries,
                                                      bool deterministic, cudaStream_t st) {
    if (entries <= 0) return 0;
    const long long blocks = (entries + kTilePixels - 1) / kTilePixels;
    splat_backward_kernel<<<static_cast<unsigned int>(blocks), kTilePixels, 0, st>>>(
        v, b.records, b.keys_out, b.sorted_gid, b.vals_out, params, grads, target, output, entries,
        deterministic ? b.entry_grads : nullptr);
    count_launch();
    return last_error();
}

}  // namespace xyzb
