// csrc/splat_kernels.cuh -- forward rasterise and backward kernels of the splat pipeline.
// Included by splat_fast.cu (nvcc -use_fast_math: ex2.approx.ftz / rcp.approx / sin.approx, the
// reference training app's flags, examples/mini-gaussian-splatting/CMakeLists.txt:22-31) and by
// splat_precise.cu (IEEE expf/sinf/cosf/div, the reference's test builds); XYZ_SPLAT_FLAVOR names
// the exported launchers.  Formula order per pair follows gaussian_splatting_kernel.cu:44-61 /
// :84-110 and the Logic structs cited inline.
#pragma once

#include <cstdlib>

#include "splat_common.cuh"

#ifndef XYZ_SPLAT_FLAVOR
#error "define XYZ_SPLAT_FLAVOR (fast|precise) before including splat_kernels.cuh"
#endif
#ifndef XYZ_BWD_MINBLOCKS
#define XYZ_BWD_MINBLOCKS 5  // resident backward CTAs per SM the register budget is sized for (96 registers)
#endif
#ifndef XYZ_BWD_DX2
#define XYZ_BWD_DX2 1  // backward pixel loop keeps dx^2 in registers: 13 instead of 14 FMA-pipe operations per pixel
#endif
#ifndef XYZ_BWD_ROW_UNROLL
#define XYZ_BWD_ROW_UNROLL 1  // rows of the tile per iteration of the backward pixel loop
#endif
#ifndef XYZ_FWD_UNROLL
#define XYZ_FWD_UNROLL 4      // Gaussians per iteration of the forward loop
#endif
#define XYZ_PRAGMA(x) _Pragma(#x)
#define XYZ_UNROLL(n) XYZ_PRAGMA(unroll n)
#define XYZ_CAT2(a, b) a##b
#define XYZ_CAT(a, b) XYZ_CAT2(a, b)

namespace xyzb {
namespace {

// Exponent of one pair.  The reference evaluates exp(-(d2 * 0.5)) (mul_constant / neg / exp nodes); both
// flavours fold the constant into the per-Gaussian conic (A2, B2, C2) = kappa * (ia, 2 ib, ic) so the pair costs
// two FMAs and one special-function op:
//   fast    kappa = -0.5 * log2(e), e = ex2.approx.ftz(arg)   (what -use_fast_math makes of expf: ex2(x * log2 e))
//   precise kappa = -0.5,           e = expf(arg)             (IEEE flavour)
// Both underflow to exactly 0.0f beyond the d2 bounds in splat_common.cuh (checked by the cull parity test).
#if XYZ_SPLAT_IS_FAST
constexpr float kKappa = kKappaFast;
__device__ __forceinline__ float pair_exp(float arg) { return exp2f(arg); }
#else
constexpr float kKappa = kKappaPrecise;
__device__ __forceinline__ float pair_exp(float arg) { return expf(arg); }
#endif

// v with the sign of `from` XOR-ed in: one LOP3 (v ^ (from & 0x80000000)), lut 0x78
__device__ __forceinline__ float xor_sign(float v, float from) {
    unsigned int r;
    asm("lop3.b32 %0, %1, %2, 0x80000000, 0x78;" : "=r"(r) : "r"(__float_as_uint(v)), "r"(__float_as_uint(from)));
    return __uint_as_float(r);
}

// Packed fp32 (Blackwell FFMA2 / FADD2 / FMUL2 = PTX fma/add/mul.rn.f32x2): two independent IEEE operations per instruction
// on the halves of 64-bit registers; the same bits as two scalar operations.
struct F2 {
    unsigned long long v;
};
#if XYZ_SPLAT_IS_FAST
#define XYZ_F2_FTZ ".ftz"
#else
#define XYZ_F2_FTZ ""
#endif
__device__ __forceinline__ F2 f2_pack(float lo, float hi) {
    F2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(F2 a, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v));
}
__device__ __forceinline__ F2 f2_fma(F2 a, F2 b, F2 c) {
    F2 r;
    asm("fma.rn" XYZ_F2_FTZ ".f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
__device__ __forceinline__ F2 f2_mul(F2 a, F2 b) {
    F2 r;
    asm("mul.rn" XYZ_F2_FTZ ".f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ F2 f2_add(F2 a, F2 b) {
    F2 r;
    asm("add.rn" XYZ_F2_FTZ ".f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ float f2_hsum(F2 a) {
    float lo, hi;
    f2_unpack(a, lo, hi);
    return lo + hi;
}

// ---- forward: one CTA (64 threads) per tile, four pixels per thread -------------------------------------
// Staged per Gaussian (shared memory, 32 B): {cx, cy, A2, B2} {C2, so*r, so*g, so*b}.  ncu on the one-pixel-per-
// thread version: 87 % of the shared-memory pipe (two broadcast LDS.128 = 4 wavefronts per pair per warp), so a
// thread now owns the pixels (x, y), (x, y+4), (x, y+8), (x, y+12) and reads each Gaussian once for all four:
//   dx = x - cx ; t0 = (A2 dx) dx ; bdx = B2 dx                          shared by the four pixels
//   dy = y_k - cy ; arg = dy (C2 dy + bdx) + t0 ; e = exp(arg) ; out_k,i += (so c_i) e
// = 1 shared-memory wavefront, 1 MUFU and ~8.5 issue slots per pair.  Summation runs in ascending Gaussian
// index per pixel, the reference's order (gaussian_splatting_kernel.cu:33-62).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Staging is a three-deep software pipeline per CTA, so that a CTA never waits for global memory even when it is alone
// on its SM sub-partition: while stage s is being rendered, the 32-byte forward records of stage s + 1 are in flight
// (cp.async into the other shared-memory buffer) and the Gaussian ids of stage s + 2 are in flight into registers.
//
// kParts = CTAs per tile.  1: one CTA renders the 16 x 16 tile (kThreads = 64: four pixels per thread, the throughput
// configuration).  2: two CTAs of 64 threads render the upper and the lower 16 x 8 half (two pixels per thread) -- for
// launches with FEW tiles: a row band of an image sharded over 8 GPUs has 512 tiles = 3.46 per SM, all resident at once,
// so SMs with 4 tiles run 33 % longer than SMs with 3 and the kernel ends with a fifth of the machine idle (ncu,
// profiles/ncu_r02_band_kernels.csv); with 1024 half tiles the quantum halves.  Per pixel the arithmetic and the
// summation order are the same in every configuration: the image is bit-identical (tested).
// The loss partial is kept per HALF tile (tile_loss[2 tile + half], each a fixed tree over its 128 pixels), so the sum
// the last CTA forms over the launch's half tiles has the same bits in every configuration too.
// The CTAs run longest list first (tile_order, from the tile scan), and the CTA of a tile (its upper half with two
// CTAs per tile) also lists the backward work items of the tile: see "backward work items" below.

// Smallest value of the conic form q(dx, dy) = ia dx^2 + 2 ib dx dy + ic dy^2 (positive definite) over the pixel
// rectangle [x0, x1] x [y0, y1] given relative to the Gaussian's centre: 0 when the centre lies inside, else the smallest
// of the four edge minima (q is convex; on an edge it is a parabola whose vertex is clamped to the edge).  A lower bound
// of q over the rectangle's pixels (tests/test_oracle.py restates it against a dense sampling).
__device__ __forceinline__ float conic_min_over_rect(float ia, float ib, float ic, float kx, float ky, float x0, float x1,
                                                     float y0, float y1) {
    if (x0 <= 0.f && x1 >= 0.f && y0 <= 0.f && y1 >= 0.f) return 0.f;
    const float ib2 = 2.0f * ib;
    auto q = [&](float dx, float dy) { return fmaf(dx, fmaf(ia, dx, ib2 * dy), (ic * dy) * dy); };
    float m = q(fminf(fmaxf(kx * y0, x0), x1), y0);
    m = fminf(m, q(fminf(fmaxf(kx * y1, x0), x1), y1));
    m = fminf(m, q(x0, fminf(fmaxf(ky * x0, y0), y1)));
    m = fminf(m, q(x1, fminf(fmaxf(ky * x1, y0), y1)));
    return m;
}

constexpr int kFwdStage = 128;  // Gaussians staged per pass
constexpr int kHalfPixels = kTilePixels / 2;

#ifndef XYZ_FWD_MINBLOCKS
#define XYZ_FWD_MINBLOCKS 20  // resident whole-tile forward CTAs per SM the register budget is sized for: 48 registers, no spills
#endif                        // (56 unconstrained = 18 CTAs: 373 us at C4 against 367; 24 CTAs = 40 registers spill: 401)
template <int kThreads, int kParts>
__global__ void __launch_bounds__(kThreads, kThreads == 64 ? XYZ_FWD_MINBLOCKS : 1)
    splat_forward_kernel(SplatView v, const float4* __restrict__ fwd_records, const int* __restrict__ sorted_gid,
                         const int2* __restrict__ tile_ranges, const float* __restrict__ target,
                         float* __restrict__ output, float* __restrict__ tile_loss, float4* __restrict__ rest_tiles,
                         int tile_y0, unsigned int* __restrict__ ticket, float* total_loss, int* __restrict__ bwd_items,
                         const int* __restrict__ chunk_offsets, int4* __restrict__ chunk_info, int first_tile,
                         float d2_bwd_scaled, const unsigned int* __restrict__ sorted_orig, float* __restrict__ entry_grads,
                         const int* __restrict__ tile_order) {
    static_assert(kParts == 1 || (kParts == 2 && kThreads == 64), "half tiles are rendered by 64 threads");
    constexpr int kPixels = kTilePixels / kParts;   // pixels of this CTA
    constexpr int kRows = kPixels / kThreads;       // pixel rows per thread (4, 2, 1)
    constexpr int kRowStep = kThreads / kTile;      // distance between a thread's rows
    constexpr int kPerThread = (kFwdStage + kThreads - 1) / kThreads;  // Gaussians a thread stages per pass
    __shared__ __align__(16) float4 s_a[2][kFwdStage];
    __shared__ __align__(16) float4 s_b[2][kFwdStage];
    __shared__ int s_items;  // length of the tile's list of backward work items

    const int tid = threadIdx.x;
    // grid: (tiles of the band) x kParts CTAs, a tile's parts next to each other; tile_order = the tiles by falling list length
    const int slot = static_cast<int>(blockIdx.x) / kParts;
    const int band_tile = tile_order ? __ldg(tile_order + slot) : slot;
    const int tile_x = band_tile % v.tiles_x, tile_y = tile_y0 + band_tile / v.tiles_x;
    const int part = static_cast<int>(blockIdx.x) % kParts;  // 0: rows 0..7 (or the whole tile), 1: rows 8..15
    const int tile = tile_y * v.tiles_x + tile_x;
    const int pxi = tile_x * kTile + (tid & (kTile - 1));
    const int row0 = part * (kTile / kParts) + (tid >> 4);  // this thread's rows inside the tile: row0 + kRowStep k
    const int pyi0 = tile_y * kTile + row0;
    const float px = static_cast<float>(pxi);
    float py[kRows];
#pragma unroll
    for (int k = 0; k < kRows; ++k) py[k] = static_cast<float>(pyi0 + kRowStep * k);

    const int2 range = tile_ranges[tile];
    float o[kRows][3];
#pragma unroll
    for (int k = 0; k < kRows; ++k) o[k][0] = o[k][1] = o[k][2] = 0.f;

    int gid[kPerThread];  // ids of the stage that is fetched next
    auto load_ids = [&](int base) {
#pragma unroll
        for (int q = 0; q < kPerThread; ++q) {
            const int t = tid + q * kThreads;
            gid[q] = (t < kFwdStage && base + t < range.y) ? __ldg(sorted_gid + base + t) : -1;
        }
    };
    auto fetch = [&](int buf) {  // records of the ids in gid[] -> shared-memory buffer `buf` (asynchronous)
#pragma unroll
        for (int q = 0; q < kPerThread; ++q) {
            const int t = tid + q * kThreads;
            if (gid[q] >= 0) {
                cp_async16(&s_a[buf][t], fwd_records + 2 * gid[q]);
                cp_async16(&s_b[buf][t], fwd_records + 2 * gid[q] + 1);
            }
        }
        cp_async_commit();
    };
    // backward work items (see "backward work items" below); with two CTAs per tile the upper one lists them
    const bool lists_items = kParts == 1 || part == 0;
    const bool by_entry = entry_grads != nullptr;  // deterministic mode: an item is the entry's index (one row per entry)
    const float tx0 = static_cast<float>(tile_x * kTile), ty0 = static_cast<float>(tile_y * kTile);
    int* const items0 = bwd_items + range.x;  // the tile's items take the first slots of its range
    const int list_len = range.y - range.x;
    if (tid == 0) s_items = 0;  // (the barriers at the top of the loop / after it order this with every use)

    load_ids(range.x);
    int id_cur[kPerThread];  // ids of the stage being rendered
#pragma unroll
    for (int q = 0; q < kPerThread; ++q) id_cur[q] = gid[q];
    fetch(0);
    load_ids(range.x + kFwdStage);
    int buf = 0;
    for (int base = range.x; base < range.y; base += kFwdStage, buf ^= 1) {
        const int n = min(kFwdStage, range.y - base);
        // every thread is done READING buffer buf ^ 1 (the previous stage) before anybody refills it
        __syncthreads();
        int id_nxt[kPerThread];
#pragma unroll
        for (int q = 0; q < kPerThread; ++q) id_nxt[q] = gid[q];
        fetch(buf ^ 1);                        // stage s + 1 (an empty group past the end)
        load_ids(base + 2 * kFwdStage);        // ids of stage s + 2
        cp_async_wait<1>();                    // this thread's copies of stage s have landed ...
        __syncthreads();                       // ... and everybody else's
        const float4* __restrict__ sa = s_a[buf];
        const float4* __restrict__ sb = s_b[buf];
        // the stage's backward work items: min d2 over the tile against the backward cull's bound, on the staged
        // (kappa-scaled, hence negated) conic
        if (lists_items) {
#pragma unroll
            for (int q = 0; q < kPerThread; ++q) {
                const int t = tid + q * kThreads;
                bool keep = false;
                if (t < n) {
                    const float4 a = sa[t];
                    const float ia = -a.z, ib = -0.5f * a.w, ic = -sb[t].x;
                    // anything but a positive definite conic of finite numbers is kept (the comparisons fail on NaN)
                    const bool pd = ia > 0.f && ic > 0.f && ia * ic > ib * ib;
                    const float x0 = tx0 - a.x, y0 = ty0 - a.y;
                    keep = !(pd && conic_min_over_rect(ia, ib, ic, -ib / ia, -ib / ic, x0, x0 + static_cast<float>(kTile - 1), y0,
                                                       y0 + static_cast<float>(kTile - 1)) > d2_bwd_scaled);
                    if (by_entry && !keep) {  // the per-Gaussian sum reads every entry's row
                        float* row = entry_grads + static_cast<size_t>(sorted_orig[base + t]) * 9;
#pragma unroll
                        for (int k = 0; k < 9; ++k) row[k] = 0.f;
                    }
                }
                const unsigned int votes = __ballot_sync(0xffffffffu, keep);
                if (votes == 0u) continue;
                const int lane = tid & 31;
                int at = 0;
                if (lane == 0) at = atomicAdd(&s_items, __popc(votes));
                at = __shfl_sync(0xffffffffu, at, 0) + __popc(votes & ((1u << lane) - 1u));
                if (keep) items0[at] = by_entry ? base + t : id_cur[q];
            }
        }
        XYZ_UNROLL(XYZ_FWD_UNROLL)
        for (int j = 0; j < n; ++j) {
            const float4 a = sa[j];
            const float4 b = sb[j];
            const float dx = px - a.x;
            const float t0 = (a.z * dx) * dx;
            const float bdx = a.w * dx;
#pragma unroll
            for (int k = 0; k < kRows; ++k) {
                const float dy = py[k] - a.y;
                const float e = pair_exp(fmaf(dy, fmaf(b.x, dy, bdx), t0));
                o[k][0] = fmaf(b.y, e, o[k][0]);
                o[k][1] = fmaf(b.z, e, o[k][1]);
                o[k][2] = fmaf(b.w, e, o[k][2]);
            }
        }
#pragma unroll
        for (int q = 0; q < kPerThread; ++q) id_cur[q] = id_nxt[q];
    }
    cp_async_wait<0>();
    __shared__ float s_l[kTilePixels];  // per-pixel |out - target| by pixel index inside the tile (row-major)
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
        const int row = row0 + kRowStep * k, pyi = pyi0 + kRowStep * k;
        const int pix = row * kTile + (tid & (kTile - 1));
        // rest_sum = target_color - pixel_out (gaussian_splatting_kernel.cu:99-101) for the backward pass, stored
        // tile-major (4 KB contiguous per tile); .w = 1 for pixels of this launch, 0 outside the image / row band
        float4 rest = make_float4(0.f, 0.f, 0.f, 0.f);
        float l = 0.f;
        if (pxi < v.width && pyi >= v.row_begin && pyi < v.row_end) {
            const size_t p = static_cast<size_t>(pyi) * v.width + pxi;
            output[3 * p] = o[k][0];
            output[3 * p + 1] = o[k][1];
            output[3 * p + 2] = o[k][2];
            rest.x = __ldg(target + 3 * p) - o[k][0];
            rest.y = __ldg(target + 3 * p + 1) - o[k][1];
            rest.z = __ldg(target + 3 * p + 2) - o[k][2];
            rest.w = 1.f;
            // gaussian_splatting_kernel.cu:68-70: |out - target|
            l = fabsf(rest.x) + fabsf(rest.y) + fabsf(rest.z);
        }
        rest_tiles[static_cast<size_t>(tile) * kTilePixels + pix] = rest;
        s_l[pix] = l;
    }
    __syncthreads();
    if (lists_items && list_len > 0) {  // (no entries: no records set aside, and with no Gaussians at all no offsets either)
        // the tile's backward work records: one per kBwdChunk items, in the ceil(len / kBwdChunk) records the tile scan
        // set aside for this tile; the rest of them: none
        const int lt = tile - first_tile;
        const int rec0 = chunk_offsets[lt], rec1 = chunk_offsets[lt + 1];
        const int n_items = s_items;
        for (int r = rec0 + tid; r < rec1; r += kThreads) {
            const int first = (r - rec0) * kBwdChunk;
            chunk_info[r] = first < n_items ? make_int4(tile, range.x + first, min(kBwdChunk, n_items - first), 0)
                                            : make_int4(-1, -1, -1, -1);
        }
    }
    // loss partial of a half tile: lane i adds its pixels i, i + 32, i + 64, i + 96, then a shuffle tree.  One warp per
    // half this CTA owns (every configuration has at least two warps).
    __shared__ int s_last;
    {
        const int warp = tid >> 5, lane = tid & 31;
        for (int w = warp; w < 2 / kParts; w += kThreads / 32) {
            const int half = kParts == 2 ? part : w;
            float l = 0.f;
#pragma unroll
            for (int j = 0; j < kHalfPixels / 32; ++j) l += s_l[half * kHalfPixels + lane + 32 * j];
            l = warp_sum(l);
            if (lane == 0) tile_loss[2 * tile + half] = l;
        }
    }
    __syncthreads();
    if (tid == 0) {
        // the last CTA of the launch adds the partials to the caller's loss (in half-tile order: no float atomics, the
        // reference's 3 atomicAdds per pixel on one address become one add per launch)
        __threadfence();
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // 64 partial sums over the launch's half tiles h, h + 64, ... ; then a fixed tree
    const int h_begin = 2 * tile_y0 * v.tiles_x, h_end = h_begin + 2 * static_cast<int>(gridDim.x) / kParts;
    for (int t = tid; t < 64; t += kThreads) {
        float acc = 0.f;
        for (int h = h_begin + t; h < h_end; h += 64) acc += __ldcg(tile_loss + h);
        s_l[t] = acc;
    }
    __syncthreads();
    if (tid < 32) {
        float l = s_l[tid] + s_l[tid + 32];
        l = warp_sum(l);
        if (tid == 0) {
            *total_loss += l;  // the reference accumulates into the caller's value (gaussian_splatting_kernel.cu:68-70)
            *ticket = 0u;
        }
    }
}

// ---- backward: one thread per (tile, Gaussian) list entry ----------------------------------------
// Per pair (gaussian_splatting_kernel.cu:84-110, run() of the l1_norm root), with w = e * so:
//   cd_i = c_i w - (tgt_i - out_i)         (rest_sum += un-forwarded weighted_color adds 0: Q2)
//   s_i  = sign(cd_i)                      l1_norm_logic.cuh:29-37
//   d color_i += s_i w ; g_w = sum s_i c_i                  mul_logic.cuh:33-41
//   d so += g_w e ; g_d2 = -0.5 so (g_w e)                   exp / neg / mul_constant backward
//   d center -= g_d2 (2 ia dx + 2 ib dy , 2 ib dx + 2 ic dy)  mahalanobis_distance.cuh:100-113
//   d inv += g_d2 (dx^2, 2 dx dy, dy^2)                      :116-118
// Everything is LINEAR in the per-pair quantity t = g_w e with per-Gaussian (and per-row: dy) coefficients, so
// the pixel loop only accumulates   sum s_i e (3),  and per row  sum t, sum t dx, sum t dx^2 ;
// rows fold into  T0 = sum t, Tx, Ty, Txx, Txy, Tyy ; the conic / center / opacity / covariance chain rule is
// applied once per entry.  ~15 instructions per pair (13 on the FMA pipe per pixel, packed two pixels per
// instruction) instead of the reference's ~200 + 9 atomics.
// sign(): s_i e is formed by XOR-ing the sign bit of cd_i into e.  cd_i == 0 with e != 0
// (an exact cancellation c_i w == tgt_i - out_i) gets +-1 where the reference's l1_norm gives 0: one pair's term,
// inside the stated kink tolerance; with e == 0 every product is 0 either way.
// Packed fp32 (Blackwell FFMA2 / FADD2 / FMUL2 = PTX fma/add/mul.rn.f32x2): the pixel loop handles the pixels
// (2p, 2p+1) of a row in the two halves of 64-bit registers, so the ~15 floating-point operations per pixel
// cost 7.5 issue slots; only the exponential (MUFU) and the sign transfer (LOP3) work on single lanes.
// Measured on B200 (dev/pipe_lab.cu): FFMA2 issues at half the FFMA rate (same flops) but leaves the other issue
// slots to the ALU / MUFU pipes -- the scalar version of this loop was issue bound (ncu: 79 % issue active).
// Shared-memory layout of the residuals for the packed loop: per pixel pair (2p, 2p+1) two float4
//   A = {-rest.x[2p], -rest.x[2p+1], -rest.y[2p], -rest.y[2p+1]}   B = {-rest.z[2p], -rest.z[2p+1], m[2p], m[2p+1]}
// (negated so cd_i = fma(cs_i, e, -rest_i) takes them as the addend; m = 1 active / 0 inactive pixel).
struct RestPair {
    float4 A, B;
};

template <bool kMasked>
__device__ __forceinline__ void entry_tile_pass(const RestPair* __restrict__ s_rest, float px0, float py0, float cx,
                                                float cy, float A2, float B2, float C2, float cs0, float cs1,
                                                float cs2, float c0, float c1, float c2, float (&ac)[3],
                                                float (&T)[6]) {
    const float dx0 = px0 - cx;
    F2 dxs[kTile / 2];  // (px0 - cx) + j: one rounding more than the forward pass
#pragma unroll
    for (int p = 0; p < kTile / 2; ++p) dxs[p] = f2_pack(dx0 + static_cast<float>(2 * p), dx0 + static_cast<float>(2 * p + 1));
#if XYZ_BWD_DX2
    F2 dx2s[kTile / 2];  // dx^2, so that the two moment sums are one FMA each (costs 16 registers)
#pragma unroll
    for (int p = 0; p < kTile / 2; ++p) dx2s[p] = f2_mul(dxs[p], dxs[p]);
#endif
    const F2 A2p = f2_pack(A2, A2);
    const F2 cs0p = f2_pack(cs0, cs0), cs1p = f2_pack(cs1, cs1), cs2p = f2_pack(cs2, cs2);
    const F2 c0p = f2_pack(c0, c0), c1p = f2_pack(c1, c1), c2p = f2_pack(c2, c2);
    XYZ_UNROLL(XYZ_BWD_ROW_UNROLL)
    for (int r = 0; r < kTile; ++r) {
        const float dy = (py0 + static_cast<float>(r)) - cy;
        const float u = B2 * dy;
        const float t = (C2 * dy) * dy;
        const F2 up = f2_pack(u, u), tp = f2_pack(t, t);
        F2 acp0 = f2_pack(0.f, 0.f), acp1 = acp0, acp2 = acp0;  // this row's sum s_i e
        F2 Sx = acp0, Sxx = acp0;
#pragma unroll
        for (int p = 0; p < kTile / 2; ++p) {
            const RestPair rp = s_rest[r * (kTile / 2) + p];
            const F2 dx = dxs[p];
            const F2 arg = f2_fma(dx, f2_fma(A2p, dx, up), tp);
            float a_lo, a_hi;
            f2_unpack(arg, a_lo, a_hi);
            float e_lo = pair_exp(a_lo), e_hi = pair_exp(a_hi);
            if (kMasked) {  // 0 for pixels outside the image / row band
                e_lo *= rp.B.z;
                e_hi *= rp.B.w;
            }
            const F2 e = f2_pack(e_lo, e_hi);
            float d_lo, d_hi;
            f2_unpack(f2_fma(cs0p, e, f2_pack(rp.A.x, rp.A.y)), d_lo, d_hi);  // cd_0 of both pixels
            const F2 se0 = f2_pack(xor_sign(e_lo, d_lo), xor_sign(e_hi, d_hi));  // s_0 e
            f2_unpack(f2_fma(cs1p, e, f2_pack(rp.A.z, rp.A.w)), d_lo, d_hi);
            const F2 se1 = f2_pack(xor_sign(e_lo, d_lo), xor_sign(e_hi, d_hi));
            f2_unpack(f2_fma(cs2p, e, f2_pack(rp.B.x, rp.B.y)), d_lo, d_hi);
            const F2 se2 = f2_pack(xor_sign(e_lo, d_lo), xor_sign(e_hi, d_hi));
            acp0 = f2_add(acp0, se0);
            acp1 = f2_add(acp1, se1);
            acp2 = f2_add(acp2, se2);
            const F2 tt = f2_fma(c2p, se2, f2_fma(c1p, se1, f2_mul(c0p, se0)));  // t = g_w e = sum c_i (s_i e)
#if XYZ_BWD_DX2
            Sx = f2_fma(tt, dx, Sx);
            Sxx = f2_fma(tt, dx2s[p], Sxx);
#else
            const F2 tx = f2_mul(tt, dx);
            Sx = f2_add(Sx, tx);
            Sxx = f2_fma(tx, dx, Sxx);
#endif
        }
        // the row's sum of t follows from its sums of s_i e (t is linear in them): no per-pixel accumulator
        const float r0 = f2_hsum(acp0), r1 = f2_hsum(acp1), r2 = f2_hsum(acp2);
        ac[0] += r0;
        ac[1] += r1;
        ac[2] += r2;
        const float s0 = fmaf(c2, r2, fmaf(c1, r1, c0 * r0)), sx = f2_hsum(Sx), sxx = f2_hsum(Sxx);
        T[0] += s0;
        T[1] += sx;
        T[2] = fmaf(dy, s0, T[2]);
        T[3] += sxx;
        T[4] = fmaf(dy, sx, T[4]);
        T[5] = fmaf(dy * dy, s0, T[5]);
    }
}

// ---- backward work items (written by the forward pass, see splat_forward_kernel) ----
// The tile lists hold every pair whose weight is not EXACTLY zero (d2 <= 176: 13 sigma), which the forward pass needs for
// a bit-identical image.  A gradient is an atomically accumulated sum held to 1e-4 of the sum of its terms' magnitudes, and
// the terms of a pair carry the factor e = exp(-d2 / 2): beyond d2 = d2_bwd (default 48: e < 3.8e-11) they are far below
// the fp32 resolution of the sums they would join (splat_common.cuh: kD2Backward).  So the backward pass works on ITEMS =
// the list entries of a tile on which min d2 over the tile is within the bound -- 35 % of the entries at BASELINE's C4.
// The forward CTA of a tile has every entry's record in shared memory anyway: it takes the minimum of the conic form over
// the tile (closed form, conic_min_over_rect) and appends the surviving entries to the tile's slots of bwd_items.  An item
// is the Gaussian id (deterministic mode: the entry's index, which leads to the id and to the entry's row of
// entry_grads; rows of entries that are left out are zeroed by the forward pass).  The order of the items depends on warp
// timing and never enters a result.  XYZ_FLAG_BWD_ALL_PAIRS / XYZ_FLAG_NO_CULL pass d2_bwd = inf: every entry is an item.
// (Measured: items of 16 x 8 half tiles -- three lists per tile: both halves / upper / lower -- evaluate 14 % fewer
// pixels and take the same time, 273 against 275 us at C4: more, shorter, emptier CTAs.  Whole tiles it is.)
//
// The forward CTA also writes the tile's work records, one per backward CTA:
//     chunk_info[c] = {tile, slot of the CTA's first item, items (1 .. kBwdChunk), -}     (tile = -1: nothing to do)
// The tile scan sets aside ceil(len / kBwdChunk) records per tile, so the grid (all records) is an upper bound the host
// knows without reading anything back; at C4 two thirds of the CTAs find tile = -1 and leave at once.
__global__ void __launch_bounds__(kBwdChunk, XYZ_BWD_MINBLOCKS)
    splat_backward_kernel(SplatView v, const float4* __restrict__ records, const int4* __restrict__ chunk_info,
                          const int* __restrict__ bwd_items, const float4* __restrict__ rest_tiles,
                          const int* __restrict__ sorted_gid, const unsigned int* __restrict__ sorted_orig,
                          xyz_gaussian_grads* grads, float* __restrict__ entry_grads) {
    __shared__ RestPair s_rest[kTilePixels / 2];  // -(tgt - out) and the active mask, pixel pairs (see RestPair)

    const int tid = threadIdx.x;
    const int4 info = __ldg(chunk_info + blockIdx.x);
    const int tile = info.x;
    if (tile < 0) return;
    const bool valid = tid < info.z;
    const int item = valid ? __ldg(bwd_items + info.y + tid) : 0;
    const bool by_entry = entry_grads != nullptr;  // deterministic mode: the item is the entry's index
    const int tile_x = tile % v.tiles_x, tile_y = tile / v.tiles_x;
#pragma unroll
    for (int p = tid; p < kTilePixels; p += kBwdChunk) {
        const float4 rest = __ldg(rest_tiles + static_cast<size_t>(tile) * kTilePixels + p);
        float* pair = reinterpret_cast<float*>(&s_rest[p >> 1]) + (p & 1);
        pair[0] = -rest.x;
        pair[2] = -rest.y;
        pair[4] = -rest.z;
        pair[6] = rest.w;
    }
    const bool all_active = (tile_x * kTile + kTile <= v.width) && (tile_y * kTile >= v.row_begin) &&
                            (tile_y * kTile + kTile <= v.row_end);

    float cx = 0.f, cy = 0.f, ia = 0.f, ib = 0.f, ic = 0.f, so = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
    int g = 0;
    float4 r2 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) {
        g = by_entry ? sorted_gid[item] : item;
        const float4 r0 = __ldg(records + 4 * g), r1 = __ldg(records + 4 * g + 1);
        r2 = __ldg(records + 4 * g + 2);
        cx = r0.x; cy = r0.y; ia = r0.z; ib = r0.w;
        ic = r1.x; so = r1.y; c0 = r1.z; c1 = r1.w; c2 = r2.x;
    }
    __syncthreads();
    if (!valid) return;
    {
        float ac[3] = {0.f, 0.f, 0.f};
        float T[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // sum t {1, dx, dy, dx^2, dx dy, dy^2}
        {
            const float px0 = static_cast<float>(tile_x * kTile), py0 = static_cast<float>(tile_y * kTile);
            const float A2 = kKappa * ia, B2 = (2.0f * kKappa) * ib, C2 = kKappa * ic;
            const float cs0 = so * c0, cs1 = so * c1, cs2 = so * c2;
            if (all_active)
                entry_tile_pass<false>(s_rest, px0, py0, cx, cy, A2, B2, C2, cs0, cs1, cs2, c0, c1, c2, ac, T);
            else
                entry_tile_pass<true>(s_rest, px0, py0, cx, cy, A2, B2, C2, cs0, cs1, cs2, c0, c1, c2, ac, T);
        }
        // per-pair g_d2 = gamma * t with gamma = -0.5 so
        const float gamma = -0.5f * so;
        const float Gx = gamma * T[1], Gy = gamma * T[2];
        const float a_ia = gamma * T[3], a_ib = 2.0f * (gamma * T[4]), a_ic = gamma * T[5];
        const float ib2 = 2.0f * ib;
        const float a_cx = -(2.0f * ia * Gx + ib2 * Gy);
        const float a_cy = -(ib2 * Gx + 2.0f * ic * Gy);
        const float a_so = T[0];

        // Per-Gaussian chain rule, once per entry (linear in the sums above).
        // sym_matrix2_inv backward (sym_matrix2_inv_logic.cuh:43-77) needs Sigma = (A, B, C): rebuilt from exp(scale), cos
        // and sin AS THE PREPROCESS KERNEL COMPUTED THEM (IEEE expf / cosf / sinf in both flavours, kept in the record), so
        // the chain rule is applied at the Sigma whose inverse the forward pass rendered with -- the fast-math flavour's
        // ex2.approx / sin.approx (no range reduction) never enter it.  exp_logic.cuh:17-36, covariance_generation.cuh:154-172
        const float sn = __ldg(reinterpret_cast<const float*>(records + 4 * g + 3));
        const float es0 = r2.y, es1 = r2.z, ct = r2.w;
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
        out9[5] = so * ac[0];  // sum s_i w = so * sum s_i e
        out9[6] = so * ac[1];
        out9[7] = so * ac[2];
        out9[8] = a_so * (so * (1.0f - so));  // sigmoid_logic.cuh:27-36
        if (entry_grads) {
            float* row = entry_grads + static_cast<size_t>(sorted_orig[item]) * 9;
#pragma unroll
            for (int k = 0; k < 9; ++k) row[k] = out9[k];
        } else {
            float* gg = reinterpret_cast<float*>(grads + g);
#pragma unroll
            for (int k = 0; k < 9; ++k) atomicAdd(gg + k, out9[k]);  // VariableRef::add_grad, variable.cuh:48-50
        }
    }
}

}  // namespace

int XYZ_CAT(splat_forward_launch_, XYZ_SPLAT_FLAVOR)(const SplatView& v, const SplatBuffers& b, const float* target,
                                                     float* output, float* total_loss, unsigned int* ticket, bool deterministic,
                                                     float d2_bwd, int first_tile, const int* tile_order, cudaStream_t st) {
    const int ty0 = v.row_begin / kTile, ty1 = (v.row_end + kTile - 1) / kTile;
    if (ty1 <= ty0) return 0;
    // configuration by how many tiles an SM gets (see the kernel); XYZ_SPLAT_FWD_THREADS = 64 | 128 (threads of a
    // whole-tile CTA) or 32 (two half-tile CTAs of 64 threads) overrides.  Measured (dev/fwd_threads_sweep.py,
    // profiles/fwd_threads_sweep_r02.log): whole tiles with 64 threads win from 12 tiles per SM on, half tiles below;
    // 128 / 256 threads per tile (more shared-memory reads per pair) never win, nor does one warp per tile with eight
    // pixels per lane (fewer issue slots per pair, 96 registers: 388 us against 387 at C4).
    static const int forced = [] {
        const char* e = std::getenv("XYZ_SPLAT_FWD_THREADS");
        const int x = e ? std::atoi(e) : 0;
        return (x == 32 || x == 64 || x == 128) ? x : 0;
    }();
    const long long tiles = static_cast<long long>(v.tiles_x) * (ty1 - ty0);
    const int sms = sm_count();
    const int cfg = forced ? forced : (tiles >= 12LL * sms ? 64 : 32);
    const unsigned int grid = static_cast<unsigned int>(tiles) * (cfg == 32 ? 2u : 1u);
    const float d2s = -kKappa * d2_bwd;  // on the staged conic (scaled by kappa < 0); inf stays inf
#define XYZ_FWD_ARGS v, b.fwd_records, b.sorted_gid, b.tile_ranges, target, output, b.tile_loss, b.rest_tiles, ty0, ticket, total_loss, \
                     b.bwd_items, b.chunk_offsets, b.chunk_info, first_tile, d2s, b.vals_out, deterministic ? b.entry_grads : nullptr, \
                     tile_order
    if (cfg == 32) splat_forward_kernel<64, 2><<<grid, 64, 0, st>>>(XYZ_FWD_ARGS);
    else if (cfg == 64) splat_forward_kernel<64, 1><<<grid, 64, 0, st>>>(XYZ_FWD_ARGS);
    else splat_forward_kernel<128, 1><<<grid, 128, 0, st>>>(XYZ_FWD_ARGS);
#undef XYZ_FWD_ARGS
    count_launch();
    return last_error();
}

int XYZ_CAT(splat_backward_launch_, XYZ_SPLAT_FLAVOR)(const SplatView& v, const SplatBuffers& b, xyz_gaussian_grads* grads,
                                                      long long bwd_ctas, bool deterministic, cudaStream_t st) {
    if (bwd_ctas <= 0) return 0;
    splat_backward_kernel<<<static_cast<unsigned int>(bwd_ctas), kBwdChunk, 0, st>>>(
        v, b.records, b.chunk_info, b.bwd_items, b.rest_tiles, b.sorted_gid, b.vals_out, grads,
        deterministic ? b.entry_grads : nullptr);
    count_launch();
    return last_error();
}

}  // namespace xyzb
