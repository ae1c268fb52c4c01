// csrc/splat_host.cu -- host side of the splat pipeline: per-Gaussian preprocessing, the integer
// tile-binning work, scratch management and the C entry points that replace
// launch_gaussian_splatting (reference examples/mini-gaussian-splatting/gaussian_splatting_kernel.cu:114-149).
// See splat_common.cuh for the pipeline overview.  This translation unit is ALWAYS compiled with
// IEEE arithmetic: the tile rectangles are integer results that the CPU oracle reproduces
// bit for bit from the same record floats (oracle/xyz_oracle.cpp::tile_rect).
#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "splat_common.cuh"

namespace xyzb {
namespace {

// ---- 1. per-Gaussian records + tile rectangles -----------------------------------------------------
// Every float op that feeds an integer decision is a single rounded IEEE operation (__f*_rn: no
// FMA contraction), mirrored by oracle/xyz_oracle.cpp::tile_rect.
__device__ __forceinline__ int4 gaussian_tile_rect(float cx, float cy, float ia, float ib, float ic, const SplatView& v,
                                                   float d2max, int no_cull) {
    const int ty_lo = v.row_begin / kTile, ty_hi = (v.row_end + kTile - 1) / kTile;
    int4 r = make_int4(0, ty_lo, v.tiles_x, ty_hi);
    const float det = __fsub_rn(__fmul_rn(ia, ic), __fmul_rn(ib, ib));
    const bool ok = !no_cull && det > 0.0f && ia > 0.0f && ic > 0.0f && isfinite(det) && isfinite(cx) && isfinite(cy) &&
                    isfinite(ia) && isfinite(ic);
    if (ok) {
        const float hx = __fadd_rn(__fmul_rn(__fsqrt_rn(__fdiv_rn(__fmul_rn(d2max, ic), det)), 1.001f), 1.0f);
        const float hy = __fadd_rn(__fmul_rn(__fsqrt_rn(__fdiv_rn(__fmul_rn(d2max, ia), det)), 1.001f), 1.0f);
        if (isfinite(hx) && isfinite(hy)) {
            const float x_lo = floorf(__fsub_rn(cx, hx)), x_hi = ceilf(__fadd_rn(cx, hx));
            const float y_lo = floorf(__fsub_rn(cy, hy)), y_hi = ceilf(__fadd_rn(cy, hy));
            if (x_hi < 0.0f || y_hi < static_cast<float>(v.row_begin) || x_lo > static_cast<float>(v.width - 1) ||
                y_lo > static_cast<float>(v.row_end - 1)) {
                r = make_int4(0, 0, 0, 0);
            } else {
                const int xi0 = static_cast<int>(fmaxf(x_lo, 0.0f));
                const int xi1 = static_cast<int>(fminf(x_hi, static_cast<float>(v.width - 1)));
                const int yi0 = static_cast<int>(fmaxf(y_lo, static_cast<float>(v.row_begin)));
                const int yi1 = static_cast<int>(fminf(y_hi, static_cast<float>(v.row_end - 1)));
                r = make_int4(xi0 / kTile, yi0 / kTile, xi1 / kTile + 1, yi1 / kTile + 1);
            }
        }
    }
    return r;
}

// Exact-zero cull, second level: inside the rectangle, tile row ty only needs the tiles between the leftmost and
// the rightmost pixel of the ellipse d2 <= d2max over that row's pixel rows (the ellipse is convex, so the tiles of
// a row form one span; at C4 this drops 21 % of the rectangle's tiles).  For dy = y - cy the ellipse covers
//   dx in [-k dy - w(dy), -k dy + w(dy)],  k = ib / ia,  w = sqrt(d2max / ia - (det / ia^2) dy^2);
// the right edge is concave in dy with its maximum hx at dy_R = -(ib / ic) hx, the left edge is its mirror image.
// Same margins as the rectangle (0.1 % + 1 pixel); every float op is a single rounded IEEE operation, mirrored by
// oracle/xyz_oracle.cpp::tile_row_span.  Returns the half-open tile span [x, y) (empty: x >= y).
struct SpanCoef {
    float cx, cy, k, a, b, hxv, hyv_m, dyR;
    int ok;
};
__device__ __forceinline__ SpanCoef span_coef(float cx, float cy, float ia, float ib, float ic, float d2max, int no_cull) {
    SpanCoef c;
    c.cx = cx;
    c.cy = cy;
    const float det = __fsub_rn(__fmul_rn(ia, ic), __fmul_rn(ib, ib));
    c.ok = !no_cull && det > 0.0f && ia > 0.0f && ic > 0.0f && isfinite(det) && isfinite(cx) && isfinite(cy) &&
           isfinite(ia) && isfinite(ic);
    c.k = __fdiv_rn(ib, ia);
    c.a = __fdiv_rn(d2max, ia);
    c.b = __fdiv_rn(det, __fmul_rn(ia, ia));
    c.hxv = __fsqrt_rn(__fdiv_rn(__fmul_rn(d2max, ic), det));
    c.hyv_m = __fadd_rn(__fmul_rn(__fsqrt_rn(__fdiv_rn(__fmul_rn(d2max, ia), det)), 1.001f), 1.0f);
    c.dyR = __fmul_rn(-__fdiv_rn(ib, ic), c.hxv);
    if (!(isfinite(c.k) && isfinite(c.a) && isfinite(c.b) && isfinite(c.hxv) && isfinite(c.hyv_m) && isfinite(c.dyR))) c.ok = 0;
    return c;
}
__device__ __forceinline__ float span_edge(const SpanCoef& c, float dy, float sign) {
    const float rad = fmaxf(__fsub_rn(c.a, __fmul_rn(c.b, __fmul_rn(dy, dy))), 0.0f);
    return __fadd_rn(__fmul_rn(-c.k, dy), __fmul_rn(sign, __fsqrt_rn(rad)));
}
__device__ __forceinline__ int2 tile_row_span(const SpanCoef& c, const int4& r, int ty, const SplatView& v) {
    if (!c.ok) return make_int2(r.x, r.z);
    const int y0 = max(ty * kTile, v.row_begin), y1 = min(ty * kTile + kTile - 1, v.row_end - 1);
    const float lo = fmaxf(__fsub_rn(static_cast<float>(y0), c.cy), -c.hyv_m);
    const float hi = fminf(__fsub_rn(static_cast<float>(y1), c.cy), c.hyv_m);
    if (lo > hi) return make_int2(0, 0);
    const float fmx = (lo <= c.dyR && c.dyR <= hi) ? c.hxv : fmaxf(span_edge(c, lo, 1.0f), span_edge(c, hi, 1.0f));
    const float gmn = (lo <= -c.dyR && -c.dyR <= hi) ? -c.hxv : fminf(span_edge(c, lo, -1.0f), span_edge(c, hi, -1.0f));
    const float xr = ceilf(__fadd_rn(c.cx, __fadd_rn(fmx, __fadd_rn(__fmul_rn(fabsf(fmx), 0.001f), 1.0f))));
    const float xl = floorf(__fsub_rn(c.cx, __fadd_rn(-gmn, __fadd_rn(__fmul_rn(fabsf(gmn), 0.001f), 1.0f))));
    if (!(isfinite(xr) && isfinite(xl))) return make_int2(r.x, r.z);
    if (xr < 0.0f || xl > static_cast<float>(v.width - 1)) return make_int2(0, 0);
    const int xi0 = static_cast<int>(fmaxf(xl, 0.0f));
    const int xi1 = static_cast<int>(fminf(xr, static_cast<float>(v.width - 1)));
    return make_int2(max(xi0 / kTile, r.x), min(xi1 / kTile + 1, r.z));
}

// kCount (counting-sort binning, section 2b): the grid is one CTA per chunk of `chunk_size` consecutive Gaussians and
// the kernel also builds the chunk's tile histogram in shared memory -> hist[chunk][tile].  A thread counts the tiles
// of its own Gaussian right where it derives the spans; Gaussians with more than kBigGaussian tiles are left to the
// whole warp afterwards (one lane per tile row), so that a screen-filling Gaussian costs ~tiles_x + tiles_y steps
// instead of tiles_x * tiles_y.
constexpr unsigned int kBigGaussian = 192;

template <bool kCount>
__global__ void __launch_bounds__(256)
    splat_preprocess_kernel(SplatView v, const xyz_gaussian_params* __restrict__ params, float4* __restrict__ records,
                            int4* __restrict__ rects, unsigned int* __restrict__ touched, int2* __restrict__ spans,
                            float d2max, int no_cull, int chunk_size, int n_tiles, unsigned int* __restrict__ hist,
                            unsigned int* __restrict__ ticket, unsigned int* __restrict__ chunk_total,
                            float4* __restrict__ fwd_records, float kappa) {
    // kCount: n_tiles counters, one per tile of THIS launch's row band (tile ids relative to the band's first tile row:
    // a band of 1/8 of the image has 1/8 of the counters, of the histogram traffic and of the column scan)
    extern __shared__ unsigned int s_cnt[];
    __shared__ unsigned int s_total;
    const int tid = threadIdx.x;
    const int per_cta = kCount ? chunk_size : 256;
    const int g_begin = blockIdx.x * per_cta, g_end = min(g_begin + per_cta, v.num_gaussians);
    const int ty_lo = v.row_begin / kTile;
    const bool band_only = !no_cull && (v.row_begin > 0 || v.row_end < v.height);
    if (kCount) {
        if (blockIdx.x == 0 && tid == 0) *ticket = 0u;  // for the column-scan kernel that follows
        if (tid == 0) s_total = 0u;
        for (int t = tid; t < n_tiles; t += 256) s_cnt[t] = 0u;
        __syncthreads();
    }
    unsigned int my_total = 0u;
    for (int gb = g_begin; gb < g_end; gb += 256) {  // warp-uniform trip count
        const int g = gb + tid;
        const bool live = g < g_end;
        int4 r = make_int4(0, 0, 0, 0);
        SpanCoef sc{};
        unsigned int cnt = 0;
        bool big_one = false;  // rectangle of more than kBigGaussian tiles: counted by the whole warp below
        bool misses_band = false;
        if (live) {
            const xyz_gaussian_params p = params[g];
            // exp_logic.cuh:17-25, covariance_generation.cuh:154-172, sym_matrix2_inv_logic.cuh:21-40,
            // math.cuh:200-204 (sigmoid) -- computed once per Gaussian instead of once per pair per pass
            const float es0 = expf(p.scale[0]), es1 = expf(p.scale[1]);
            if (band_only) {
                // A launch that renders a row band of the image (one of G GPUs): 1 - 1/G of the Gaussians cannot reach it.
                // Sigma_yy <= max(exp s)^2, so the ellipse d2 <= d2max stays within sqrt(d2max) max(exp s) rows of the
                // centre; with the margins of gaussian_tile_rect (0.1 % + 1 pixel + the rounding of its own chain) a
                // Gaussian beyond that has the EMPTY rectangle there too -- same integer result, without the sine, the
                // inverse, the records and the spans.  Only for a covariance the det rule (Q15) leaves alone and whose
                // axes differ by less than 64 x (the rectangle's own "positive definite in fp32" test then holds with a
                // margin of 1e-3; anything else takes the full path and is kept as before).
                const float reach = fmaf(sqrtf(d2max) * fmaxf(es0, es1), 1.002f, 2.0f);
                const float cy = p.center[1];
                if (es0 * es1 > 1.1e-4f && fmaxf(es0, es1) < 64.0f * fminf(es0, es1) && isfinite(reach) && isfinite(cy) && isfinite(p.center[0]) &&
                    (cy + reach < static_cast<float>(v.row_begin) || cy - reach > static_cast<float>(v.row_end - 1))) {
                    rects[g] = make_int4(0, 0, 0, 0);
                    touched[g] = 0u;
                    misses_band = true;
                }
            }
            if (!misses_band) {
                const float ct = cosf(p.rotation[0]), sn = sinf(p.rotation[0]);
                const float m00 = __fmul_rn(es0, ct), m01 = __fmul_rn(-es1, sn), m10 = __fmul_rn(es0, sn), m11 = __fmul_rn(es1, ct);
                const float A = __fadd_rn(__fmul_rn(m00, m00), __fmul_rn(m01, m01));
                const float B = __fadd_rn(__fmul_rn(m00, m10), __fmul_rn(m01, m11));
                const float C = __fadd_rn(__fmul_rn(m10, m10), __fmul_rn(m11, m11));
                float det = __fsub_rn(__fmul_rn(A, C), __fmul_rn(B, B));
                if (fabsf(det) < 1e-8f) det = 1e-8f;
                const float inv_det = __fdiv_rn(1.0f, det);
                const float ia = __fmul_rn(C, inv_det), ib = __fmul_rn(-B, inv_det), ic = __fmul_rn(A, inv_det);
                const float so = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-p.opacity[0])));
                records[4 * g] = make_float4(p.center[0], p.center[1], ia, ib);
                records[4 * g + 1] = make_float4(ic, so, p.color[0], p.color[1]);
                // exp(scale), cos and sin ride along for the backward pass: its per-Gaussian chain rule is applied at exactly
                // the Sigma the forward pass used, in both math flavours (this translation unit is always IEEE)
                records[4 * g + 2] = make_float4(p.color[2], es0, es1, ct);
                records[4 * g + 3] = make_float4(sn, 0.f, 0.f, 0.f);
                // what the forward pass stages per Gaussian, ready to copy (cp.async): the conic pre-scaled by
                // kappa (-0.5 log2(e) for the fast-math flavour's ex2, -0.5 for expf) and the colours by sigmoid(opacity)
                fwd_records[2 * g] = make_float4(p.center[0], p.center[1], __fmul_rn(kappa, ia), __fmul_rn(2.0f * kappa, ib));
                fwd_records[2 * g + 1] = make_float4(__fmul_rn(kappa, ic), __fmul_rn(so, p.color[0]), __fmul_rn(so, p.color[1]),
                                                     __fmul_rn(so, p.color[2]));
                r = gaussian_tile_rect(p.center[0], p.center[1], ia, ib, ic, v, d2max, no_cull);
                rects[g] = r;
                // (a rectangle that misses the image / the row band is empty: no spans to derive -- 7 of 8 Gaussians when
                // one image is split over 8 GPUs)
                if (r.w > r.y && r.z > r.x) sc = span_coef(p.center[0], p.center[1], ia, ib, ic, d2max, no_cull);
                big_one = static_cast<unsigned int>((r.z - r.x) * (r.w - r.y)) > kBigGaussian;
                for (int ty = r.y; ty < r.w; ++ty) {
                    const int2 s = tile_row_span(sc, r, ty, v);
                    cnt += static_cast<unsigned int>(max(s.y - s.x, 0));
                    // the first kSpanRows rows are kept for the binning kernels (rows beyond that are recomputed there)
                    if (ty - r.y < kSpanRows) spans[static_cast<size_t>(g) * kSpanRows + (ty - r.y)] = make_int2(s.x, max(s.y, s.x));
                    if (kCount && !big_one)
                        for (int tx = s.x; tx < s.y; ++tx) atomicAdd(&s_cnt[(ty - ty_lo) * v.tiles_x + tx], 1u);
                }
                touched[g] = cnt;
                my_total += cnt;
            }
        }
        if (kCount) {
            unsigned int big = __ballot_sync(0xffffffffu, big_one);
            const int lane = tid & 31;
            while (big) {
                const int src = __ffs(big) - 1;
                big &= big - 1;
                int4 rb;
                SpanCoef cb;
                rb.x = __shfl_sync(0xffffffffu, r.x, src); rb.y = __shfl_sync(0xffffffffu, r.y, src);
                rb.z = __shfl_sync(0xffffffffu, r.z, src); rb.w = __shfl_sync(0xffffffffu, r.w, src);
                cb.cx = __shfl_sync(0xffffffffu, sc.cx, src); cb.cy = __shfl_sync(0xffffffffu, sc.cy, src);
                cb.k = __shfl_sync(0xffffffffu, sc.k, src); cb.a = __shfl_sync(0xffffffffu, sc.a, src);
                cb.b = __shfl_sync(0xffffffffu, sc.b, src); cb.hxv = __shfl_sync(0xffffffffu, sc.hxv, src);
                cb.hyv_m = __shfl_sync(0xffffffffu, sc.hyv_m, src); cb.dyR = __shfl_sync(0xffffffffu, sc.dyR, src);
                cb.ok = __shfl_sync(0xffffffffu, sc.ok, src);
                for (int ty = rb.y + lane; ty < rb.w; ty += 32) {
                    const int2 s = tile_row_span(cb, rb, ty, v);
                    for (int tx = s.x; tx < s.y; ++tx) atomicAdd(&s_cnt[(ty - ty_lo) * v.tiles_x + tx], 1u);
                }
            }
        }
    }
    if (kCount) {
        // the chunk's number of list entries (integer adds: order-independent), for the deterministic mode's
        // positions in Gaussian order (scanned over the chunks by the column-scan kernel's last CTA)
        my_total = warp_sum(my_total);
        if ((tid & 31) == 0 && my_total) atomicAdd(&s_total, my_total);
        __syncthreads();
        unsigned int* out = hist + static_cast<size_t>(blockIdx.x) * n_tiles;
        for (int t = tid; t < n_tiles; t += 256) out[t] = s_cnt[t];
        if (tid == 0) chunk_total[blockIdx.x] = s_total;
    }
}

// ---- 2a. radix path: (tile, Gaussian) keys in Gaussian order -----------------------------------------
// (more than kBinMaxTiles tiles, or XYZ_FLAG_RADIX_BINNING; section 2b is the default)
// Half a warp per Gaussian: the 16 lanes take the spans of 16 tile rows at a time (stored by the preprocess kernel
// for the first 16 rows, recomputed beyond), a shuffle scan gives each row its offset, then every row's span is
// written with consecutive lanes -> coalesced stores.  Order inside a Gaussian: row-major (ty, tx).
__global__ void __launch_bounds__(256)
    splat_emit_keys_kernel(SplatView v, const float4* __restrict__ records, const int4* __restrict__ rects,
                           const unsigned int* __restrict__ touched, const unsigned long long* __restrict__ offsets_incl,
                           const int2* __restrict__ spans, unsigned int* __restrict__ keys, unsigned int* __restrict__ vals,
                           float d2max, int no_cull, int by_gid) {
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int hl = threadIdx.x & 15;
    const unsigned int half_mask = 0xffffu << (threadIdx.x & 16);
    if (g >= v.num_gaussians) return;
    const unsigned int cnt = touched[g];
    if (cnt == 0u) return;
    const int4 r = rects[g];
    unsigned int o = static_cast<unsigned int>(offsets_incl[g] - cnt);
    for (int rb = r.y; rb < r.w; rb += kSpanRows) {
        const int ty = rb + hl;
        int2 s = make_int2(0, 0);
        if (ty < r.w) {
            if (rb == r.y) {
                s = spans[static_cast<size_t>(g) * kSpanRows + hl];
            } else {
                const float4 r0 = __ldg(records + 4 * g), r1 = __ldg(records + 4 * g + 1);
                const SpanCoef sc = span_coef(r0.x, r0.y, r0.z, r0.w, r1.x, d2max, no_cull);
                s = tile_row_span(sc, r, ty, v);
            }
        }
        const int wdt = max(s.y - s.x, 0);
        int incl = wdt;
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) {
            const int y = __shfl_up_sync(half_mask, incl, d, 16);
            if (hl >= d) incl += y;
        }
        const int excl = incl - wdt;
        const int rows = min(kSpanRows, r.w - rb);
        for (int j = 0; j < rows; ++j) {
            const int wj = __shfl_sync(half_mask, wdt, j, 16);
            const int oj = __shfl_sync(half_mask, excl, j, 16);
            const int sj = __shfl_sync(half_mask, s.x, j, 16);
            for (int t = hl; t < wj; t += 16) {
                const unsigned int e = o + static_cast<unsigned int>(oj + t);
                keys[e] = static_cast<unsigned int>((rb + j) * v.tiles_x + sj + t);
                // payload: the Gaussian id itself, or (deterministic mode) the entry's position in Gaussian
                // order, which names its row of entry_grads
                vals[e] = by_gid ? static_cast<unsigned int>(g) : e;
            }
        }
        o += static_cast<unsigned int>(__shfl_sync(half_mask, incl, 15, 16));
    }
}

// ---- 2b. stable counting sort by tile (the default binning for up to kBinMaxTiles tiles) --------------------------
// The keys are never materialised.  The Gaussians are cut into `n_chunks` consecutive chunks, one CTA each:
//   splat_preprocess_kernel<true>  also builds the per-chunk tile histogram in shared memory -> hist[chunk][tile]
//   splat_bin_colscan_kernel  per tile: exclusive prefix over the chunks (in place) + the tile's list length
//                             its last CTA: exclusive scan over the tiles -> tile_ranges, backward work-list offsets, total
//   splat_bin_scatter_kernel  every chunk walks its Gaussians again, ascending id; every tile is owned by one lane
//                             (see the kernel), slot = next[tile]++  -- ascending Gaussian id inside every tile list
//                             = exactly the stable sort the radix path produces (tested bit for bit against it and
//                             against the CPU restatement).  Also marks the unused backward work records.
constexpr int kBinThreads = 256;
constexpr int kBinBatch = 32;  // chunk sizes are multiples of this
// 4 bytes of shared memory per tile of the launch's row band (histogram / next free slot): up to 32 KB (8192 tiles, e.g.
// 2048 x 1024) the kernels keep several CTAs per SM; beyond that they opt in to the SM's full 227 KB (57 344 tiles =
// 14.7 Mpixel per band) with fewer CTAs per SM.  Only bands larger than that fall back to the radix path.
constexpr int kBinMaxTiles = 57344;
constexpr int kBinSmallTiles = 8192;
constexpr size_t kBinSmemBudget = 220 * 1024;
constexpr long long kBinLongList = 40LL << 20;  // predicted list length beyond which the chunks become one per SM

// one CTA of 1024 threads: exclusive scans over the tiles of (list length) and of the backward work records per tile
// `capacity` (0 = unlimited) is the list length the entry-sized buffers were sized for (XYZ_FLAG_ASYNC): a longer list
// publishes EMPTY tile ranges and an empty work list instead and bumps the sticky counter total_entries[1] (which never
// falls below `overflow_base`, the count the host has seen), so that no later kernel of the launch reads or writes
// beyond the buffers.
// Exclusive scan (64-bit) of n unsigned ints by one CTA of 1024 threads: out[i] = in[0] + ... + in[i - 1].
__device__ __forceinline__ void cta_exclusive_scan_u32_to_u64(const unsigned int* in, int n, unsigned long long* out, int tid) {
    __shared__ unsigned long long s_w[32];
    __shared__ unsigned long long s_c;
    const int lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_c = 0ull;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + tid;
        const unsigned long long mine = i < n ? static_cast<unsigned long long>(__ldcg(in + i)) : 0ull;
        unsigned long long x = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_w[warp] = x;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = s_w[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            s_w[lane] = w;
        }
        __syncthreads();
        const unsigned long long carry = s_c;
        if (i < n) out[i] = carry + (warp > 0 ? s_w[warp - 1] : 0ull) + (x - mine);
        __syncthreads();
        if (tid == 1023) s_c = carry + s_w[31];
        __syncthreads();
    }
}

// backward work records set aside for a tile with `len` list entries: the forward pass keeps at most `len` work items
// and each record takes kBwdChunk of them (splat_kernels.cuh)
__host__ __device__ __forceinline__ int bwd_records_of(unsigned int len) {
    return static_cast<int>((len + kBwdChunk - 1) / kBwdChunk);
}

__device__ __forceinline__ void bin_tilescan(const unsigned int* tile_total, int n_tiles, int2* __restrict__ tile_ranges,
                                             int* __restrict__ chunk_offsets, unsigned long long* __restrict__ total_entries,
                                             unsigned long long capacity, unsigned long long overflow_base, int tid,
                                             int* __restrict__ tile_order) {
    __shared__ unsigned long long s_warp[32];
    __shared__ int s_warp_c[32];
    __shared__ unsigned long long s_carry;
    __shared__ int s_carry_c;
    const int lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        s_carry = 0ull;
        s_carry_c = 0;
    }
    __syncthreads();
    for (int base = 0; base < n_tiles; base += 1024) {
        const int t = base + tid;
        const unsigned int len = t < n_tiles ? __ldcg(tile_total + t) : 0u;  // written by other CTAs of this launch
        const int c = bwd_records_of(len);
        unsigned long long x = len;
        int xc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
            const int yc = __shfl_up_sync(0xffffffffu, xc, o);
            if (lane >= o) {
                x += y;
                xc += yc;
            }
        }
        if (lane == 31) {
            s_warp[warp] = x;
            s_warp_c[warp] = xc;
        }
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = s_warp[lane];
            int wc = s_warp_c[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long y = __shfl_up_sync(0xffffffffu, w, o);
                const int yc = __shfl_up_sync(0xffffffffu, wc, o);
                if (lane >= o) {
                    w += y;
                    wc += yc;
                }
            }
            s_warp[lane] = w;  // inclusive over warps
            s_warp_c[lane] = wc;
        }
        __syncthreads();
        const unsigned long long carry = s_carry;
        const int carry_c = s_carry_c;
        if (t < n_tiles) {
            // beyond 2^31 entries the int ranges are meaningless; the host sees total_entries and refuses the launch
            const unsigned long long before = carry + (warp > 0 ? s_warp[warp - 1] : 0ull) + (x - len);
            tile_ranges[t] = make_int2(static_cast<int>(before), static_cast<int>(before + len));
            chunk_offsets[t] = carry_c + (warp > 0 ? s_warp_c[warp - 1] : 0) + (xc - c);
        }
        __syncthreads();
        if (tid == 1023) {
            s_carry = carry + s_warp[31];
            s_carry_c = carry_c + s_warp_c[31];
        }
        __syncthreads();
    }
    // Launch order of the forward CTAs: longest lists first (a counting sort into 1024 length classes), so that the
    // hardware's in-order CTA dispatch is longest-processing-time-first list scheduling and the kernel does not end with
    // a few SMs finishing the heaviest tiles.  Ties inside a class fall in arrival order; no result depends on the order.
    if (tile_order) {
        __shared__ unsigned int s_class[1024];
        __shared__ unsigned int s_cw[32];
        __shared__ unsigned int s_longest;
        s_class[tid] = 0u;
        if (tid == 0) s_longest = 0u;
        __syncthreads();
        unsigned int longest = 0u;
        for (int t = tid; t < n_tiles; t += 1024) longest = max(longest, __ldcg(tile_total + t));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) longest = max(longest, __shfl_xor_sync(0xffffffffu, longest, o));
        if (lane == 0) atomicMax(&s_longest, longest);
        __syncthreads();
        const float scale = s_longest ? 1023.0f / static_cast<float>(s_longest) : 0.0f;
        auto class_of = [&](unsigned int len) { return 1023u - min(1023u, static_cast<unsigned int>(static_cast<float>(len) * scale)); };
        for (int t = tid; t < n_tiles; t += 1024) atomicAdd(&s_class[class_of(__ldcg(tile_total + t))], 1u);
        __syncthreads();
        const unsigned int mine = s_class[tid];
        unsigned int x = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_cw[warp] = x;
        __syncthreads();
        if (warp == 0) {
            unsigned int w = s_cw[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            s_cw[lane] = w;
        }
        __syncthreads();
        s_class[tid] = (warp > 0 ? s_cw[warp - 1] : 0u) + (x - mine);  // first position of the class
        __syncthreads();
        for (int t = tid; t < n_tiles; t += 1024) tile_order[atomicAdd(&s_class[class_of(__ldcg(tile_total + t))], 1u)] = t;
    }
    const bool overflow = capacity != 0ull && s_carry > capacity;  // s_carry: settled by the loop's last barrier
    if (overflow) {
        for (int t = tid; t < n_tiles; t += 1024) {
            tile_ranges[t] = make_int2(0, 0);
            chunk_offsets[t] = 0;
        }
    }
    if (tid == 0) {
        chunk_offsets[n_tiles] = overflow ? 0 : s_carry_c;
        total_entries[0] = s_carry;
        // sticky counter, kept at or above the count the host has already seen (the scratch may have been reallocated)
        const unsigned long long seen = total_entries[1] > overflow_base ? total_entries[1] : overflow_base;
        total_entries[1] = overflow ? seen + 1ull : seen;
        total_entries[3] = capacity;  // what this launch was sized for (xyz_splat_workspace_status)
    }
}

// block (32, 32): 32 consecutive tiles x 32 groups of consecutive chunks
// The last CTA to finish (ticket) goes on to scan the tile totals: tile_ranges, the backward work-list offsets and the
// list length come out of the same launch.  `ticket` is zeroed by the preprocess kernel.
__global__ void __launch_bounds__(1024)
    splat_bin_colscan_kernel(unsigned int* __restrict__ hist, int n_chunks, int n_tiles, unsigned int* tile_total,
                             unsigned int* ticket, int2* __restrict__ tile_ranges, int* __restrict__ chunk_offsets,
                             unsigned long long* __restrict__ total_entries, unsigned long long capacity,
                             unsigned long long overflow_base, const unsigned int* chunk_total,
                             unsigned long long* __restrict__ chunk_base, int* __restrict__ tile_order) {
    __shared__ unsigned int s_part[32][33];
    __shared__ bool s_last;
    const int tx = threadIdx.x, gy = threadIdx.y;
    const int tile = blockIdx.x * 32 + tx;
    const int per = (n_chunks + 31) / 32;
    const int c0 = min(gy * per, n_chunks), c1 = min(c0 + per, n_chunks);
    // A thread owns `per` consecutive chunks of one tile column.  Up to kColRegs of them stay in registers between the
    // two passes: all loads of the column are in flight together and nothing is read twice (592 chunks: per = 19).  The
    // kernel is one latency chain (1 MB of histograms, 16 .. 128 CTAs): round trips are what it costs.
    constexpr int kColRegs = 24;
    unsigned int sum = 0;
    if (per <= kColRegs) {
        unsigned int x[kColRegs];
#pragma unroll
        for (int k = 0; k < kColRegs; ++k) {
            const int c = c0 + k;
            x[k] = (tile < n_tiles && c < c1) ? hist[static_cast<size_t>(c) * n_tiles + tile] : 0u;
        }
#pragma unroll
        for (int k = 0; k < kColRegs; ++k) sum += x[k];
        s_part[gy][tx] = sum;
        __syncthreads();
        unsigned int run = 0;
        for (int q = 0; q < gy; ++q) run += s_part[q][tx];
        if (tile < n_tiles) {
#pragma unroll
            for (int k = 0; k < kColRegs; ++k) {
                const int c = c0 + k;
                if (c < c1) hist[static_cast<size_t>(c) * n_tiles + tile] = run;
                run += x[k];
            }
            if (gy == 31) tile_total[tile] = run;
        }
    } else {
        if (tile < n_tiles) {
#pragma unroll 4
            for (int c = c0; c < c1; ++c) sum += hist[static_cast<size_t>(c) * n_tiles + tile];
        }
        s_part[gy][tx] = sum;
        __syncthreads();
        unsigned int run = 0;
        for (int q = 0; q < gy; ++q) run += s_part[q][tx];
        if (tile < n_tiles) {
#pragma unroll 4
            for (int c = c0; c < c1; ++c) {
                const size_t at = static_cast<size_t>(c) * n_tiles + tile;
                const unsigned int x = hist[at];
                hist[at] = run;
                run += x;
            }
            if (gy == 31) tile_total[tile] = run;
        }
    }
    __threadfence();
    __syncthreads();
    const int tid = gy * 32 + tx;
    if (tid == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    bin_tilescan(tile_total, n_tiles, tile_ranges, chunk_offsets, total_entries, capacity, overflow_base, tid, tile_order);
    // deterministic mode: list entries before every chunk, in Gaussian order (the scatter kernel turns them into
    // per-Gaussian positions) -- what used to be a library scan over all Gaussians
    if (chunk_base) cta_exclusive_scan_u32_to_u64(chunk_total, n_chunks, chunk_base, tid);
}

// Scatter: one CTA per chunk, next free slot of every tile in shared memory (start = tile begin + the chunk's column
// prefix).  Every warp OWNS a band of consecutive tile rows and walks the chunk's Gaussians in ascending id, taking
// only the spans inside its band: a tile is only ever written by its owner warp, in Gaussian order, and the tiles of
// one Gaussian are all different -- so `slot = next[tile]++` needs neither atomics nor block barriers, and every tile
// list comes out ascending in Gaussian id (= the stable sort).  Lanes of a warp: RL tile rows x XP column phases
// (8 x 4 for bands of 8 rows), i.e. up to 32 entries of one Gaussian per step.
template <bool kDeterministic>
__global__ void __launch_bounds__(kBinThreads)
    splat_bin_scatter_kernel(SplatView v, const float4* __restrict__ records, const int4* __restrict__ rects,
                             const int2* __restrict__ spans, const unsigned int* __restrict__ touched,
                             unsigned long long* __restrict__ offsets_incl, int chunk_size, int n_tiles, int tile0,
                             const unsigned int* __restrict__ hist, const int2* __restrict__ tile_ranges,
                             unsigned int* __restrict__ vals_out, int* __restrict__ sorted_gid, float d2max, int no_cull,
                             const int* __restrict__ chunk_offsets, int4* __restrict__ chunk_info, int chunk_info_size,
                             const unsigned long long* __restrict__ total_entries, unsigned long long capacity,
                             const unsigned long long* __restrict__ chunk_base) {
    extern __shared__ unsigned int s_next[];  // next free slot per tile
    __shared__ unsigned char s_hit[kBinThreads / 32][32];  // per warp: rank among the group's hits -> lane
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g_begin = blockIdx.x * chunk_size, g_end = min(g_begin + chunk_size, v.num_gaussians);
    // the backward work records of the tiles are written by the forward pass; the records beyond them: tile = -1
    for (int c = chunk_offsets[n_tiles] + blockIdx.x * kBinThreads + tid; c < chunk_info_size; c += gridDim.x * kBinThreads)
        chunk_info[c] = make_int4(-1, -1, -1, -1);
    if (capacity != 0ull && total_entries[0] > capacity) return;  // XYZ_FLAG_ASYNC overflow: every list is empty
    const unsigned int* mine = hist + static_cast<size_t>(blockIdx.x) * n_tiles;
    for (int t = tid; t < n_tiles; t += kBinThreads) s_next[t] = static_cast<unsigned int>(tile_ranges[t].x) + mine[t];
    if (kDeterministic) {
        // offsets_incl[g] = list entries of Gaussians 0 .. g (Gaussian order) = the end of g's block of entry_grads rows
        __shared__ unsigned long long s_scan_w[kBinThreads / 32];
        unsigned long long carry = chunk_base[blockIdx.x];
        for (int gb = g_begin; gb < g_end; gb += kBinThreads) {
            const int g = gb + tid;
            const unsigned long long mine_c = g < g_end ? touched[g] : 0u;
            unsigned long long xs = mine_c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long y = __shfl_up_sync(0xffffffffu, xs, o);
                if (lane >= o) xs += y;
            }
            __syncthreads();  // the previous round's readers of s_scan_w are done
            if (lane == 31) s_scan_w[warp] = xs;
            __syncthreads();
            unsigned long long before = 0ull, total = 0ull;
#pragma unroll
            for (int w = 0; w < kBinThreads / 32; ++w) {
                if (w < warp) before += s_scan_w[w];
                total += s_scan_w[w];
            }
            if (g < g_end) offsets_incl[g] = carry + before + xs;
            carry += total;
        }
    }
    __syncthreads();

    // this warp's band of tile rows (of the rows this launch renders), taken RL rows at a time: lane = (row rl, column
    // phase xl); a lane only ever touches the tiles (S0 + rl, x) with x = xl (mod XP), so consecutive Gaussians need
    // no synchronisation at all -- program order inside the lane is the Gaussian order.
    constexpr int kWarps = kBinThreads / 32;
    const int ty_lo = v.row_begin / kTile, ty_hi = (v.row_end + kTile - 1) / kTile;
    const int band = (ty_hi - ty_lo + kWarps - 1) / kWarps;
    const int R0 = ty_lo + warp * band, R1 = min(R0 + band, ty_hi);
    if (R0 >= R1) return;
    const int rl_bits = band >= 8 ? 3 : (band >= 4 ? 2 : (band >= 2 ? 1 : 0));
    const int RL = 1 << rl_bits, XP = 32 >> rl_bits;
    const int rl = lane >> (5 - rl_bits), xl = lane & (XP - 1);

    // Loads are batched so that their latency is paid once per group, not once per Gaussian: the rectangles of the
    // NEXT 32 candidates are in flight while this group is processed, and the spans of up to 8 hits are fetched by two
    // warp-wide loads (lane (rl, xl) fetches row rl of hits xl & 3 and 4 + (xl & 3)) and handed out by shuffles.
    auto load_rect = [&](int gb) {
        const int g = gb + lane;
        // an empty rectangle (culled, or outside the row band) has r.w == r.y and never hits; a non-empty rectangle whose
        // spans all turn out empty hits, finds nothing to write and costs a few instructions -- cheaper than a dependent
        // load of touched[g] in front of every rectangle
        int4 r = make_int4(0, 0, 0, 0);
        if (g < g_end) r = rects[g];
        return r;
    };
    for (int S0 = R0; S0 < R1; S0 += RL) {
        const int S1 = min(S0 + RL, R1);
        const int ty = S0 + rl;  // this lane's tile row
        unsigned int* row_next = s_next + (ty - ty_lo) * v.tiles_x;  // counters are relative to the band's first tile row
        int4 r_next = load_rect(g_begin);
        for (int gb = g_begin; gb < g_end; gb += 32) {
            const int4 r = r_next;
            if (gb + 32 < g_end) r_next = load_rect(gb + 32);
            const bool hit = r.y < S1 && r.w > S0 && r.w > r.y;
            const unsigned int hits = __ballot_sync(0xffffffffu, hit);
            const int n_hits = __popc(hits);
            __syncwarp();  // the previous group's readers of s_hit are done
            if (hit) s_hit[warp][__popc(hits & ((1u << lane) - 1u))] = static_cast<unsigned char>(lane);  // rank -> lane
            __syncwarp();
            for (int base = 0; base < n_hits; base += 8) {
                const int n_sub = min(8, n_hits - base);
                int2 staged[2];
                unsigned int staged_e[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    staged[q] = make_int2(0, 0);
                    staged_e[q] = 0u;
                    const int h = 4 * q + (lane & 3);
                    const int src = s_hit[warp][min(base + h, 31)] & 31;  // slots beyond n_hits hold stale lanes: guarded below
                    const int ry = __shfl_sync(0xffffffffu, r.y, src), rw = __shfl_sync(0xffffffffu, r.w, src);
                    if (h < n_sub && ty < S1 && ty >= ry && ty < rw) {
                        const int gg = gb + src, j = ty - ry;
                        int2 sp;
                        if (j < kSpanRows) {
                            sp = spans[static_cast<size_t>(gg) * kSpanRows + j];
                        } else {
                            const float4 r0 = __ldg(records + 4 * gg), r1 = __ldg(records + 4 * gg + 1);
                            const SpanCoef sc = span_coef(r0.x, r0.y, r0.z, r0.w, r1.x, d2max, no_cull);
                            sp = tile_row_span(sc, rects[gg], ty, v);
                        }
                        staged[q] = sp;
                        if (kDeterministic) {
                            // payload = the entry's position in Gaussian order (row-major inside a Gaussian)
                            unsigned int before = 0u;
                            for (int jj = 0; jj < min(j, kSpanRows); ++jj) {
                                const int2 t = spans[static_cast<size_t>(gg) * kSpanRows + jj];
                                before += static_cast<unsigned int>(max(t.y - t.x, 0));
                            }
                            if (j > kSpanRows) {
                                const float4 r0 = __ldg(records + 4 * gg), r1 = __ldg(records + 4 * gg + 1);
                                const SpanCoef sc = span_coef(r0.x, r0.y, r0.z, r0.w, r1.x, d2max, no_cull);
                                const int4 rr = rects[gg];
                                for (int jj = kSpanRows; jj < j; ++jj) {
                                    const int2 t = tile_row_span(sc, rr, ry + jj, v);
                                    before += static_cast<unsigned int>(max(t.y - t.x, 0));
                                }
                            }
                            staged_e[q] = static_cast<unsigned int>(offsets_incl[gg] - touched[gg]) + before;
                        }
                    }
                }
                for (int i = 0; i < n_sub; ++i) {
                    const int gg = gb + s_hit[warp][base + i];
                    // the span of (hit i, my row) was staged by lane (rl, i & 3) in slot i >> 2
                    const int from = (rl << (5 - rl_bits)) | (i & 3);
                    const int sx = __shfl_sync(0xffffffffu, (i >> 2) ? staged[1].x : staged[0].x, from);
                    const int sy = __shfl_sync(0xffffffffu, (i >> 2) ? staged[1].y : staged[0].y, from);
                    unsigned int eb = 0u;
                    if (kDeterministic) eb = __shfl_sync(0xffffffffu, (i >> 2) ? staged_e[1] : staged_e[0], from);
                    for (int x = sx + ((xl - sx) & (XP - 1)); x < sy; x += XP) {
                        const unsigned int pos = row_next[x];
                        row_next[x] = pos + 1u;
                        if (kDeterministic) {
                            vals_out[pos] = eb + static_cast<unsigned int>(x - sx);
                            sorted_gid[pos] = gg;
                        } else {
                            vals_out[pos] = static_cast<unsigned int>(gg);
                        }
                    }
                }
            }
        }
    }
}

// entry index in Gaussian order -> Gaussian id (binary search over the inclusive scan)
__device__ __forceinline__ int entry_to_gaussian(const unsigned long long* __restrict__ offsets_incl, int n,
                                                 unsigned int e) {
    int lo = 0, hi = n - 1;  // first g with offsets_incl[g] > e
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (offsets_incl[mid] > e) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// ---- per-tile ranges + Gaussian id per sorted entry ---------------------------------------------------
__global__ void __launch_bounds__(256)
    splat_ranges_kernel(long long entries, const unsigned int* __restrict__ keys_sorted,
                        const unsigned int* __restrict__ vals_sorted, const unsigned long long* __restrict__ offsets_incl, int n,
                        int2* __restrict__ tile_ranges, int* __restrict__ sorted_gid) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= entries) return;
    const unsigned int k = keys_sorted[i];
    if (i == 0 || keys_sorted[i - 1] != k) tile_ranges[k].x = static_cast<int>(i);
    if (i == entries - 1 || keys_sorted[i + 1] != k) tile_ranges[k].y = static_cast<int>(i + 1);
    // sorted_gid == nullptr: the sort payload already is the Gaussian id
    if (sorted_gid) sorted_gid[i] = entry_to_gaussian(offsets_incl, n, vals_sorted[i]);
}

// ---- backward work records (radix path): exclusive scan over the tiles of bwd_records_of(list length) ------------
__global__ void __launch_bounds__(1024)
    splat_chunk_scan_kernel(const int2* __restrict__ tile_ranges, int n_tiles, int* __restrict__ chunk_offsets) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n_tiles; base += 1024) {
        const int t = base + tid;
        int c = 0;
        if (t < n_tiles) {
            const int2 r = tile_ranges[t];
            c = bwd_records_of(static_cast<unsigned int>(r.y - r.x));
        }
        int x = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            s_warp[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const int carry = s_carry;
        const int before = carry + (warp > 0 ? s_warp[warp - 1] : 0) + (x - c);
        if (t < n_tiles) chunk_offsets[t] = before;
        __syncthreads();
        if (tid == 1023) s_carry = carry + s_warp[31];
        __syncthreads();
    }
    if (tid == 0) chunk_offsets[n_tiles] = s_carry;
}

// ---- deterministic mode: per-Gaussian sum of its entries' rows, in entry order ----------------------------
__global__ void __launch_bounds__(256)
    splat_grads_finish_kernel(int n, const unsigned int* __restrict__ touched,
                              const unsigned long long* __restrict__ offsets_incl, const float* __restrict__ entry_grads,
                              xyz_gaussian_grads* grads) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const unsigned int end = static_cast<unsigned int>(offsets_incl[g]), begin = end - touched[g];
    float s[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (unsigned int e = begin; e < end; ++e) {
#pragma unroll
        for (int k = 0; k < 9; ++k) s[k] += entry_grads[static_cast<size_t>(e) * 9 + k];
    }
    float* gg = reinterpret_cast<float*>(grads + g);
#pragma unroll
    for (int k = 0; k < 9; ++k) gg[k] += s[k];
}

struct ToU64 {
    __host__ __device__ unsigned long long operator()(unsigned int x) const { return x; }
};
using TouchedIter = cub::TransformInputIterator<unsigned long long, ToU64, const unsigned int*>;  // radix path only

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// ---- the plan of one launch: what runs and where every buffer lives --------------------------------------------------
// A workspace is  [header 256 B][fixed part: depends on N, the image and the chunking][entry part: depends on the list
// capacity].  The library-owned scratch of the classic entry points keeps the two parts in two arenas (the entry part
// grows with the scene); xyz_launch_gaussian_splatting_ws lays both out in the caller's buffer.
constexpr size_t kHeaderBytes = 256;  // u64 {list length, overflow counter (sticky), ticket, capacity of that launch}

struct SplatPlan {
    SplatView v;
    int n_tiles;             // whole image
    int tile0, n_tiles_l;    // tiles of the launch's row band: global ids [tile0, tile0 + n_tiles_l)
    bool counting, deterministic, precise;
    int no_cull;
    float d2max;
    float d2_bwd;            // backward cull (splat_kernels.cuh); infinity = every listed pair
    int chunk_size, n_chunks;
    size_t o_rec, o_frec, o_rect, o_touched, o_spans, o_offsets, o_ranges, o_tloss, o_chunks, o_rest, o_hist, o_ttotal, o_ctotal,
        o_cbase, o_order, o_scan_tmp;
    size_t scan_tmp_bytes;
    size_t fixed_bytes;  // without the header
};

struct EntryLayout {
    size_t o_kin, o_kout, o_vin, o_vout, o_gid, o_cinfo, o_eg, o_items, o_sort_tmp;
    size_t sort_tmp_bytes;
    size_t bytes;
    int key_bits;
    int chunk_info_size;
};

// predicted_entries: only a tuning input (CTAs per SM of the counting sort); never changes a result bit.
int make_plan(SplatPlan& p, int W, int H, int N, int row_begin, int row_end, int flags, long long predicted_entries,
              cudaStream_t st) {
    if (W <= 0 || H <= 0 || N < 0 || row_begin < 0 || row_end > H || row_begin > row_end) return XYZ_ERR_INVALID_ARGUMENT;
    p.v.width = W; p.v.height = H; p.v.num_gaussians = N;
    p.v.row_begin = row_begin; p.v.row_end = row_end;
    p.v.tiles_x = (W + kTile - 1) / kTile;
    p.v.tiles_y = (H + kTile - 1) / kTile;
    p.n_tiles = p.v.tiles_x * p.v.tiles_y;
    p.precise = (flags & XYZ_FLAG_PRECISE_MATH) != 0;
    p.deterministic = (flags & XYZ_FLAG_DETERMINISTIC) != 0;
    p.no_cull = (flags & XYZ_FLAG_NO_CULL) ? 1 : 0;
    p.d2max = (flags & XYZ_FLAG_TAIL_CULL) ? kD2MaxTail : (p.precise ? kD2MaxPrecise : kD2MaxFast);
    static const float d2_bwd_default = [] {  // measuring knob (dev/bwd_cull_sweep.py): the bound of the backward cull
        const char* e = std::getenv("XYZ_SPLAT_BWD_D2");
        const float x = e ? static_cast<float>(std::atof(e)) : 0.0f;
        return x > 0.0f ? x : kD2Backward;
    }();
    p.d2_bwd = (flags & (XYZ_FLAG_BWD_ALL_PAIRS | XYZ_FLAG_NO_CULL)) ? INFINITY : d2_bwd_default;
    const int ty_lo = row_begin / kTile, ty_hi = (row_end + kTile - 1) / kTile;
    const int band_tiles = (ty_hi - ty_lo) * p.v.tiles_x;
    if (band_tiles <= 0) p.v.num_gaussians = N = 0;  // an empty row band renders nothing: no kernel has work
    // binning: stable counting sort by tile (default) or, for very large tile counts / on request, the radix path.
    // The counting sort appends 4 bytes at a time to n_chunks x n_tiles output streams; it is fast as long as the open
    // 32-byte sectors of all streams stay in L2 until they are complete, so long lists get fewer, longer chunks
    // (measured at 1024^2, dev/splat_knob3m.sh: 4 CTAs per SM win up to ~3e7 entries, 1 per SM beyond -- at 1.2e8
    // entries 4 per SM are 0.8 ms SLOWER than 1 per SM and lose to the radix sort).
    p.counting = !(flags & XYZ_FLAG_RADIX_BINNING) && band_tiles <= kBinMaxTiles;
    p.tile0 = p.counting ? ty_lo * p.v.tiles_x : 0;
    p.n_tiles_l = p.counting ? band_tiles : p.n_tiles;
    const int ng = N > 0 ? N : 1;
    p.chunk_size = kBinBatch;
    p.n_chunks = 1;
    if (p.counting) {
        static const int forced = [] {  // tuning knob: CTAs of the count / scatter kernels per SM
            const char* e = std::getenv("XYZ_SPLAT_BIN_CTAS_PER_SM");
            const int x = e ? std::atoi(e) : 0;
            return x > 0 && x <= 64 ? x : 0;
        }();
        int per_sm = forced ? forced : (predicted_entries <= kBinLongList ? 4 : 1);
        if (band_tiles > kBinSmallTiles)  // CTAs that fit next to each other with 4 bytes of shared memory per tile
            per_sm = std::max(1, std::min(per_sm, static_cast<int>(kBinSmemBudget / (sizeof(unsigned int) * band_tiles))));
        const int want = per_sm * sm_count();
        p.chunk_size = ((ng + want - 1) / want + kBinBatch - 1) / kBinBatch * kBinBatch;
        p.n_chunks = (ng + p.chunk_size - 1) / p.chunk_size;
    }
    size_t off = 0;
    auto take = [&off](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
    const size_t nl = static_cast<size_t>(p.n_tiles_l > 0 ? p.n_tiles_l : 1);
    p.o_rec = take(sizeof(float4) * 4 * ng);
    p.o_frec = take(sizeof(float4) * 2 * ng);
    p.o_rect = take(sizeof(int4) * ng);
    p.o_touched = take(sizeof(unsigned int) * ng);
    p.o_spans = take(sizeof(int2) * kSpanRows * ng);
    p.o_offsets = take((p.deterministic || !p.counting) ? sizeof(unsigned long long) * ng : 0);
    p.o_ranges = take(sizeof(int2) * p.n_tiles);
    p.o_tloss = take(sizeof(float) * 2 * p.n_tiles);  // one loss partial per half tile
    p.o_chunks = take(sizeof(int) * (nl + 1));
    p.o_rest = take(sizeof(float4) * kTilePixels * static_cast<size_t>(p.n_tiles));
    p.o_hist = take(p.counting ? sizeof(unsigned int) * static_cast<size_t>(p.n_chunks) * nl : 0);
    p.o_ttotal = take(sizeof(unsigned int) * nl);
    p.o_ctotal = take(sizeof(unsigned int) * p.n_chunks);
    p.o_cbase = take(sizeof(unsigned long long) * p.n_chunks);
    p.o_order = take(sizeof(int) * nl);  // launch order of the forward CTAs (counting path)
    p.scan_tmp_bytes = 0;
    if (!p.counting)
        cub::DeviceScan::InclusiveSum(nullptr, p.scan_tmp_bytes, TouchedIter(nullptr, ToU64()),
                                      static_cast<unsigned long long*>(nullptr), ng, st);
    p.o_scan_tmp = take(p.counting ? 0 : p.scan_tmp_bytes + 16);
    p.fixed_bytes = off;
    return 0;
}

// capacity = list entries the entry part is laid out for
EntryLayout make_entry_layout(const SplatPlan& p, long long capacity, cudaStream_t st) {
    EntryLayout L{};
    L.key_bits = 1;
    while ((1 << L.key_bits) < p.n_tiles) ++L.key_bits;
    const long long ne = capacity > 0 ? capacity : 1;
    size_t soff = 0;
    auto stake = [&soff](size_t bytes) { size_t o = soff; soff += align_up(bytes); return o; };
    L.o_kin = stake(p.counting ? 0 : 4 * ne);
    L.o_kout = stake(p.counting ? 0 : 4 * ne);
    L.o_vin = stake(p.counting ? 0 : 4 * ne);
    L.o_vout = stake(4 * ne);
    L.o_gid = stake(p.deterministic ? 4 * ne : 0);
    L.chunk_info_size = static_cast<int>(capacity / kBwdChunk + p.n_tiles_l);  // an upper bound of the sum of ceil(len / chunk)
    L.o_cinfo = stake(sizeof(int4) * static_cast<size_t>(ne / kBwdChunk + p.n_tiles_l));
    L.o_eg = stake(p.deterministic ? 36 * ne : 0);
    L.o_items = stake(4 * ne);  // the backward work items: at most one per entry
    L.sort_tmp_bytes = 0;
    if (!p.counting)
        cub::DeviceRadixSort::SortPairs(nullptr, L.sort_tmp_bytes, static_cast<unsigned int*>(nullptr),
                                        static_cast<unsigned int*>(nullptr), static_cast<unsigned int*>(nullptr),
                                        static_cast<unsigned int*>(nullptr), static_cast<int>(ne), 0, L.key_bits, st);
    L.o_sort_tmp = stake(p.counting ? 0 : L.sort_tmp_bytes + 16);
    L.bytes = soff;
    return L;
}

// XYZ_FLAG_TIMING: CUDA events between the stages of a launch (a measuring aid: the events cost a little themselves)
struct StageEvents {
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // begin | preprocess | scans | scatter | forward | backward
    bool armed = false;
    int ensure() {
        for (auto& e : ev)
            if (!e) {
                cudaError_t ce = cudaEventCreate(&e);
                if (ce != cudaSuccess) return static_cast<int>(ce);
            }
        return 0;
    }
};
inline void mark(StageEvents* t, int i, cudaStream_t st) {
    if (t) cudaEventRecord(t->ev[i], st);
}

struct Bound {  // a plan bound to memory
    SplatBuffers b{};
    unsigned long long* header = nullptr;  // {list length, overflow counter, ticket, capacity}
    unsigned int* hist = nullptr;
    unsigned int* tile_total = nullptr;
    unsigned int* chunk_total = nullptr;
    unsigned long long* chunk_base = nullptr;
    unsigned char* scan_tmp = nullptr;
    unsigned char* sort_tmp = nullptr;
};

void bind_fixed(const SplatPlan& p, unsigned char* header, unsigned char* base, Bound& o) {
    o.header = reinterpret_cast<unsigned long long*>(header);
    o.b.records = reinterpret_cast<float4*>(base + p.o_rec);
    o.b.fwd_records = reinterpret_cast<float4*>(base + p.o_frec);
    o.b.rects = reinterpret_cast<int4*>(base + p.o_rect);
    o.b.touched = reinterpret_cast<unsigned int*>(base + p.o_touched);
    o.b.spans = reinterpret_cast<int2*>(base + p.o_spans);
    o.b.offsets = reinterpret_cast<unsigned long long*>(base + p.o_offsets);
    o.b.tile_ranges = reinterpret_cast<int2*>(base + p.o_ranges);
    o.b.tile_loss = reinterpret_cast<float*>(base + p.o_tloss);
    o.b.chunk_offsets = reinterpret_cast<int*>(base + p.o_chunks);
    o.b.rest_tiles = reinterpret_cast<float4*>(base + p.o_rest);
    o.hist = reinterpret_cast<unsigned int*>(base + p.o_hist);
    o.tile_total = reinterpret_cast<unsigned int*>(base + p.o_ttotal);
    o.chunk_total = reinterpret_cast<unsigned int*>(base + p.o_ctotal);
    o.chunk_base = reinterpret_cast<unsigned long long*>(base + p.o_cbase);
    o.b.tile_order = reinterpret_cast<int*>(base + p.o_order);
    o.scan_tmp = base + p.o_scan_tmp;
}

void bind_entry(const SplatPlan& p, const EntryLayout& L, unsigned char* sbase, Bound& o) {
    o.b.keys_in = reinterpret_cast<unsigned int*>(sbase + L.o_kin);
    o.b.keys_out = reinterpret_cast<unsigned int*>(sbase + L.o_kout);
    o.b.vals_in = reinterpret_cast<unsigned int*>(sbase + L.o_vin);
    o.b.vals_out = reinterpret_cast<unsigned int*>(sbase + L.o_vout);
    // fast mode: the sort payload is the Gaussian id; deterministic mode: payload = entry position, ids kept apart
    o.b.sorted_gid = p.deterministic ? reinterpret_cast<int*>(sbase + L.o_gid) : reinterpret_cast<int*>(o.b.vals_out);
    o.b.entry_grads = p.deterministic ? reinterpret_cast<float*>(sbase + L.o_eg) : nullptr;
    o.b.chunk_info = reinterpret_cast<int4*>(sbase + L.o_cinfo);
    o.b.bwd_items = reinterpret_cast<int*>(sbase + L.o_items);
    o.sort_tmp = sbase + L.o_sort_tmp;
}

// ---- front half: records, rectangles, histograms, tile ranges, the list length (on the device) -----------------------
// capacity != 0: the entry part holds `capacity` entries and nobody will look at the length before the back half runs
// (sync-free launches): a longer list is published EMPTY and the header's overflow counter goes up.
int enqueue_front(const SplatPlan& p, const Bound& o, const xyz_gaussian_params* gaussians, unsigned long long capacity,
                  unsigned long long overflow_base, cudaStream_t st, const unsigned long long** total_src,
                  StageEvents* tm) {
    const int N = p.v.num_gaussians;
    const float kappa = p.precise ? kKappaPrecise : kKappaFast;
    *total_src = o.header;
    mark(tm, 0, st);
    if (N <= 0) {
        mark(tm, 1, st);
        mark(tm, 2, st);
        return 0;
    }
    if (p.counting) {
        if (p.n_tiles_l > kBinSmallTiles) {  // more than 32 KB of counters: opt in to the large shared-memory carve-out
            static const cudaError_t once = [] {
                const int bytes = static_cast<int>(sizeof(unsigned int) * kBinMaxTiles);
                cudaError_t e = cudaFuncSetAttribute(splat_preprocess_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
                if (e == cudaSuccess)
                    e = cudaFuncSetAttribute(splat_bin_scatter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
                if (e == cudaSuccess)
                    e = cudaFuncSetAttribute(splat_bin_scatter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
                return e;
            }();
            if (once != cudaSuccess) return static_cast<int>(once);
        }
        splat_preprocess_kernel<true><<<p.n_chunks, 256, sizeof(unsigned int) * p.n_tiles_l, st>>>(
            p.v, gaussians, o.b.records, o.b.rects, o.b.touched, o.b.spans, p.d2max, p.no_cull, p.chunk_size, p.n_tiles_l,
            o.hist, reinterpret_cast<unsigned int*>(o.header + 2), o.chunk_total, o.b.fwd_records, kappa);
        count_launch();
        mark(tm, 1, st);
        splat_bin_colscan_kernel<<<(p.n_tiles_l + 31) / 32, dim3(32, 32), 0, st>>>(
            o.hist, p.n_chunks, p.n_tiles_l, o.tile_total, reinterpret_cast<unsigned int*>(o.header + 2),
            o.b.tile_ranges + p.tile0, o.b.chunk_offsets, o.header, capacity, overflow_base, o.chunk_total,
            p.deterministic ? o.chunk_base : nullptr, o.b.tile_order);
        count_launch();
        mark(tm, 2, st);
        return last_error();
    }
    // radix path: positions in Gaussian order come from a library scan (more than kBinMaxTiles tiles / on request)
    cudaError_t ce = cudaMemsetAsync(o.b.tile_ranges, 0, sizeof(int2) * p.n_tiles, st);  // empty tiles: (0, 0)
    if (ce != cudaSuccess) return static_cast<int>(ce);
    splat_preprocess_kernel<false><<<(N + 255) / 256, 256, 0, st>>>(p.v, gaussians, o.b.records, o.b.rects, o.b.touched,
                                                                    o.b.spans, p.d2max, p.no_cull, 0, 0, nullptr, nullptr,
                                                                    nullptr, o.b.fwd_records, kappa);
    count_launch();
    mark(tm, 1, st);
    size_t tmp = p.scan_tmp_bytes;
    ce = cub::DeviceScan::InclusiveSum(o.scan_tmp, tmp, TouchedIter(o.b.touched, ToU64()), o.b.offsets, N, st);
    if (ce != cudaSuccess) return static_cast<int>(ce);
    count_launch();
    mark(tm, 2, st);
    *total_src = o.b.offsets + (N - 1);
    return last_error();
}

// ---- back half: lists, forward, loss, backward ------------------------------------------------------------------------
// entries: the list length where the host knows it (radix path: must), else the capacity the buffers were sized for.
int enqueue_back(const SplatPlan& p, const EntryLayout& L, const Bound& o, const xyz_gaussian_params* gaussians,
                 xyz_gaussian_grads* gradients, const float* target, float* output, float* total_loss, long long entries,
                 unsigned long long capacity, cudaStream_t st, StageEvents* tm) {
    const int N = p.v.num_gaussians;
    const SplatBuffers& b = o.b;
    if (entries > 0) {
        if (p.counting) {
            const size_t smem = sizeof(unsigned int) * p.n_tiles_l;  // <= 224 KB (kBinMaxTiles; opt-in set by enqueue_front)
            if (p.deterministic)
                splat_bin_scatter_kernel<true><<<p.n_chunks, kBinThreads, smem, st>>>(
                    p.v, b.records, b.rects, b.spans, b.touched, b.offsets, p.chunk_size, p.n_tiles_l, p.tile0, o.hist,
                    b.tile_ranges + p.tile0, b.vals_out, b.sorted_gid, p.d2max, p.no_cull, b.chunk_offsets, b.chunk_info,
                    L.chunk_info_size, o.header, capacity, o.chunk_base);
            else
                splat_bin_scatter_kernel<false><<<p.n_chunks, kBinThreads, smem, st>>>(
                    p.v, b.records, b.rects, b.spans, b.touched, nullptr, p.chunk_size, p.n_tiles_l, p.tile0, o.hist,
                    b.tile_ranges + p.tile0, b.vals_out, nullptr, p.d2max, p.no_cull, b.chunk_offsets, b.chunk_info,
                    L.chunk_info_size, o.header, capacity, nullptr);
            count_launch();
        } else {
            splat_emit_keys_kernel<<<(N + 15) / 16, 256, 0, st>>>(p.v, b.records, b.rects, b.touched, b.offsets, b.spans,
                                                                  b.keys_in, b.vals_in, p.d2max, p.no_cull,
                                                                  p.deterministic ? 0 : 1);
            count_launch();
            size_t tmp = L.sort_tmp_bytes;
            cudaError_t ce = cub::DeviceRadixSort::SortPairs(o.sort_tmp, tmp, b.keys_in, b.keys_out, b.vals_in, b.vals_out,
                                                             static_cast<int>(entries), 0, L.key_bits, st);
            if (ce != cudaSuccess) return static_cast<int>(ce);
            count_launch(3);
            splat_ranges_kernel<<<static_cast<unsigned int>((entries + 255) / 256), 256, 0, st>>>(
                entries, b.keys_out, b.vals_out, b.offsets, N, b.tile_ranges, p.deterministic ? b.sorted_gid : nullptr);
            splat_chunk_scan_kernel<<<1, 1024, 0, st>>>(b.tile_ranges, p.n_tiles, b.chunk_offsets);
            count_launch(2);
            // work records nobody writes (the forward pass fills in those of its tiles) read tile = -1
            ce = cudaMemsetAsync(b.chunk_info, 0xff, sizeof(int4) * static_cast<size_t>(L.chunk_info_size), st);
            if (ce != cudaSuccess) return static_cast<int>(ce);
        }
    } else if (p.counting && N > 0) {
        // no entries: the column scan has published empty ranges for the band already
    } else {
        const int ty0 = p.v.row_begin / kTile, ty1 = (p.v.row_end + kTile - 1) / kTile;
        if (ty1 > ty0) {
            cudaError_t ce = cudaMemsetAsync(b.tile_ranges + ty0 * p.v.tiles_x, 0,
                                             sizeof(int2) * static_cast<size_t>(ty1 - ty0) * p.v.tiles_x, st);
            if (ce != cudaSuccess) return static_cast<int>(ce);
        }
    }
    mark(tm, 3, st);
    unsigned int* fwd_ticket = reinterpret_cast<unsigned int*>(o.header + 2) + 1;  // next to the column scan's ticket
    // (the launch order comes out of the counting path's tile scan, which only runs when there are Gaussians)
    const int* order = (p.counting && N > 0) ? b.tile_order : nullptr;
    int err = p.precise ? splat_forward_launch_precise(p.v, b, target, output, total_loss, fwd_ticket, p.deterministic, p.d2_bwd, p.tile0, order, st)
                        : splat_forward_launch_fast(p.v, b, target, output, total_loss, fwd_ticket, p.deterministic, p.d2_bwd, p.tile0, order, st);
    if (err) return err;
    mark(tm, 4, st);
    const long long bwd_ctas = entries > 0 ? static_cast<long long>(L.chunk_info_size) : 0;
    err = p.precise ? splat_backward_launch_precise(p.v, b, gradients, bwd_ctas, p.deterministic, st)
                    : splat_backward_launch_fast(p.v, b, gradients, bwd_ctas, p.deterministic, st);
    if (err) return err;
    if (p.deterministic && entries > 0) {
        splat_grads_finish_kernel<<<(N + 255) / 256, 256, 0, st>>>(N, b.touched, b.offsets, b.entry_grads, gradients);
        count_launch();
    }
    mark(tm, 5, st);
    return last_error();
}

// ---- what the classic entry points remember between launches, per (device, stream) -------------------------------------
struct SplatState {
    bool valid = false;
    bool external = false;       // buffers live in a caller-provided workspace (no generation to check)
    SplatView view{};
    SplatBuffers buf{};
    long long entries = 0;       // list length; -1: not known on the host yet (sync-free launch)
    long long known_entries = 0;  // most recent length the host has seen for this scene shape
    long long capacity = 0;      // sync-free launches: the length the entry part was sized for
    const unsigned long long* header = nullptr;
    long long stats[4] = {0, 0, 0, 0};
    int work_records = 0;        // size of buf.chunk_info (the backward grid)
    uint64_t gen_fixed = 0, gen_entry = 0;
    cudaStream_t stream = nullptr;
    // XYZ_FLAG_ASYNC: pinned-host mirror of {list length, overflow counter}, refreshed by every launch without a
    // synchronisation; read at the next call (whatever has arrived by then is a valid earlier value)
    volatile unsigned long long* mirror = nullptr;
    unsigned long long overflows_seen = 0;
    StageEvents timing;
};
std::mutex g_state_mu;
std::map<std::pair<int, cudaStream_t>, SplatState> g_states;  // node addresses are stable
thread_local SplatState* g_last = nullptr;                     // the state this host thread launched on last
thread_local SplatState g_last_ws;                             // ... or its last workspace launch

SplatState* state_for(cudaStream_t st) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(g_state_mu);
    return &g_states[{dev, st}];
}

bool state_usable(const SplatState* s) {
    if (!s || !s->valid) return false;
    if (s->external) return true;
    return scratch_is_current(SCRATCH_SPLAT, s->stream, s->gen_fixed) &&
           scratch_is_current(SCRATCH_SPLAT_SORT, s->stream, s->gen_entry);
}

void remember(SplatState& s, const SplatPlan& p, const Bound& o, long long entries, long long known, long long capacity) {
    s.valid = true;
    s.view = p.v;
    s.buf = o.b;
    s.entries = entries;
    s.known_entries = known;
    s.capacity = capacity;
    s.header = o.header;
    s.stats[0] = entries;
    s.stats[1] = p.n_tiles;
    s.stats[2] = -1;  // longest list: filled lazily by xyz_splat_last_stats
    s.stats[3] = entries * kTilePixels;
}

int splat_launch(const xyz_gaussian_params* gaussians, xyz_gaussian_grads* gradients, const float* target,
                 float* output, float* total_loss, int W, int H, int N, int row_begin, int row_end, void* stream,
                 int flags) {
    if (!target || !output || !total_loss || (N > 0 && (!gaussians || !gradients))) return XYZ_ERR_INVALID_ARGUMENT;
    if (W <= 0 || H <= 0 || N < 0 || row_begin < 0 || row_end > H || row_begin > row_end) return XYZ_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SplatState* sp = state_for(st);
    if (!sp) return XYZ_ERR_NOT_INITIALISED;
    SplatState& S = *sp;
    S.stream = st;
    const bool same_shape = state_usable(&S) && S.view.num_gaussians == N && S.view.width == W && S.view.height == H &&
                            S.view.row_begin == row_begin && S.view.row_end == row_end;
    if (!same_shape) S.valid = false;
    // XYZ_FLAG_ASYNC: what an earlier sync-free launch on this stream reported meanwhile (list length, overflows)
    if (S.mirror) {
        const unsigned long long seen_total = S.mirror[0], seen_overflows = S.mirror[1];
        if (same_shape && S.entries < 0 && seen_total != 0ull) S.known_entries = static_cast<long long>(seen_total);
        if (seen_overflows != S.overflows_seen) {
            S.overflows_seen = seen_overflows;
            S.capacity = 0;            // sized afresh from the reported length
            return XYZ_ERR_WORKSPACE;  // an earlier XYZ_FLAG_ASYNC launch was longer than its buffers: its outputs are void
        }
    }
    const long long predicted = (same_shape && S.known_entries > 0) ? S.known_entries : static_cast<long long>(N) * 40;
    SplatPlan p;
    int err = make_plan(p, W, H, N, row_begin, row_end, flags, predicted, st);
    if (err) return err;
    // sync-free launch: needs the counting sort (no host-side item counts) and a known length of this scene shape
    N = p.v.num_gaussians;  // 0 for an empty row band
    const bool async = (flags & XYZ_FLAG_ASYNC) && p.counting && N > 0 && same_shape && S.known_entries > 0;
    long long capacity = 0;
    if (async) {
        capacity = S.known_entries + S.known_entries / 2 + 4096;
        if (S.capacity >= S.known_entries + S.known_entries / 8) capacity = S.capacity;  // keep the buffers
        if (capacity >= (1LL << 31) - 512) return XYZ_ERR_WORKSPACE;
    }
    // the header {list length, overflow counter, ticket} comes FIRST in the arena: the overflow counter is sticky across
    // launches, so its place must not move with the scene shape (a fresh arena is zero-filled, common.cu)
    unsigned char* base = nullptr;
    uint64_t gen_fixed = 0, gen_entry = 0;
    err = scratch_get(SCRATCH_SPLAT, kHeaderBytes + p.fixed_bytes, reinterpret_cast<void**>(&base), st, &gen_fixed);
    if (err) return err;
    Bound o;
    bind_fixed(p, base, base + kHeaderBytes, o);
    const unsigned long long* total_src = nullptr;
    StageEvents* tm = nullptr;
    S.timing.armed = false;
    if (flags & XYZ_FLAG_TIMING) {
        if ((err = S.timing.ensure())) return err;
        tm = &S.timing;
    }
    err = enqueue_front(p, o, gaussians, static_cast<unsigned long long>(capacity), S.overflows_seen, st, &total_src, tm);
    if (err) return err;
    long long entries = 0;
    if (N > 0) {
        if (async) {
            // no read-back: the buffers are sized for `capacity`; the length goes to the pinned mirror for later calls
            if (!S.mirror) {
                void* m = nullptr;
                cudaError_t ce = cudaHostAlloc(&m, 2 * sizeof(unsigned long long), cudaHostAllocDefault);
                if (ce != cudaSuccess) return static_cast<int>(ce);
                S.mirror = static_cast<volatile unsigned long long*>(m);
                S.mirror[0] = 0ull;
                S.mirror[1] = S.overflows_seen;
            }
            cudaError_t ce = cudaMemcpyAsync(const_cast<unsigned long long*>(S.mirror), o.header,
                                             2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
            if (ce != cudaSuccess) return static_cast<int>(ce);
            entries = capacity;
        } else {
            // the list length is data dependent: one 8-byte read-back (the only synchronisation of this entry point;
            // xyz_launch_gaussian_splatting_ws has none)
            unsigned long long total = 0;
            cudaError_t ce = cudaMemcpyAsync(&total, total_src, sizeof(total), cudaMemcpyDeviceToHost, st);
            if (ce != cudaSuccess) return static_cast<int>(ce);
            ce = cudaStreamSynchronize(st);
            if (ce != cudaSuccess) return static_cast<int>(ce);
            entries = static_cast<long long>(total);
        }
    }
    if (entries >= (1LL << 31) - 512) return XYZ_ERR_WORKSPACE;
    const EntryLayout L = make_entry_layout(p, entries, st);
    unsigned char* sbase = nullptr;
    err = scratch_get(SCRATCH_SPLAT_SORT, L.bytes, reinterpret_cast<void**>(&sbase), st, &gen_entry);
    if (err) return err;
    bind_entry(p, L, sbase, o);
    err = enqueue_back(p, L, o, gaussians, gradients, target, output, total_loss, entries,
                       static_cast<unsigned long long>(capacity), st, tm);
    if (err) return err;
    S.timing.armed = tm != nullptr;
    const long long known_before = same_shape ? S.known_entries : 0;
    S.external = false;
    S.gen_fixed = gen_fixed;
    S.gen_entry = gen_entry;
    remember(S, p, o, async ? -1 : entries, async ? known_before : entries, async ? capacity : 0);
    S.work_records = entries > 0 ? L.chunk_info_size : 0;
    g_last = &S;
    return last_error();
}

// After a sync-free launch the host does not know the list length: fetch it (synchronises the stream).
// Returns XYZ_ERR_WORKSPACE if that launch was longer than its buffers (its outputs are void).
int resolve_async_length(SplatState& S) {
    if (!S.valid || S.entries >= 0) return 0;
    unsigned long long both[2] = {0ull, 0ull};
    cudaError_t e = cudaStreamSynchronize(S.stream);
    if (e != cudaSuccess) return static_cast<int>(e);
    e = cudaMemcpy(both, S.header, sizeof(both), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return static_cast<int>(e);
    S.known_entries = static_cast<long long>(both[0]);
    if (both[1] != S.overflows_seen) {
        S.overflows_seen = both[1];
        S.entries = 0;  // every list of that launch was published empty
        S.capacity = 0;
        S.stats[0] = 0;
        S.stats[3] = 0;
        return XYZ_ERR_WORKSPACE;
    }
    S.entries = static_cast<long long>(both[0]);
    S.stats[0] = S.entries;
    S.stats[3] = S.entries * kTilePixels;
    return 0;
}

int ws_plan(SplatPlan& p, EntryLayout& L, int W, int H, int N, int row_begin, int row_end, long long max_entries, int flags,
            cudaStream_t st) {
    if (max_entries < 0 || max_entries >= (1LL << 31) - 512) return XYZ_ERR_INVALID_ARGUMENT;
    if (flags & XYZ_FLAG_RADIX_BINNING) return XYZ_ERR_INVALID_ARGUMENT;
    int err = make_plan(p, W, H, N, row_begin, row_end, flags, max_entries, st);
    if (err) return err;
    if (!p.counting) return XYZ_ERR_INVALID_ARGUMENT;  // more than 57 344 tiles in the band: the radix path needs host-side counts
    L = make_entry_layout(p, max_entries, st);
    return 0;
}

}  // namespace
}  // namespace xyzb

extern "C" int xyz_launch_gaussian_splatting(const xyz_gaussian_params* gaussians, xyz_gaussian_grads* gradients,
                                             const float* target_image, float* output_image, float* total_loss,
                                             int image_width, int image_height, int num_gaussians, void* stream,
                                             int flags) {
    return xyzb::splat_launch(gaussians, gradients, target_image, output_image, total_loss, image_width, image_height,
                              num_gaussians, 0, image_height, stream, flags);
}

extern "C" int xyz_launch_gaussian_splatting_rows(const xyz_gaussian_params* gaussians, xyz_gaussian_grads* gradients,
                                                  const float* target_image, float* output_image, float* total_loss,
                                                  int image_width, int image_height, int num_gaussians, int row_begin,
                                                  int row_end, void* stream, int flags) {
    return xyzb::splat_launch(gaussians, gradients, target_image, output_image, total_loss, image_width, image_height,
                              num_gaussians, row_begin, row_end, stream, flags);
}

// ---- caller-provided workspace: no allocation, no synchronisation, no library state ----------------------------------
extern "C" size_t xyz_splat_workspace_bytes(int image_width, int image_height, int num_gaussians, int row_begin,
                                            int row_end, long long max_entries, int flags) {
    using namespace xyzb;
    SplatPlan p;
    EntryLayout L;
    if (ws_plan(p, L, image_width, image_height, num_gaussians, row_begin, row_end, max_entries, flags, nullptr)) return 0;
    return kHeaderBytes + align_up(p.fixed_bytes) + L.bytes;
}

extern "C" int xyz_splat_workspace_init(void* workspace, size_t workspace_bytes, void* stream) {
    if (!workspace || workspace_bytes < xyzb::kHeaderBytes) return XYZ_ERR_INVALID_ARGUMENT;
    return static_cast<int>(cudaMemsetAsync(workspace, 0, xyzb::kHeaderBytes, static_cast<cudaStream_t>(stream)));
}

extern "C" int xyz_launch_gaussian_splatting_ws(const xyz_gaussian_params* gaussians, xyz_gaussian_grads* gradients,
                                                const float* target_image, float* output_image, float* total_loss,
                                                int image_width, int image_height, int num_gaussians, int row_begin,
                                                int row_end, void* workspace, size_t workspace_bytes,
                                                long long max_entries, void* stream, int flags) {
    using namespace xyzb;
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return XYZ_ERR_INVALID_ARGUMENT;
    if (!target_image || !output_image || !total_loss || (num_gaussians > 0 && (!gaussians || !gradients)))
        return XYZ_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SplatPlan p;
    EntryLayout L;
    int err = ws_plan(p, L, image_width, image_height, num_gaussians, row_begin, row_end, max_entries, flags, st);
    if (err) return err;
    if (workspace_bytes < kHeaderBytes + align_up(p.fixed_bytes) + L.bytes) return XYZ_ERR_WORKSPACE;
    unsigned char* w = static_cast<unsigned char*>(workspace);
    Bound o;
    bind_fixed(p, w, w + kHeaderBytes, o);
    bind_entry(p, L, w + kHeaderBytes + align_up(p.fixed_bytes), o);
    const unsigned long long* total_src = nullptr;
    // capacity 0 would mean "unlimited" to the kernels: a workspace for zero entries still bounds the lists at 0 + 1
    const unsigned long long capacity = static_cast<unsigned long long>(max_entries > 0 ? max_entries : 1);
    num_gaussians = p.v.num_gaussians;  // 0 for an empty row band
    if (num_gaussians == 0) {           // no kernel will write the header: the list is empty
        cudaError_t ce = cudaMemsetAsync(o.header, 0, sizeof(unsigned long long), st);
        if (ce != cudaSuccess) return static_cast<int>(ce);
    }
    SplatState& S = g_last_ws;
    StageEvents* tm = nullptr;
    S.timing.armed = false;
    if (flags & XYZ_FLAG_TIMING) {
        if ((err = S.timing.ensure())) return err;
        tm = &S.timing;
    }
    err = enqueue_front(p, o, gaussians, capacity, 0ull, st, &total_src, tm);
    if (err) return err;
    err = enqueue_back(p, L, o, gaussians, gradients, target_image, output_image, total_loss,
                       num_gaussians > 0 ? static_cast<long long>(capacity) : 0, capacity, st, tm);
    if (err) return err;
    S.timing.armed = tm != nullptr;
    S.external = true;
    S.stream = st;
    remember(S, p, o, num_gaussians > 0 ? -1 : 0, 0, static_cast<long long>(capacity));
    S.work_records = num_gaussians > 0 ? L.chunk_info_size : 0;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    S.overflows_seen = ~0ull;  // unknown: xyz_splat_last_stats compares with the header's own previous value instead
    g_last = (cap == cudaStreamCaptureStatusNone) ? &S : nullptr;  // a captured launch has not run: nothing to report
    return last_error();
}

extern "C" int xyz_splat_workspace_status(const void* workspace, void* stream, long long status_host[4]) {
    if (!workspace || !status_host) return XYZ_ERR_INVALID_ARGUMENT;
    cudaError_t e = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return static_cast<int>(e);
    unsigned long long h[4] = {0, 0, 0, 0};
    e = cudaMemcpy(h, workspace, sizeof(h), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return static_cast<int>(e);
    status_host[0] = static_cast<long long>(h[0]);  // list length of the most recent launch
    status_host[1] = static_cast<long long>(h[1]);  // launches that did not fit (sticky)
    status_host[2] = static_cast<long long>(h[3]);  // max_entries of the most recent launch
    status_host[3] = h[3] != 0 && h[0] > h[3] ? 1 : 0;  // did the most recent launch overflow?
    return 0;
}

extern "C" int xyz_splat_last_stats(long long stats_host[4]) {
    using namespace xyzb;
    if (!stats_host || !g_last || !state_usable(g_last)) return XYZ_ERR_NOT_INITIALISED;
    SplatState& S = *g_last;
    if (S.external) {
        // workspace launch: read the header (synchronises the stream)
        long long st4[4];
        int err = xyz_splat_workspace_status(S.header, S.stream, st4);
        if (err) return err;
        if (st4[3]) return XYZ_ERR_WORKSPACE;
        S.entries = st4[0];
        S.stats[0] = S.entries;
        S.stats[3] = S.entries * kTilePixels;
    } else if (int err = resolve_async_length(S)) {
        return err;
    }
    if (S.stats[2] < 0) {
        const int n_tiles = S.view.tiles_x * S.view.tiles_y;
        const int ty0 = S.view.row_begin / kTile, ty1 = (S.view.row_end + kTile - 1) / kTile;
        std::vector<int2> r(n_tiles);
        cudaError_t e = cudaMemcpy(r.data(), S.buf.tile_ranges, sizeof(int2) * n_tiles, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) return static_cast<int>(e);
        long long mx = 0;
        for (int t = ty0 * S.view.tiles_x; t < ty1 * S.view.tiles_x; ++t) mx = std::max<long long>(mx, r[t].y - r[t].x);
        S.stats[2] = mx;
    }
    for (int i = 0; i < 4; ++i) stats_host[i] = S.stats[i];
    return 0;
}

extern "C" int xyz_splat_last_backward_stats(long long stats_host[3]) {
    using namespace xyzb;
    if (!stats_host) return XYZ_ERR_INVALID_ARGUMENT;
    long long st4[4];
    if (int err = xyz_splat_last_stats(st4)) return err;  // (also: the launch did not overflow)
    SplatState& S = *g_last;
    cudaError_t e = cudaStreamSynchronize(S.stream);
    if (e != cudaSuccess) return static_cast<int>(e);
    stats_host[0] = stats_host[1] = stats_host[2] = 0;
    if (S.work_records <= 0 || st4[0] <= 0) return 0;
    std::vector<int4> rec(static_cast<size_t>(S.work_records));
    e = cudaMemcpy(rec.data(), S.buf.chunk_info, sizeof(int4) * rec.size(), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return static_cast<int>(e);
    for (const int4& r : rec) {
        if (r.x < 0) continue;
        stats_host[0] += r.z;
        stats_host[1] += static_cast<long long>(r.z) * kTilePixels;
        stats_host[2] += 1;
    }
    return 0;
}

extern "C" int xyz_splat_debug_binning(int32_t* rects_host, int32_t* tile_ranges_host, int32_t* sorted_ids_host,
                                       float* records_host) {
    using namespace xyzb;
    if (!g_last || !state_usable(g_last)) return XYZ_ERR_NOT_INITIALISED;
    long long st4[4];
    if (int err = xyz_splat_last_stats(st4)) return err;
    SplatState& S = *g_last;
    cudaError_t e = cudaStreamSynchronize(S.stream);
    if (e != cudaSuccess) return static_cast<int>(e);
    const int n = S.view.num_gaussians, n_tiles = S.view.tiles_x * S.view.tiles_y;
    if (rects_host && n > 0) {
        e = cudaMemcpy(rects_host, S.buf.rects, sizeof(int4) * n, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) return static_cast<int>(e);
    }
    if (tile_ranges_host) {
        // tiles outside the launch's row band were not binned: reported as (0, 0)
        const int ty0 = S.view.row_begin / kTile, ty1 = (S.view.row_end + kTile - 1) / kTile;
        memset(tile_ranges_host, 0, sizeof(int2) * n_tiles);
        if (ty1 > ty0) {
            e = cudaMemcpy(tile_ranges_host + 2 * ty0 * S.view.tiles_x, S.buf.tile_ranges + ty0 * S.view.tiles_x,
                           sizeof(int2) * static_cast<size_t>(ty1 - ty0) * S.view.tiles_x, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) return static_cast<int>(e);
        }
    }
    if (sorted_ids_host && S.entries > 0) {
        e = cudaMemcpy(sorted_ids_host, S.buf.sorted_gid, sizeof(int) * S.entries, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) return static_cast<int>(e);
    }
    if (records_host && n > 0) {
        // the first 9 floats of every 16-float device record into the documented 12-float rows (9..11 stay untouched)
        e = cudaMemcpy2D(records_host, 12 * sizeof(float), S.buf.records, 16 * sizeof(float), 9 * sizeof(float), n,
                         cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) return static_cast<int>(e);
    }
    return 0;
}

extern "C" int xyz_splat_last_timing(float stage_us_host[6]) {
    using namespace xyzb;
    if (!stage_us_host || !g_last || !g_last->timing.armed) return XYZ_ERR_NOT_INITIALISED;
    SplatState& S = *g_last;
    cudaError_t e = cudaEventSynchronize(S.timing.ev[5]);
    if (e != cudaSuccess) return static_cast<int>(e);
    float total = 0.f;
    for (int i = 0; i < 5; ++i) {
        float ms = 0.f;
        e = cudaEventElapsedTime(&ms, S.timing.ev[i], S.timing.ev[i + 1]);
        if (e != cudaSuccess) return static_cast<int>(e);
        stage_us_host[i] = ms * 1000.0f;
    }
    e = cudaEventElapsedTime(&total, S.timing.ev[0], S.timing.ev[5]);
    if (e != cudaSuccess) return static_cast<int>(e);
    stage_us_host[5] = total * 1000.0f;
    return 0;
}
