// csrc/splat_common.cuh -- shared definitions of the splat pipeline (C4/C5).
//
// The reference kernel (examples/mini-gaussian-splatting/gaussian_splatting_kernel.cu:8-112)
// evaluates EVERY (pixel, Gaussian) pair twice and issues 9 scalar atomicAdds per pair.  Its cull
// (:48-50, :90-92) is dead code.  This pipeline produces the same image, loss and
// gradients from
//   1. splat_preprocess_kernel : per Gaussian, once: exp(scale), R(theta), Sigma, Sigma^-1,
//      sigmoid(opacity) -> a 64-byte record; plus the rectangle of 16x16 tiles outside which the
//      pair weight exp(-d2/2) is EXACTLY 0.0f (so skipping those pairs cannot change any result);
//   2. integer work: per-tile lists of Gaussian ids, each ascending (= the reference's summation
//      order), built by a stable counting sort by tile (splat_host.cu section 2b) or, for more than
//      57 344 tiles per row band, (tile, Gaussian) keys + a stable radix sort; per-tile [begin, end) ranges;
//   3. splat_forward_kernel  : one CTA per tile (longest list first), four pixels per thread, records
//      staged in shared memory; writes the image, one loss partial per half tile, the residuals for the
//      backward pass -- and lists the backward work items: the tile's entries that reach a weight of
//      exp(-24) somewhere on the tile (kD2Backward below; everything else is left out of the gradient sums);
//   4. splat_backward_kernel : one THREAD per work item looping over the tile's 256 pixels (pixel
//      residuals broadcast from shared memory), 9 adjoint sums in registers, the per-Gaussian chain
//      rule applied once per item, then 9 REDs (or a partial row in deterministic mode) -- instead
//      of 9 atomics per PAIR.
#pragma once

#include "common.cuh"

namespace xyzb {

constexpr int kTile = 16;                 // TILE_SIZE, gaussian_splatting_kernel.cuh:21
constexpr int kTilePixels = kTile * kTile;
#ifndef XYZ_BWD_CHUNK
#define XYZ_BWD_CHUNK 128
#endif
constexpr int kBwdChunk = XYZ_BWD_CHUNK;   // work items (= threads) per backward CTA; chunks never straddle tiles
constexpr int kSpanRows = 16;              // tile-row spans kept per Gaussian between preprocess and key emission
constexpr int kRecFloats = 16;            // {cx, cy, ia, ib | ic, sigmoid(opacity), r, g | b, exp(s0), exp(s1), cos | sin, 0, 0, 0}

// exp(-d2/2) == 0.0f exactly beyond these:
//   fast-math/FTZ (ex2.approx.ftz): 0.5*d2*log2(e) > 126   <=> 0.5*d2 > 87.34 ; margin -> 88
//   IEEE expf:                      0.5*d2 > 103.98 (below half the smallest denormal) ; margin -> 104.5
constexpr float kD2MaxFast = 176.0f;
constexpr float kD2MaxPrecise = 209.0f;
// exp(-d2 / 2) as one special-function op on a pre-scaled conic: fast flavour ex2(kappa d2) with kappa = -0.5 log2(e)
// (what -use_fast_math makes of expf), IEEE flavour expf(kappa d2) with kappa = -0.5
constexpr float kKappaFast = -0.72134752044448170368f;
constexpr float kKappaPrecise = -0.5f;
constexpr float kD2MaxTail = 56.0f;  // XYZ_FLAG_TAIL_CULL (opt-in, bounded error): weights below exp(-28)
// Backward cull: list entries with min d2 over their tile beyond this are left out of the gradient sums -- every term
// of such a pair carries exp(-d2 / 2) < exp(-24) = 3.8e-11 (see splat_kernels.cuh; XYZ_FLAG_BWD_ALL_PAIRS turns it off).
// The bound is chosen from a measurement of what the cull changes, in deterministic mode where every kept entry is
// computed by the same instructions with and without it (dev/bwd_cull_sweep.py, profiles/bwd_cull_sweep_r02.log, C4 scene,
// largest change of an fp32 gradient sum relative to the fp64 sum of |terms|, 66 sampled Gaussians): D = 64: 3e-15,
// 56: 1e-13, 48: 1.1e-11, 40: 1.2e-7, 32: 5e-7.  In exact arithmetic the dropped tail of the largest coefficient
// (d2 exp(-d2 / 2)) beyond 48 is 1e-9 of its total: 1/60 of an fp32 epsilon -- it can move the rounding of a sum by its last
// bit and no further (tests/test_gpu_parity.py); beyond 40 whole ulps go.  The parity bar for accumulated sums is 1e-4.
// Backward time at C4 for D = 64 / 56 / 48 / 40 / 32: 324 / 301 / 273 / 244 / 223 us (measured with half-tile items; whole-tile
// items, the final form, take the same 275 us at 48).
constexpr float kD2Backward = 48.0f;

struct SplatView {  // what one launch renders
    int width, height, num_gaussians;
    int row_begin, row_end;  // pixel rows [row_begin, row_end)
    int tiles_x, tiles_y;    // full image, in tiles
};

struct SplatBuffers {  // device scratch of one launch (library-owned)
    float4* records;          // N x 4 float4 (kRecFloats)
    float4* fwd_records;      // N x 2 float4: {cx, cy, kappa ia, 2 kappa ib} {kappa ic, so r, so g, so b} (forward staging)
    int4* rects;              // N: tx0, ty0, tx1, ty1 (half-open)
    unsigned int* touched;    // N: tiles per Gaussian
    int2* spans;              // N x kSpanRows: [tx0, tx1) of the first kSpanRows tile rows of the rectangle
    unsigned long long* offsets;  // N: inclusive scan of touched (64-bit: N x tiles can exceed 2^32)
    unsigned int* keys_in;    // entries: tile id, Gaussian order
    unsigned int* keys_out;   // entries: sorted
    unsigned int* vals_in;    // entries: entry index in Gaussian order (== position)
    unsigned int* vals_out;   // entries: sorted -> original entry index
    int* sorted_gid;          // entries: Gaussian id per sorted entry (fast mode: aliases vals_out)
    int2* tile_ranges;        // tiles: [begin, end)
    int* chunk_offsets;       // tiles of the launch + 1: exclusive scan of the backward work records set aside per tile
                              // (ceil(list length / kBwdChunk))
    int4* chunk_info;         // backward work records = CTAs (entries / kBwdChunk + tiles, an upper bound):
                              // {tile or -1, first item's slot, items, -}, written by the forward pass
    float4* rest_tiles;       // tiles x 256: (target - output, active) per pixel, tile-major, written by the forward pass
    float* tile_loss;         // 2 x tiles: one partial per half tile (rows 0..7, rows 8..15)
    float* entry_grads;       // deterministic mode: entries x 9 (indexed by ORIGINAL entry index)
    int* tile_order;          // tiles of the launch: band-local tile ids, longest list first (launch order of the forward CTAs)
    int* bwd_items;           // entries: the backward work items, written by the forward pass into the first slots of
                              // every tile's range (see splat_kernels.cuh)
};

// flavour launchers (splat_fast.cu is built with -use_fast_math like the reference's training app,
// splat_precise.cu without, like the reference's tests)
// ticket: one unsigned int that is zero before the launch (the last CTA resets it): the forward pass itself adds the
// per-tile loss partials, in tile order, to *total_loss
// d2_bwd: the backward cull's bound on d2 (infinity: every listed pair becomes a backward work item);
// first_tile: the tile chunk_offsets[0] belongs to (the first tile of the row band; 0 on the radix path);
// tile_order: launch order of the tiles of the band (ids relative to first_tile), or nullptr = row-major
int splat_forward_launch_fast(const SplatView&, const SplatBuffers&, const float* target, float* output, float* total_loss,
                              unsigned int* ticket, bool deterministic, float d2_bwd, int first_tile, const int* tile_order, cudaStream_t);
int splat_forward_launch_precise(const SplatView&, const SplatBuffers&, const float* target, float* output,
                                 float* total_loss, unsigned int* ticket, bool deterministic, float d2_bwd, int first_tile,
                                 const int* tile_order, cudaStream_t);
// bwd_ctas = size of the backward work list (an upper bound of the CTAs needed; surplus records hold tile = -1)
int splat_backward_launch_fast(const SplatView&, const SplatBuffers&, xyz_gaussian_grads* grads, long long bwd_ctas,
                               bool deterministic, cudaStream_t);
int splat_backward_launch_precise(const SplatView&, const SplatBuffers&, xyz_gaussian_grads* grads, long long bwd_ctas,
                                  bool deterministic, cudaStream_t);

}  // namespace xyzb
