// csrc/covproj_kernels.cu -- C3: batched covariance projection S' = (J W) S (J W)^T, forward +
// reverse, one element per thread, fp32.
//
// What it replaces: a per-thread graph of op::matmul<2,3,3>(J,W), op::matmul<2,3,3>(T,S),
// op::matmul<2,3,2>(U,T^T) nodes (reference include/xyz_autodiff/operations/binary/
// matmul_logic.cuh:33-81) over a packed symmetric 3x3 (symmetric_matrix_view.cuh:24-29) followed
// by node.backward().  In the reference every leaf adjoint term is one scalar atomicAdd
// (variable.cuh:48-50); here node values and adjoints live in registers and each output is
// written exactly once.
//
// Roofline: 192 algorithmic bytes per element (96 in: J6 W9 S6 g3; 96 out: out3 gJ6 gW9 gS6),
// ~330 flops -> HBM bound.  Data path (sm_100a):
//   HBM --TMA 1-D bulk copy (UBLKCP), 4 contiguous chunks per 128-element tile, mbarrier
//   complete_tx--> shared memory (3-stage ring) --conflict-free LDS (strides 6/9/6/3 words)-->
//   registers --> STS --> shared memory (2-stage ring) --TMA bulk store--> HBM.
// A tile of 128 elements is contiguous in every one of the 8 arrays, so no tensor map is needed
// and every byte is moved exactly once by full-line requests.  Persistent CTAs (a multiple of
// 148), static round-robin tile assignment.
#include "common.cuh"

namespace xyzb {
namespace {

constexpr int kTileE = 128;    // elements per tile == threads per CTA
constexpr int kInStages = 3;
constexpr int kOutStages = 2;
constexpr int kFloatsPerElem = 24;  // both directions
constexpr int kCtasPerSM = 3;

struct CovTile {  // one stage, input or output side: [J|gJ: 6][W|gW: 9][S|gS: 6][g|out: 3] x 128
    float a6[kTileE * 6];
    float b9[kTileE * 9];
    float c6[kTileE * 6];
    float d3[kTileE * 3];
};
static_assert(sizeof(CovTile) == kTileE * kFloatsPerElem * 4, "tile layout");

struct CovSmem {
    CovTile in[kInStages];
    CovTile out[kOutStages];
    uint64_t full[kInStages];
};

// Node values and adjoints of one element, all in registers.  Formula order follows the oracle
// (oracle/xyz_oracle.cpp::covproj_one): T = J W, U = T S, P = U T^T; G = [[g0,g1],[0,g2]];
// dU = G T, dT_b = G^T U, dT_a = dU S^T, dS = T^T dU, dJ = dT_a W^T + dT_b W^T, dW = J^T dT_a + J^T dT_b.
__device__ __forceinline__ void covproj_element(const float (&J)[6], const float (&W)[9], const float (&S)[6],
                                                const float (&g)[3], float (&out)[3], float (&gJ)[6], float (&gW)[9],
                                                float (&gS)[6]) {
    const float Sf[9] = {S[0], S[1], S[2], S[1], S[3], S[4], S[2], S[4], S[5]};
    float T[6], U[6];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) s += J[i * 3 + k] * W[k * 3 + j];
            T[i * 3 + j] = s;
        }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) s += T[i * 3 + k] * Sf[k * 3 + j];
            U[i * 3 + j] = s;
        }
    {
        float p00 = 0.f, p01 = 0.f, p11 = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            p00 += U[k] * T[k];
            p01 += U[k] * T[3 + k];
            p11 += U[3 + k] * T[3 + k];
        }
        out[0] = p00;
        out[1] = p01;
        out[2] = p11;
    }
    const float G[4] = {g[0], g[1], 0.f, g[2]};
    float dU[6], dTa[6], dTb[6];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 2; ++j) s += G[i * 2 + j] * T[j * 3 + k];
            dU[i * 3 + k] = s;
        }
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 2; ++i) s += G[i * 2 + j] * U[i * 3 + k];
            dTb[j * 3 + k] = s;
        }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) s += dU[i * 3 + j] * Sf[k * 3 + j];
            dTa[i * 3 + k] = s;
        }
    float dSf[9];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 2; ++i) s += T[i * 3 + k] * dU[i * 3 + j];
            dSf[k * 3 + j] = s;
        }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float sa = 0.f, sb = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                sa += dTa[i * 3 + j] * W[k * 3 + j];
                sb += dTb[i * 3 + j] * W[k * 3 + j];
            }
            gJ[i * 3 + k] = sa + sb;
        }
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float sa = 0.f, sb = 0.f;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                sa += J[i * 3 + k] * dTa[i * 3 + j];
                sb += J[i * 3 + k] * dTb[i * 3 + j];
            }
            gW[k * 3 + j] = sa + sb;
        }
    gS[0] = dSf[0];
    gS[1] = dSf[1] + dSf[3];
    gS[2] = dSf[2] + dSf[6];
    gS[3] = dSf[4];
    gS[4] = dSf[5] + dSf[7];
    gS[5] = dSf[8];
}

struct CovArgs {
    const float* J;
    const float* W;
    const float* S;
    const float* g;
    float* out;
    float* gJ;
    float* gW;
    float* gS;
};

__device__ __forceinline__ void issue_tile_load(CovTile* st, uint64_t* bar, const CovArgs& a, long long tile) {
    const long long e0 = tile * kTileE;
    mbar_arrive_expect_tx(bar, sizeof(CovTile));
    bulk_load(st->a6, a.J + e0 * 6, kTileE * 6 * 4, bar);
    bulk_load(st->b9, a.W + e0 * 9, kTileE * 9 * 4, bar);
    bulk_load(st->c6, a.S + e0 * 6, kTileE * 6 * 4, bar);
    bulk_load(st->d3, a.g + e0 * 3, kTileE * 3 * 4, bar);
}

// Full tiles only (n_tiles * 128 elements); 16-byte aligned bases.
__global__ void __launch_bounds__(kTileE, kCtasPerSM) covproj_tma_kernel(CovArgs a, long long n_tiles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    CovSmem& sm = *reinterpret_cast<CovSmem*>(smem_raw);
    const int tid = threadIdx.x;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kInStages; ++s) mbar_init(&sm.full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const long long first = blockIdx.x, stride = gridDim.x;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kInStages; ++s) {
            const long long t = first + s * stride;
            if (t < n_tiles) issue_tile_load(&sm.in[s], &sm.full[s], a, t);
        }
    }

    int it = 0;
    for (long long tile = first; tile < n_tiles; tile += stride, ++it) {
        const int s = it % kInStages;
        const uint32_t parity = (it / kInStages) & 1;
        mbar_wait(&sm.full[s], parity);

        float J[6], W[9], S[6], g[3];
        {
            const CovTile& in = sm.in[s];
            const float2* j2 = reinterpret_cast<const float2*>(in.a6 + tid * 6);
            const float2* s2 = reinterpret_cast<const float2*>(in.c6 + tid * 6);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float2 v = j2[k];
                J[2 * k] = v.x;
                J[2 * k + 1] = v.y;
                const float2 u = s2[k];
                S[2 * k] = u.x;
                S[2 * k + 1] = u.y;
            }
#pragma unroll
            for (int k = 0; k < 9; ++k) W[k] = in.b9[tid * 9 + k];
#pragma unroll
            for (int k = 0; k < 3; ++k) g[k] = in.d3[tid * 3 + k];
        }
        // the output stage we are about to fill was handed to the TMA unit two tiles ago
        if (tid == 0) bulk_wait_read<kOutStages - 1>();
        __syncthreads();  // stage s consumed by everyone; out stage free
        if (tid == 0) {
            const long long nt = tile + static_cast<long long>(kInStages) * stride;
            if (nt < n_tiles) issue_tile_load(&sm.in[s], &sm.full[s], a, nt);
        }

        float out[3], gJ[6], gW[9], gS[6];
        covproj_element(J, W, S, g, out, gJ, gW, gS);

        CovTile& o = sm.out[it % kOutStages];
        {
            float2* j2 = reinterpret_cast<float2*>(o.a6 + tid * 6);
            float2* s2 = reinterpret_cast<float2*>(o.c6 + tid * 6);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                j2[k] = make_float2(gJ[2 * k], gJ[2 * k + 1]);
                s2[k] = make_float2(gS[2 * k], gS[2 * k + 1]);
            }
#pragma unroll
            for (int k = 0; k < 9; ++k) o.b9[tid * 9 + k] = gW[k];
#pragma unroll
            for (int k = 0; k < 3; ++k) o.d3[tid * 3 + k] = out[k];
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            const long long e0 = tile * kTileE;
            bulk_store(a.gJ + e0 * 6, o.a6, kTileE * 6 * 4);
            bulk_store(a.gW + e0 * 9, o.b9, kTileE * 9 * 4);
            bulk_store(a.gS + e0 * 6, o.c6, kTileE * 6 * 4);
            bulk_store(a.out + e0 * 3, o.d3, kTileE * 3 * 4);
            bulk_commit();
        }
    }
    if (tid == 0) bulk_wait_all<0>();  // smem must outlive the last stores
}

// Tail / unaligned fallback: plain loads and stores, one element per thread.
__global__ void __launch_bounds__(128) covproj_plain_kernel(CovArgs a, long long e_begin, long long e_end) {
    const long long e = e_begin + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (e >= e_end) return;
    float J[6], W[9], S[6], g[3], out[3], gJ[6], gW[9], gS[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) J[k] = __ldg(a.J + e * 6 + k);
#pragma unroll
    for (int k = 0; k < 9; ++k) W[k] = __ldg(a.W + e * 9 + k);
#pragma unroll
    for (int k = 0; k < 6; ++k) S[k] = __ldg(a.S + e * 6 + k);
#pragma unroll
    for (int k = 0; k < 3; ++k) g[k] = __ldg(a.g + e * 3 + k);
    covproj_element(J, W, S, g, out, gJ, gW, gS);
#pragma unroll
    for (int k = 0; k < 6; ++k) a.gJ[e * 6 + k] = gJ[k];
#pragma unroll
    for (int k = 0; k < 9; ++k) a.gW[e * 9 + k] = gW[k];
#pragma unroll
    for (int k = 0; k < 6; ++k) a.gS[e * 6 + k] = gS[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) a.out[e * 3 + k] = out[k];
}


// ---- variant B: W is ONE shared 3x3 parameter -----------------------------------------------------------------
// The complete north-star pattern in one launch: per-element forward + reverse, and the per-element adjoints of the
// shared parameter accumulated into its 9 gradients (what E threads calling VariableRef::add_grad on the shared W
// do in the reference: 9 same-address atomics per element, variable.cuh:48-50).
// 120 algorithmic bytes per element (in J6 S6 g3, out out3 gJ6 gS6).  Same TMA ring as above with three arrays per
// direction; dW stays in 9 registers per thread for the whole persistent loop, then shuffle tree -> shared memory ->
// one row per CTA -> the last CTA (ticket) adds the rows in CTA order: no floating-point atomics, bit-identical run
// to run.  With a peer group the last CTA also exchanges the row with the other ranks over NVLink mailboxes.
constexpr int kSwInStages = 4;
constexpr int kSwOutStages = 2;
constexpr int kSwCtasPerSM = 4;
constexpr int kSwAcc = 9;

struct SwTile {  // [J|gJ: 6][S|gS: 6][g|out: 3] x 128
    float a6[kTileE * 6];
    float c6[kTileE * 6];
    float d3[kTileE * 3];
};
struct SwSmem {
    SwTile in[kSwInStages];
    SwTile out[kSwOutStages];
    uint64_t full[kSwInStages];
    float red[kTileE / 32][kSwAcc];
    int is_last;
};

__device__ __forceinline__ void issue_sw_tile_load(SwTile* st, uint64_t* bar, const CovArgs& a, long long tile) {
    const long long e0 = tile * kTileE;
    mbar_arrive_expect_tx(bar, sizeof(SwTile));
    bulk_load(st->a6, a.J + e0 * 6, kTileE * 6 * 4, bar);
    bulk_load(st->c6, a.S + e0 * 6, kTileE * 6 * 4, bar);
    bulk_load(st->d3, a.g + e0 * 3, kTileE * 3 * 4, bar);
}

// a.W: the shared 3x3 (9 floats), a.gW: its 9 gradient accumulators (+=).  partials: [gridDim.x][9]; ticket: zero before
// the launch, reset by the last CTA.
template <bool kTma>
__global__ void __launch_bounds__(kTileE, kSwCtasPerSM)
    covproj_sharedw_kernel(CovArgs a, long long n, float* partials, unsigned int* ticket, PeerArgs peer) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SwSmem& sm = *reinterpret_cast<SwSmem*>(smem_raw);
    const int tid = threadIdx.x;
    float W[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) W[k] = __ldg(a.W + k);
    float acc[kSwAcc];
#pragma unroll
    for (int k = 0; k < kSwAcc; ++k) acc[k] = 0.f;

    const long long n_tiles = kTma ? n / kTileE : 0;
    if constexpr (kTma) {
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < kSwInStages; ++s) mbar_init(&sm.full[s], 1);
            mbar_fence_init();
        }
        __syncthreads();
        const long long first = blockIdx.x, stride = gridDim.x;
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < kSwInStages; ++s) {
                const long long t = first + s * stride;
                if (t < n_tiles) issue_sw_tile_load(&sm.in[s], &sm.full[s], a, t);
            }
        }
        int it = 0;
        for (long long tile = first; tile < n_tiles; tile += stride, ++it) {
            const int s = it % kSwInStages;
            mbar_wait(&sm.full[s], (it / kSwInStages) & 1);
            float J[6], S[6], g[3];
            {
                const SwTile& in = sm.in[s];
                const float2* j2 = reinterpret_cast<const float2*>(in.a6 + tid * 6);
                const float2* s2 = reinterpret_cast<const float2*>(in.c6 + tid * 6);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float2 v = j2[k];
                    J[2 * k] = v.x;
                    J[2 * k + 1] = v.y;
                    const float2 u = s2[k];
                    S[2 * k] = u.x;
                    S[2 * k + 1] = u.y;
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) g[k] = in.d3[tid * 3 + k];
            }
            if (tid == 0) bulk_wait_read<kSwOutStages - 1>();
            __syncthreads();
            if (tid == 0) {
                const long long nt = tile + static_cast<long long>(kSwInStages) * stride;
                if (nt < n_tiles) issue_sw_tile_load(&sm.in[s], &sm.full[s], a, nt);
            }
            float out[3], gJ[6], gW[9], gS[6];
            covproj_element(J, W, S, g, out, gJ, gW, gS);
#pragma unroll
            for (int k = 0; k < kSwAcc; ++k) acc[k] += gW[k];
            SwTile& o = sm.out[it % kSwOutStages];
            {
                float2* j2 = reinterpret_cast<float2*>(o.a6 + tid * 6);
                float2* s2 = reinterpret_cast<float2*>(o.c6 + tid * 6);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    j2[k] = make_float2(gJ[2 * k], gJ[2 * k + 1]);
                    s2[k] = make_float2(gS[2 * k], gS[2 * k + 1]);
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) o.d3[tid * 3 + k] = out[k];
            }
            fence_proxy_async_smem();
            __syncthreads();
            if (tid == 0) {
                const long long e0 = tile * kTileE;
                bulk_store(a.gJ + e0 * 6, o.a6, kTileE * 6 * 4);
                bulk_store(a.gS + e0 * 6, o.c6, kTileE * 6 * 4);
                bulk_store(a.out + e0 * 3, o.d3, kTileE * 3 * 4);
                bulk_commit();
            }
        }
        if (tid == 0) bulk_wait_all<0>();
    }
    // tail (and the whole range when a base pointer is not 16-byte aligned): plain loads and stores
    for (long long e = n_tiles * kTileE + blockIdx.x * static_cast<long long>(kTileE) + tid; e < n;
         e += static_cast<long long>(gridDim.x) * kTileE) {
        float J[6], S[6], g[3], out[3], gJ[6], gW[9], gS[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) J[k] = __ldg(a.J + e * 6 + k);
#pragma unroll
        for (int k = 0; k < 6; ++k) S[k] = __ldg(a.S + e * 6 + k);
#pragma unroll
        for (int k = 0; k < 3; ++k) g[k] = __ldg(a.g + e * 3 + k);
        covproj_element(J, W, S, g, out, gJ, gW, gS);
#pragma unroll
        for (int k = 0; k < kSwAcc; ++k) acc[k] += gW[k];
#pragma unroll
        for (int k = 0; k < 6; ++k) a.gJ[e * 6 + k] = gJ[k];
#pragma unroll
        for (int k = 0; k < 6; ++k) a.gS[e * 6 + k] = gS[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) a.out[e * 3 + k] = out[k];
    }

    // CTA reduction (fixed order), one row per CTA, last CTA adds the rows in CTA order
#pragma unroll
    for (int k = 0; k < kSwAcc; ++k) acc[k] = warp_sum(acc[k]);
    if ((tid & 31) == 0) {
#pragma unroll
        for (int k = 0; k < kSwAcc; ++k) sm.red[tid >> 5][k] = acc[k];
    }
    __syncthreads();
    if (tid < kSwAcc) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kTileE / 32; ++w) s += sm.red[w][tid];
        partials[static_cast<size_t>(blockIdx.x) * kSwAcc + tid] = s;
        __threadfence();
    }
    __syncthreads();
    if (tid == 0) sm.is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!sm.is_last) return;
    __threadfence();
    {
        float s[kSwAcc];
#pragma unroll
        for (int k = 0; k < kSwAcc; ++k) s[k] = 0.f;
        for (unsigned int r = tid; r < gridDim.x; r += kTileE) {
#pragma unroll
            for (int k = 0; k < kSwAcc; ++k) s[k] += __ldcg(partials + static_cast<size_t>(r) * kSwAcc + k);
        }
#pragma unroll
        for (int k = 0; k < kSwAcc; ++k) s[k] = warp_sum(s[k]);
        __syncthreads();
        if ((tid & 31) == 0) {
#pragma unroll
            for (int k = 0; k < kSwAcc; ++k) sm.red[tid >> 5][k] = s[k];
        }
        __syncthreads();
        float t = 0.f;
        if (tid < kSwAcc) {
#pragma unroll
            for (int w = 0; w < kTileE / 32; ++w) t += sm.red[w][tid];
        }
        if (peer.world > 1) {  // rank-ordered sum over NVLink mailboxes (bit-identical on every rank)
            const int par = static_cast<int>(peer.seq & 1ull);
            if (tid < kSwAcc) {
                for (int p = 0; p < peer.world; ++p) peer.box[p]->vec[par][peer.rank][tid] = t;
                __threadfence_system();
            }
            const bool ok = peer_publish_and_wait_cta(peer, tid);
            if (tid < kSwAcc) {
                t = 0.f;
                for (int q = 0; q < peer.world; ++q) t += ld_relaxed_sys_f32(&peer.box[peer.rank]->vec[par][q][tid]);
                if (!ok) t = __int_as_float(0x7fc00000);
            }
        }
        if (tid < kSwAcc) a.gW[tid] += t;
        if (tid == 0) *ticket = 0u;
    }
}

int covproj_sharedw_launch(const float* J, const float* W9, const float* S, const float* g, float* out, float* gJ,
                           float* gW9, float* gS, long long n, const PeerArgs& peer, void* stream) {
    if (n < 0 || !W9 || !gW9) return XYZ_ERR_INVALID_ARGUMENT;
    if (n == 0 && peer.world <= 1) return 0;
    if (n > 0 && (!J || !S || !g || !out || !gJ || !gS)) return XYZ_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CovArgs a{J, W9, S, g, out, gJ, gW9, gS};
    const long long max_ctas = static_cast<long long>(sm_count()) * kSwCtasPerSM;
    const long long want = (n + kTileE - 1) / kTileE;
    const int grid = static_cast<int>(want < 1 ? 1 : (want < max_ctas ? want : max_ctas));
    void* scratch = nullptr;
    const int err = scratch_get(SCRATCH_REDUCE, 256 + static_cast<size_t>(max_ctas) * kSwAcc * sizeof(float), &scratch, st);
    if (err) return err;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch);
    float* partials = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(scratch) + 256);
    const bool tma_ok = n >= kTileE && aligned16(J) && aligned16(S) && aligned16(g) && aligned16(out) && aligned16(gJ) &&
                        aligned16(gS);
    if (tma_ok) {
        cudaFuncSetAttribute(covproj_sharedw_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(sizeof(SwSmem)));
        covproj_sharedw_kernel<true><<<grid, kTileE, sizeof(SwSmem), st>>>(a, n, partials, ticket, peer);
    } else {
        covproj_sharedw_kernel<false><<<grid, kTileE, sizeof(SwSmem), st>>>(a, n, partials, ticket, peer);
    }
    count_launch();
    return last_error();
}

}  // namespace
}  // namespace xyzb

extern "C" int xyz_covproj_fwd_bwd_f32(const float* J, const float* W, const float* S, const float* g, float* out,
                                       float* gJ, float* gW, float* gS, long long n, void* stream, int flags) {
    using namespace xyzb;
    (void)flags;
    if (n < 0) return XYZ_ERR_INVALID_ARGUMENT;
    if (n == 0) return 0;
    if (!J || !W || !S || !g || !out || !gJ || !gW || !gS) return XYZ_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CovArgs a{J, W, S, g, out, gJ, gW, gS};
    const bool tma_ok = aligned16(J) && aligned16(W) && aligned16(S) && aligned16(g) && aligned16(out) &&
                        aligned16(gJ) && aligned16(gW) && aligned16(gS);
    long long n_tiles = tma_ok ? n / kTileE : 0;
    if (n_tiles > 0) {
        static bool attr_set = false;  // per process; harmless to repeat
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(covproj_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 static_cast<int>(sizeof(CovSmem)));
            if (e != cudaSuccess) return static_cast<int>(e);
            attr_set = true;
        }
        const long long max_ctas = static_cast<long long>(sm_count()) * kCtasPerSM;
        const int grid = static_cast<int>(n_tiles < max_ctas ? n_tiles : max_ctas);
        covproj_tma_kernel<<<grid, kTileE, sizeof(CovSmem), st>>>(a, n_tiles);
        count_launch();
    }
    const long long done = n_tiles * kTileE;
    if (done < n) {
        const long long rest = n - done;
        const int grid = static_cast<int>((rest + 127) / 128);
        covproj_plain_kernel<<<grid, 128, 0, st>>>(a, done, n);
        count_launch();
    }
    return last_error();
}

extern "C" int xyz_covproj_shared_w_fwd_bwd_f32(const float* J, const float* W9, const float* S, const float* g, float* out,
                                                float* gJ, float* gW9, float* gS, long long n, void* stream, int flags) {
    (void)flags;
    xyzb::PeerArgs pa{};
    return xyzb::covproj_sharedw_launch(J, W9, S, g, out, gJ, gW9, gS, n, pa, stream);
}

extern "C" int xyz_covproj_shared_w_fwd_bwd_f32_allreduce(const float* J, const float* W9, const float* S, const float* g,
                                                          float* out, float* gJ, float* gW9, float* gS, long long n,
                                                          const xyz_peer_group* group, unsigned long long seq, void* stream,
                                                          int flags) {
    using namespace xyzb;
    (void)flags;
    if (!group || group->world < 1 || group->world > XYZ_PEER_MAX_WORLD || group->rank < 0 || group->rank >= group->world ||
        seq == 0)
        return XYZ_ERR_INVALID_ARGUMENT;
    PeerArgs pa{};
    pa.rank = group->rank;
    pa.world = group->world;
    pa.seq = seq;
    for (int i = 0; i < group->world; ++i) {
        if (!group->mailbox[i]) return XYZ_ERR_INVALID_ARGUMENT;
        pa.box[i] = static_cast<PeerMailbox*>(group->mailbox[i]);
    }
    return covproj_sharedw_launch(J, W9, S, g, out, gJ, gW9, gS, n, pa, stream);
}
