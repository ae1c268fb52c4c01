// csrc/lsq_kernels.cu -- C1: batched least-squares forward + reverse in fp64 with the four
// shared-parameter adjoints reduced on chip.
//
// What it replaces (reference): compute_gradient_kernel<Analytical>
// (examples/optimization/tests/test_linear_regression_gradient.cu:33-79) and
// parallel_gradient_computation_kernel (examples/optimization/linear_regression_sgd.cu:86-123):
// one thread per DataPoint builds the 9-node graph r = (a-x1)^2 + b (c-x2)^2 + d - y, loss = r^2,
// calls run(), and every thread issues 4 same-address fp64 atomicAdds (variable.cuh:48-50).
//
// Here: node values and adjoints stay in registers, each thread folds many points into 5
// private fp64 sums, a warp-shuffle tree + one shared-memory step reduce the CTA, each CTA
// writes ONE partial row, and the last CTA to finish (ticket) adds the rows in CTA order into
// params->grad.  No floating-point atomics at all, so the result is bit-identical run to run
// (XYZ_FLAG_DETERMINISTIC is always honoured).  24 algorithmic bytes per point -> HBM bound;
// points arrive by TMA 1-D bulk copies (256 points = 6 KB contiguous per tile, 4-stage ring).
#include <cooperative_groups.h>

#include "common.cuh"

namespace xyzb {
namespace {

constexpr int kThreads = 256;
constexpr int kTileP = 256;  // points per tile
#ifndef XYZ_LSQ_STAGES
#define XYZ_LSQ_STAGES 4
#endif
#ifndef XYZ_LSQ_CTAS_PER_SM
#define XYZ_LSQ_CTAS_PER_SM 4
#endif
constexpr int kStages = XYZ_LSQ_STAGES;
constexpr int kCtasPerSM = XYZ_LSQ_CTAS_PER_SM;
constexpr int kAcc = 5;  // ga gb gc gd loss

struct LsqSmem {
    double pts[kStages][kTileP * 3];
    uint64_t full[kStages];
    double red[kThreads / 32][kAcc];
    int is_last;
};

struct LsqAcc {
    double v[kAcc];
};

inline int lsq_mode(int flags) {
    return (flags & XYZ_FLAG_LSQ_SHIPPED_GRAPH) ? 2 : ((flags & XYZ_FLAG_RESIDUAL_ONLY) ? 1 : 0);
}

// mode 0: the graph of the reference's gradient test (test_linear_regression_gradient.cu:52-71), root = squared residual
// mode 1: the same graph differentiated at the residual (XYZ_FLAG_RESIDUAL_ONLY)
// mode 2: the graph the SHIPPED example builds (linear_regression_sgd.cu:103-122, XYZ_FLAG_LSQ_SHIPPED_GRAPH):
//         combined_terms = x1_term + x2_term takes the UN-squared x1_term = a - x1 (x1_term2 is dead) and the root is
//         the residual: r = (a - x1) + b (c - x2)^2 + d - y, dr/da = 1
__device__ __forceinline__ void lsq_point(double x1, double x2, double yt, double a, double b, double c, double d,
                                          int mode, LsqAcc& acc) {
    const double u = a - x1;          // sub_constant(a, x1)
    const double v = c - x2;          // sub_constant(c, x2)
    const double v2 = v * v;          // squared
    const double first = mode == 2 ? u : u * u;
    const double r = ((first + b * v2) + d) - yt;
    const double seed = mode ? 1.0 : 2.0 * r;  // squared backward: g * 2.0 * x with g = 1
    acc.v[0] += mode == 2 ? seed : seed * 2.0 * u;
    acc.v[1] += seed * v2;
    acc.v[2] += seed * b * 2.0 * v;
    acc.v[3] += seed;
    acc.v[4] += mode ? r : r * r;
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// partials: [gridDim.x][kAcc]; ticket: one unsigned int, zero before the launch, reset by the last CTA.
template <bool kTma>
__global__ void __launch_bounds__(kThreads, kCtasPerSM)
    lsq_grad_kernel(const double* __restrict__ data, long long n_points, xyz_lsq_parameters* params, double* loss_sum,
                    double* partials, unsigned int* ticket, int mode, PeerArgs peer) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    LsqSmem& sm = *reinterpret_cast<LsqSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const double a = params->value[0], b = params->value[1], c = params->value[2], d = params->value[3];
    LsqAcc acc;
#pragma unroll
    for (int k = 0; k < kAcc; ++k) acc.v[k] = 0.0;

    const long long n_tiles = n_points / kTileP;
    const long long first = blockIdx.x, stride = gridDim.x;
    if constexpr (kTma) {
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < kStages; ++s) mbar_init(&sm.full[s], 1);
            mbar_fence_init();
        }
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < kStages; ++s) {
                const long long t = first + s * stride;
                if (t < n_tiles) {
                    mbar_arrive_expect_tx(&sm.full[s], kTileP * 24);
                    bulk_load(sm.pts[s], data + t * kTileP * 3, kTileP * 24, &sm.full[s]);
                }
            }
        }
        int it = 0;
        for (long long tile = first; tile < n_tiles; tile += stride, ++it) {
            const int s = it % kStages;
            mbar_wait(&sm.full[s], (it / kStages) & 1);
            const double x1 = sm.pts[s][tid * 3], x2 = sm.pts[s][tid * 3 + 1], yt = sm.pts[s][tid * 3 + 2];
            __syncthreads();  // stage consumed
            if (tid == 0) {
                const long long nt = tile + static_cast<long long>(kStages) * stride;
                if (nt < n_tiles) {
                    mbar_arrive_expect_tx(&sm.full[s], kTileP * 24);
                    bulk_load(sm.pts[s], data + nt * kTileP * 3, kTileP * 24, &sm.full[s]);
                }
            }
            lsq_point(x1, x2, yt, a, b, c, d, mode, acc);
        }
    }
    // tail (and the whole range when the base pointer is not 16-byte aligned): plain loads
    {
        const long long begin = kTma ? n_tiles * kTileP : 0;
        for (long long i = begin + blockIdx.x * static_cast<long long>(kThreads) + tid; i < n_points;
             i += static_cast<long long>(gridDim.x) * kThreads) {
            lsq_point(__ldg(data + 3 * i), __ldg(data + 3 * i + 1), __ldg(data + 3 * i + 2), a, b, c, d,
                      mode, acc);
        }
    }

    // CTA reduction: shuffle tree per warp, then 8 rows in shared memory, fixed order
#pragma unroll
    for (int k = 0; k < kAcc; ++k) acc.v[k] = warp_sum(acc.v[k]);
    if ((tid & 31) == 0) {
#pragma unroll
        for (int k = 0; k < kAcc; ++k) sm.red[tid >> 5][k] = acc.v[k];
    }
    __syncthreads();
    if (tid < kAcc) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) s += sm.red[w][tid];
        partials[static_cast<size_t>(blockIdx.x) * kAcc + tid] = s;
        __threadfence();
    }
    __syncthreads();
    if (tid == 0) {
        const unsigned int done = atomicAdd(ticket, 1u);
        sm.is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (!sm.is_last) return;
    __threadfence();
    // last CTA: thread t sums rows t, t+256, ... (independent loads, all in flight together), then the same
    // fixed-order shuffle/shared-memory tree as above -> bit-identical for a given grid
    {
        double s[kAcc];
#pragma unroll
        for (int k = 0; k < kAcc; ++k) s[k] = 0.0;
        for (unsigned int r = tid; r < gridDim.x; r += kThreads) {
#pragma unroll
            for (int k = 0; k < kAcc; ++k) s[k] += __ldcg(partials + static_cast<size_t>(r) * kAcc + k);
        }
#pragma unroll
        for (int k = 0; k < kAcc; ++k) s[k] = warp_sum(s[k]);
        __syncthreads();  // sm.red is reused
        if ((tid & 31) == 0) {
#pragma unroll
            for (int k = 0; k < kAcc; ++k) sm.red[tid >> 5][k] = s[k];
        }
        __syncthreads();
        double t = 0.0;
        if (tid < kAcc) {
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) t += sm.red[w][tid];
        }
        // multi-GPU: exchange the row with the other ranks over NVLink peer memory, rank-ordered sum (common.cuh)
        if (peer.world > 1) t = peer_allreduce_cta(peer, t, kAcc, tid);
        if (tid < kAcc) {
            if (tid < 4) params->grad[tid] += t;
            else if (loss_sum) *loss_sum += t;
        }
        if (tid == 0) *ticket = 0u;
    }
}

__global__ void lsq_sgd_update_kernel(xyz_lsq_parameters* p, double lr, double batch) {
    const int i = threadIdx.x;
    if (i < 4) p->value[i] -= lr * p->grad[i] / batch;
}

__global__ void lsq_select_batch_kernel(const xyz_data_point* __restrict__ data, long long n_total,
                                        xyz_data_point* __restrict__ batch, long long batch_size, uint64_t seed,
                                        uint64_t epoch) {
    const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (i >= batch_size) return;
    const uint64_t h = splitmix64(splitmix64(seed ^ (epoch * 0xD1B54A32D192ED03ull)) + static_cast<uint64_t>(i));
    batch[i] = data[h % static_cast<uint64_t>(n_total)];
}


// ---- one SGD epoch in ONE launch ---------------------------------------------------------------------------
// The reference's epoch body (linear_regression_sgd.cu:185-209) is four operations on a 196 KB batch: gather,
// clear the gradients, gradient kernel, parameter update -- launch bound.  Here thread i of the batch gathers
// its own sample straight from the (L2-resident) data set with the same counter-based hash as
// lsq_select_batch_kernel, runs the graph, the CTA reduces as above, and the last CTA to finish adds the rows in
// CTA order, STORES the batch gradient (what cudaMemset + accumulation leaves behind) and applies
// value -= lr * grad / batch.  Deterministic; identical sampling to xyz_lsq_select_batch.
__global__ void __launch_bounds__(kThreads)
    lsq_sgd_step_kernel(const xyz_data_point* __restrict__ data, long long n_total, xyz_lsq_parameters* params,
                        long long batch_size, uint64_t seed, uint64_t epoch, double lr, double* loss_sum,
                        double* partials, unsigned int* ticket, int mode) {
    __shared__ double red[kThreads / 32][kAcc];
    __shared__ int is_last;
    const int tid = threadIdx.x;
    const double a = params->value[0], b = params->value[1], c = params->value[2], d = params->value[3];
    LsqAcc acc;
#pragma unroll
    for (int k = 0; k < kAcc; ++k) acc.v[k] = 0.0;
    const uint64_t base = splitmix64(seed ^ (epoch * 0xD1B54A32D192ED03ull));
    for (long long i = blockIdx.x * static_cast<long long>(kThreads) + tid; i < batch_size;
         i += static_cast<long long>(gridDim.x) * kThreads) {
        const uint64_t h = splitmix64(base + static_cast<uint64_t>(i));
        const double* p = reinterpret_cast<const double*>(data + h % static_cast<uint64_t>(n_total));
        lsq_point(__ldg(p), __ldg(p + 1), __ldg(p + 2), a, b, c, d, mode, acc);
    }
#pragma unroll
    for (int k = 0; k < kAcc; ++k) acc.v[k] = warp_sum(acc.v[k]);
    if ((tid & 31) == 0) {
#pragma unroll
        for (int k = 0; k < kAcc; ++k) red[tid >> 5][k] = acc.v[k];
    }
    __syncthreads();
    if (tid < kAcc) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) s += red[w][tid];
        partials[static_cast<size_t>(blockIdx.x) * kAcc + tid] = s;
        __threadfence();
    }
    __syncthreads();
    if (tid == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    double s[kAcc];
#pragma unroll
    for (int k = 0; k < kAcc; ++k) s[k] = 0.0;
    for (unsigned int r = tid; r < gridDim.x; r += kThreads) {
#pragma unroll
        for (int k = 0; k < kAcc; ++k) s[k] += __ldcg(partials + static_cast<size_t>(r) * kAcc + k);
    }
#pragma unroll
    for (int k = 0; k < kAcc; ++k) s[k] = warp_sum(s[k]);
    __syncthreads();
    if ((tid & 31) == 0) {
#pragma unroll
        for (int k = 0; k < kAcc; ++k) red[tid >> 5][k] = s[k];
    }
    __syncthreads();
    if (tid < kAcc) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) t += red[w][tid];
        if (tid < 4) {
            params->grad[tid] = t;
            // update_parameters_kernel, linear_regression_sgd.cu:126-134 (every other CTA has long read `value`)
            params->value[tid] -= lr * t / static_cast<double>(batch_size);
        } else if (loss_sum) {
            *loss_sum += t;
        }
    }
    if (tid == 0) *ticket = 0u;
}


// ---- MANY SGD epochs in ONE cooperative launch ----------------------------------------------------------------
// The epoch loop of the driver is launch bound even at one launch per epoch (6.6 us/epoch against ~2 us of device
// work).  This kernel keeps the parameters in registers and runs `n_epochs` epochs back to back; the only global
// synchronisation per epoch is one grid barrier between "every CTA has written its partial row" and "every CTA
// reads all rows".  EVERY CTA then adds the rows with exactly the instruction sequence of lsq_sgd_step_kernel's last
// CTA (same thread -> row mapping, same shuffle tree, same warp order) and applies the same update expression, so
// all CTAs hold bit-identical parameters and the final state is bit-identical to n_epochs calls of
// xyz_lsq_sgd_step_f64.  Rows are double-buffered by epoch parity (a CTA cannot be two barriers ahead).
// Launched with cudaLaunchCooperativeKernel (co-residency is guaranteed or the launch fails).
__global__ void __launch_bounds__(kThreads)
    lsq_sgd_run_kernel(const xyz_data_point* __restrict__ data, long long n_total, xyz_lsq_parameters* params,
                       long long batch_size, uint64_t seed, uint64_t epoch0, int n_epochs, const double* __restrict__ lrs,
                       double* loss_sum, double* partials, int mode) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ double red[kThreads / 32][kAcc];
    __shared__ double s_tot[kAcc];
    const int tid = threadIdx.x;
    double a = params->value[0], b = params->value[1], c = params->value[2], d = params->value[3];
    double loss_running = (loss_sum && blockIdx.x == 0 && tid == 0) ? *loss_sum : 0.0;
    double last_grad = 0.0;  // thread t < 4 of CTA 0: the last epoch's gradient component t
    const size_t row_stride = static_cast<size_t>(gridDim.x) * kAcc;
    for (int e = 0; e < n_epochs; ++e) {
        const uint64_t epoch = epoch0 + static_cast<uint64_t>(e);
        const double lr = lrs[e];
        double* rows = partials + static_cast<size_t>(e & 1) * row_stride;
        LsqAcc acc;
#pragma unroll
        for (int k = 0; k < kAcc; ++k) acc.v[k] = 0.0;
        const uint64_t base = splitmix64(seed ^ (epoch * 0xD1B54A32D192ED03ull));
        for (long long i = blockIdx.x * static_cast<long long>(kThreads) + tid; i < batch_size;
             i += static_cast<long long>(gridDim.x) * kThreads) {
            const uint64_t h = splitmix64(base + static_cast<uint64_t>(i));
            const double* p = reinterpret_cast<const double*>(data + h % static_cast<uint64_t>(n_total));
            lsq_point(__ldg(p), __ldg(p + 1), __ldg(p + 2), a, b, c, d, mode, acc);
        }
#pragma unroll
        for (int k = 0; k < kAcc; ++k) acc.v[k] = warp_sum(acc.v[k]);
        if ((tid & 31) == 0) {
#pragma unroll
            for (int k = 0; k < kAcc; ++k) red[tid >> 5][k] = acc.v[k];
        }
        __syncthreads();
        if (tid < kAcc) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) s += red[w][tid];
            rows[static_cast<size_t>(blockIdx.x) * kAcc + tid] = s;
        }
        grid.sync();  // all rows of this epoch are visible
        double s[kAcc];
#pragma unroll
        for (int k = 0; k < kAcc; ++k) s[k] = 0.0;
        for (unsigned int r = tid; r < gridDim.x; r += kThreads) {
#pragma unroll
            for (int k = 0; k < kAcc; ++k) s[k] += __ldcg(rows + static_cast<size_t>(r) * kAcc + k);
        }
#pragma unroll
        for (int k = 0; k < kAcc; ++k) s[k] = warp_sum(s[k]);
        __syncthreads();  // red is reused
        if ((tid & 31) == 0) {
#pragma unroll
            for (int k = 0; k < kAcc; ++k) red[tid >> 5][k] = s[k];
        }
        __syncthreads();
        if (tid < kAcc) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) t += red[w][tid];
            s_tot[tid] = t;
            if (blockIdx.x == 0 && tid < 4) last_grad = t;
        }
        __syncthreads();
        // update_parameters_kernel, linear_regression_sgd.cu:126-134 -- the expression of lsq_sgd_step_kernel
        a -= lr * s_tot[0] / static_cast<double>(batch_size);
        b -= lr * s_tot[1] / static_cast<double>(batch_size);
        c -= lr * s_tot[2] / static_cast<double>(batch_size);
        d -= lr * s_tot[3] / static_cast<double>(batch_size);
        if (blockIdx.x == 0 && tid == 0) loss_running += s_tot[4];
    }
    if (blockIdx.x == 0) {
        if (tid < 4) params->grad[tid] = last_grad;
        if (tid == 0) {
            params->value[0] = a;
            params->value[1] = b;
            params->value[2] = c;
            params->value[3] = d;
            if (loss_sum) *loss_sum = loss_running;
        }
    }
}

}  // namespace
}  // namespace xyzb

namespace xyzb {
namespace {
int lsq_grad_launch(const xyz_data_point* data, long long n_points, xyz_lsq_parameters* params, double* loss_sum,
                    const PeerArgs& peer, void* stream, int flags) {
    if (n_points < 0 || !params) return XYZ_ERR_INVALID_ARGUMENT;
    if (n_points == 0 && peer.world <= 1) return 0;
    if (n_points > 0 && !data) return XYZ_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long max_ctas = static_cast<long long>(sm_count()) * kCtasPerSM;
    const long long want = (n_points + kTileP - 1) / kTileP;
    const int grid = static_cast<int>(want < 1 ? 1 : (want < max_ctas ? want : max_ctas));  // >= 1: a rank with no
    void* scratch = nullptr;                                                                // points still exchanges
    const size_t bytes = 256 + static_cast<size_t>(max_ctas) * kAcc * sizeof(double);
    int err = scratch_get(SCRATCH_REDUCE, bytes, &scratch, st);
    if (err) return err;
    // scratch arenas are zero-filled when allocated and every kernel leaves its ticket at zero
    unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch);
    double* partials = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(scratch) + 256);
    const int mode = lsq_mode(flags);
    const double* d = reinterpret_cast<const double*>(data);
    if (aligned16(data) && n_points >= kTileP) {
        cudaFuncSetAttribute(lsq_grad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(sizeof(LsqSmem)));
        lsq_grad_kernel<true><<<grid, kThreads, sizeof(LsqSmem), st>>>(d, n_points, params, loss_sum, partials, ticket,
                                                                       mode, peer);
    } else {
        lsq_grad_kernel<false><<<grid, kThreads, sizeof(LsqSmem), st>>>(d, n_points, params, loss_sum, partials, ticket,
                                                                        mode, peer);
    }
    count_launch();
    return last_error();
}
}  // namespace
}  // namespace xyzb

extern "C" int xyz_lsq_grad_f64(const xyz_data_point* data, long long n_points, xyz_lsq_parameters* params,
                                double* loss_sum, void* stream, int flags) {
    xyzb::PeerArgs none{};
    none.world = 1;
    return xyzb::lsq_grad_launch(data, n_points, params, loss_sum, none, stream, flags);
}

extern "C" int xyz_lsq_grad_f64_allreduce(const xyz_data_point* data, long long n_points, xyz_lsq_parameters* params,
                                          double* loss_sum, const xyz_peer_group* group, unsigned long long seq,
                                          void* stream, int flags) {
    if (!group || group->world < 1 || group->world > XYZ_PEER_MAX_WORLD || group->rank < 0 || group->rank >= group->world ||
        seq == 0)
        return XYZ_ERR_INVALID_ARGUMENT;
    xyzb::PeerArgs pa{};
    pa.rank = group->rank;
    pa.world = group->world;
    pa.seq = seq;
    for (int i = 0; i < group->world; ++i) {
        if (!group->mailbox[i]) return XYZ_ERR_INVALID_ARGUMENT;
        pa.box[i] = static_cast<xyzb::PeerMailbox*>(group->mailbox[i]);
    }
    return xyzb::lsq_grad_launch(data, n_points, params, loss_sum, pa, stream, flags);
}

extern "C" int xyz_lsq_sgd_update_f64(xyz_lsq_parameters* params, double learning_rate, long long batch_size,
                                      void* stream) {
    using namespace xyzb;
    if (!params || batch_size <= 0) return XYZ_ERR_INVALID_ARGUMENT;
    lsq_sgd_update_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(params, learning_rate,
                                                                            static_cast<double>(batch_size));
    count_launch();
    return last_error();
}

extern "C" int xyz_lsq_select_batch(const xyz_data_point* data, long long n_total, xyz_data_point* batch,
                                    long long batch_size, uint64_t seed, uint64_t epoch, void* stream) {
    using namespace xyzb;
    if (!data || !batch || n_total <= 0 || batch_size < 0) return XYZ_ERR_INVALID_ARGUMENT;
    if (batch_size == 0) return 0;
    const int grid = static_cast<int>((batch_size + 255) / 256);
    lsq_select_batch_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(data, n_total, batch, batch_size, seed,
                                                                                  epoch);
    count_launch();
    return last_error();
}

extern "C" int xyz_lsq_sgd_step_f64(const xyz_data_point* data, long long n_total, xyz_lsq_parameters* params,
                                    long long batch_size, uint64_t seed, uint64_t epoch, double learning_rate,
                                    double* loss_sum, void* stream, int flags) {
    using namespace xyzb;
    if (!data || !params || n_total <= 0 || batch_size <= 0) return XYZ_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long max_ctas = static_cast<long long>(sm_count()) * kCtasPerSM;
    const long long want = (batch_size + kThreads - 1) / kThreads;
    const int grid = static_cast<int>(want < max_ctas ? want : max_ctas);
    void* scratch = nullptr;
    int err = scratch_get(SCRATCH_REDUCE, 256 + static_cast<size_t>(max_ctas) * kAcc * sizeof(double), &scratch, st);
    if (err) return err;
    unsigned int* ticket = reinterpret_cast<unsigned int*>(scratch);
    double* partials = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(scratch) + 256);
    lsq_sgd_step_kernel<<<grid, kThreads, 0, st>>>(data, n_total, params, batch_size, seed, epoch, learning_rate, loss_sum,
                                                    partials, ticket, lsq_mode(flags));
    count_launch();
    return last_error();
}

extern "C" int xyz_lsq_sgd_run_f64(const xyz_data_point* data, long long n_total, xyz_lsq_parameters* params,
                                   long long batch_size, uint64_t seed, uint64_t epoch_begin, int n_epochs,
                                   const double* learning_rates_host, double* loss_sum, void* stream, int flags) {
    using namespace xyzb;
    if (!data || !params || n_total <= 0 || batch_size <= 0 || n_epochs < 0 || (n_epochs > 0 && !learning_rates_host))
        return XYZ_ERR_INVALID_ARGUMENT;
    if (n_epochs == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // the grid of xyz_lsq_sgd_step_f64 (the row order is part of the result), provided it can be co-resident
    const long long max_ctas = static_cast<long long>(sm_count()) * kCtasPerSM;
    const long long want = (batch_size + kThreads - 1) / kThreads;
    const int grid = static_cast<int>(want < max_ctas ? want : max_ctas);
    int per_sm = 0;
    cudaError_t ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, lsq_sgd_run_kernel, kThreads, 0);
    int coop = 0, dev = 0;
    if (ce == cudaSuccess) ce = cudaGetDevice(&dev);
    if (ce == cudaSuccess) ce = cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (ce != cudaSuccess) return static_cast<int>(ce);
    const int mode = lsq_mode(flags);
    if (!coop || static_cast<long long>(per_sm) * sm_count() < grid) {  // cannot be co-resident: one launch per epoch
        for (int e = 0; e < n_epochs; ++e) {
            const int err = xyz_lsq_sgd_step_f64(data, n_total, params, batch_size, seed, epoch_begin + static_cast<uint64_t>(e),
                                                 learning_rates_host[e], loss_sum, stream, flags);
            if (err) return err;
        }
        return 0;
    }
    constexpr int kMaxEpochsPerLaunch = 4096;
    void* scratch = nullptr;
    const size_t rows_bytes = 2 * static_cast<size_t>(max_ctas) * kAcc * sizeof(double);
    int err = scratch_get(SCRATCH_REDUCE, 256 + rows_bytes + kMaxEpochsPerLaunch * sizeof(double), &scratch, st);
    if (err) return err;
    double* partials = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(scratch) + 256);
    double* lrs_dev = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(scratch) + 256 + rows_bytes);
    for (int done = 0; done < n_epochs; done += kMaxEpochsPerLaunch) {
        int n = n_epochs - done < kMaxEpochsPerLaunch ? n_epochs - done : kMaxEpochsPerLaunch;
        ce = cudaMemcpyAsync(lrs_dev, learning_rates_host + done, sizeof(double) * n, cudaMemcpyHostToDevice, st);
        if (ce != cudaSuccess) return static_cast<int>(ce);
        uint64_t epoch0 = epoch_begin + static_cast<uint64_t>(done);
        const double* lrs_arg = lrs_dev;
        void* args[] = {&data, &n_total, &params, &batch_size, &seed, &epoch0, &n, &lrs_arg, &loss_sum, &partials,
                        const_cast<int*>(&mode)};
        ce = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(lsq_sgd_run_kernel), dim3(grid), dim3(kThreads), args, 0, st);
        if (ce != cudaSuccess) return static_cast<int>(ce);
        count_launch();
    }
    return last_error();
}
