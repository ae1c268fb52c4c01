// csrc/peer.cu -- peer-memory mailboxes for the fused "reduce + all-reduce" kernels (one process per GPU).
//
// The reference has no multi-GPU path; on an NVSwitch box the only exchange of this hot path is
// the sum of SHARED-parameter gradients (4 fp64 for least squares, K fp32 for the accumulation pattern).  Those
// vectors are tiny, so an NCCL all-reduce behind the kernel costs more than the kernel itself (launch + protocol
// latency).  Instead every rank owns a small MAILBOX in its own HBM, exported with CUDA IPC; the last CTA of the
// gradient kernel stores its partial row into slot[rank] of EVERY rank's mailbox (plain st.global over NVLink),
// publishes a sequence number with st.release.sys, waits for the sequence numbers of all ranks in its own
// mailbox, and adds the rows in rank order: one kernel = compute + collective, every rank gets bit-identical sums.
//
// Layout of one mailbox (device memory of the owning rank):
//   double data[2][XYZ_PEER_MAX_WORLD][XYZ_PEER_SLOT_DOUBLES]   two parities (sequence & 1) so a fast rank may run
//                                                                one call ahead of a slow reader
//   unsigned long long flag[XYZ_PEER_MAX_WORLD]                 flag[q] = last sequence number rank q published here
#include "common.cuh"

#include <cmath>

namespace xyzb {
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "xyz_peer_* passes IPC handles as 64 bytes");
}

extern "C" size_t xyz_peer_mailbox_bytes(void) { return sizeof(xyzb::PeerMailbox); }

extern "C" int xyz_peer_mailbox_create(void** local_ptr, unsigned char ipc_handle_out[64]) {
    if (!local_ptr || !ipc_handle_out) return XYZ_ERR_INVALID_ARGUMENT;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, sizeof(xyzb::PeerMailbox));
    if (e != cudaSuccess) return static_cast<int>(e);
    e = cudaMemset(p, 0, sizeof(xyzb::PeerMailbox));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return static_cast<int>(e);
    }
    memcpy(ipc_handle_out, &h, 64);
    *local_ptr = p;
    return 0;
}

extern "C" int xyz_peer_mailbox_open(const unsigned char ipc_handle[64], void** peer_ptr) {
    if (!ipc_handle || !peer_ptr) return XYZ_ERR_INVALID_ARGUMENT;
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, 64);
    return static_cast<int>(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
}

extern "C" int xyz_peer_mailbox_close(void* peer_ptr) {
    if (!peer_ptr) return 0;
    return static_cast<int>(cudaIpcCloseMemHandle(peer_ptr));
}

extern "C" int xyz_peer_mailbox_destroy(void* local_ptr) {
    if (!local_ptr) return 0;
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return static_cast<int>(e);
    return static_cast<int>(cudaFree(local_ptr));
}

extern "C" int xyz_peer_alloc(size_t bytes, void** local_ptr, unsigned char ipc_handle_out[64]) {
    if (!local_ptr || bytes == 0) return XYZ_ERR_INVALID_ARGUMENT;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return static_cast<int>(e);
    e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess && ipc_handle_out) {
        cudaIpcMemHandle_t h;
        e = cudaIpcGetMemHandle(&h, p);
        if (e == cudaSuccess) memcpy(ipc_handle_out, &h, 64);
    }
    if (e != cudaSuccess) {
        cudaFree(p);
        return static_cast<int>(e);
    }
    *local_ptr = p;
    return 0;
}

// ---- fused reduce-scatter + Adam + all-gather of the splat parameters over NVLink peer memory --------------------------
// (xyz_adam_step_individual_peer, include/xyz_b200.h).  Replaces, for a sharded splat iteration, the sequence
//   ncclAllReduce(grads) ; adam_step_individual on every replica ; zero_gradients
// (reference per-iteration steps gaussian_splatting_training.cu:131-158, made multi-GPU) by ONE kernel per rank:
//   round 1  every rank publishes "my gradients are complete" (sequence seq) into every mailbox and waits for all ranks;
//   work     the owner of a float range loads that range from EVERY rank's gradient buffer (NVLink, 16-byte loads), adds
//            the rows in rank order, applies the reference's Adam update (gaussian_parameters.cu:260-320) to its own
//            moments, stores the new parameters into EVERY rank's parameter buffer and zeroes the range in every rank's
//            gradient buffer (zero_gradients_kernel, :227-241);
//   round 2  the last CTA to finish publishes seq + 1 and waits for all ranks: when the kernel ends, every remote store
//            into this rank's buffers has landed, so the next iteration may start.
// Bytes over NVLink per rank: (world - 1) / world x N x 36 in (gradients) and out (parameters) + the same out again for
// the zeroes -- the wire traffic of a reduce-scatter + all-gather, with no staging copy and no second launch.
namespace xyzb {
namespace {

struct PeerAdamArgs {
    float* params[XYZ_PEER_MAX_WORLD];
    float* grads[XYZ_PEER_MAX_WORLD];
    float beta1, beta2, eps;
};

__device__ __forceinline__ void adam_slot_of(int j, int& group, int& m_off, int& v_off) {
    // AdamState (gaussian_parameters.h:21-32): m_center[2] v_center[2] m_scale[2] v_scale[2] m_rot v_rot m_color[3]
    // v_color[3] m_op v_op
    if (j < 2) { group = 0; m_off = j; v_off = 2 + j; }
    else if (j < 4) { group = 1; m_off = 4 + (j - 2); v_off = 6 + (j - 2); }
    else if (j == 4) { group = 2; m_off = 8; v_off = 9; }
    else if (j < 8) { group = 3; m_off = 10 + (j - 5); v_off = 13 + (j - 5); }
    else { group = 4; m_off = 16; v_off = 17; }
}

__device__ __forceinline__ float adam_component(float* __restrict__ adam, long long f, float grad, float param,
                                                const PeerAdamArgs& a, const float* lr_corrected /* shared memory, 5 */) {
    const long long g = f / 9;
    const int j = static_cast<int>(f - g * 9);
    int group, m_off, v_off;
    adam_slot_of(j, group, m_off, v_off);
    float* st = adam + g * 18;
    const float m = a.beta1 * st[m_off] + (1.0f - a.beta1) * grad;
    const float v = a.beta2 * st[v_off] + (1.0f - a.beta2) * grad * grad;
    st[m_off] = m;
    st[v_off] = v;
    return param - lr_corrected[group] * m / (sqrtf(v) + a.eps);
}

__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}

constexpr int kPeerAdamThreads = 256;

__global__ void __launch_bounds__(kPeerAdamThreads)
    peer_adam_kernel(PeerArgs pa, PeerAdamArgs a, float* __restrict__ adam, long long n_floats, float* total_loss,
                     int iteration, float lr0, float lr1, float lr2, float lr3, float lr4) {
    __shared__ int s_ok;
    __shared__ int s_last;
    __shared__ float s_lr[5];
    const int tid = threadIdx.x;
    const int W = pa.world, R = pa.rank;
    PeerMailbox* const me = pa.box[R];
    unsigned int* const ticket = &me->splat_ticket;  // in the mailbox, not in library scratch: nothing to allocate, ever
    // sequence numbers live on the device (graph replays): this call uses seq (round 1) and seq + 1 (round 2).
    // Nobody changes splat_seq / splat_iter before the LAST CTA of this launch has read them (it does so below).
    const unsigned long long seq = me->splat_seq + 1ull;
    const int par = static_cast<int>((seq >> 1) & 1ull);
    if (tid == 0) {
        s_ok = 1;
        // bias correction of the reference's host wrapper (gaussian_parameters.cu:357-358 + kernel :279-283);
        // iteration == 0: the number of steps this mailbox has counted so far + 1
        const int it = iteration > 0 ? iteration : static_cast<int>(me->splat_iter) + 1;
        const float b1t = powf(a.beta1, static_cast<float>(it)), b2t = powf(a.beta2, static_cast<float>(it));
        const float c = sqrtf(1.0f - b2t) / (1.0f - b1t);
        s_lr[0] = lr0 * c; s_lr[1] = lr1 * c; s_lr[2] = lr2 * c; s_lr[3] = lr3 * c; s_lr[4] = lr4 * c;
    }
    // A group of ONE rank has nobody to wait for and nothing to publish: no flags, no system-scope fences (ncu: the two
    // rounds and their three membar.sys cost ~25 of this kernel's 35 us, whatever the number of ranks) -- the kernel is
    // then the Adam step with the zero-gradient pass fused in.
    const bool alone = W == 1;
    // ---- round 1: gradients (and the loss) of every rank are complete
    if (blockIdx.x == 0 && !alone) {
        if (tid == 0 && total_loss) {
            const double mine = static_cast<double>(*total_loss);
            for (int p = 0; p < W; ++p) pa.box[p]->loss_splat[par][R] = mine;
            __threadfence_system();
        }
        __syncthreads();
        if (tid < W) st_release_sys(&pa.box[tid]->flag_splat[R], seq);
    }
    __syncthreads();
    if (tid < W && !alone) {
        const unsigned long long* f = &me->flag_splat[tid];
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(f) < seq) {
            if (global_timer_ns() - t0 > 4000000000ull) {
                s_ok = 0;
                break;
            }
        }
    }
    __syncthreads();
    const bool ok = s_ok != 0;
    if (blockIdx.x == 0 && tid == 0 && total_loss && !alone) {
        double s = 0.0;
        for (int q = 0; q < W; ++q) s += ld_relaxed_sys_f64(&me->loss_splat[par][q]);  // rank order: same bits everywhere
        *total_loss = ok ? static_cast<float>(s) : __int_as_float(0x7fc00000);
    }
    // ---- work: this rank's range of 16-byte groups (the last rank also takes the < 4 floats that do not fill one)
    if (ok) {
        const long long n4 = n_floats / 4;
        const long long b4 = n4 * R / W, e4 = n4 * (R + 1) / W;
        for (long long i = b4 + blockIdx.x * static_cast<long long>(kPeerAdamThreads) + tid; i < e4;
             i += static_cast<long long>(gridDim.x) * kPeerAdamThreads) {
            const long long f0 = 4 * i;
            // all ranks' rows first (up to 8 independent NVLink loads in flight), then the sum in rank order
            float4 v[XYZ_PEER_MAX_WORLD];
#pragma unroll
            for (int q = 0; q < XYZ_PEER_MAX_WORLD; ++q)
                if (q < W) v[q] = ld_peer_f4(a.grads[q] + f0);
            const float4 p = *reinterpret_cast<const float4*>(a.params[R] + f0);
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int q = 0; q < XYZ_PEER_MAX_WORLD; ++q)
                if (q < W) {
                    s.x += v[q].x; s.y += v[q].y; s.z += v[q].z; s.w += v[q].w;
                }
            float4 np;
            np.x = adam_component(adam, f0, s.x, p.x, a, s_lr);
            np.y = adam_component(adam, f0 + 1, s.y, p.y, a, s_lr);
            np.z = adam_component(adam, f0 + 2, s.z, p.z, a, s_lr);
            np.w = adam_component(adam, f0 + 3, s.w, p.w, a, s_lr);
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int q = 0; q < XYZ_PEER_MAX_WORLD; ++q)
                if (q < W) {
                    *reinterpret_cast<float4*>(a.params[q] + f0) = np;
                    *reinterpret_cast<float4*>(a.grads[q] + f0) = z;
                }
        }
        if (R == W - 1 && blockIdx.x == 0 && tid < static_cast<int>(n_floats - 4 * n4)) {
            const long long f = 4 * n4 + tid;
            float s = 0.f;
            for (int q = 0; q < W; ++q) s += ld_relaxed_sys_f32(a.grads[q] + f);
            const float np = adam_component(adam, f, s, a.params[R][f], a, s_lr);
            for (int q = 0; q < W; ++q) {
                a.params[q][f] = np;
                a.grads[q][f] = 0.f;
            }
        }
    }
    // ---- round 2: all my remote stores have landed; wait until everybody else's have landed here.
    // One system-scope fence per CTA, by the thread that takes the ticket: the CTA barrier orders every thread's stores
    // before it, and fences are cumulative (a membar.sys in all 256 threads costs microseconds per CTA).
    __syncthreads();
    if (tid == 0) {
        if (alone) __threadfence();
        else __threadfence_system();
        s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    if (!alone) __threadfence_system();
    if (tid == 0) {
        *ticket = 0u;  // ready for the next launch
        me->splat_seq = seq + 1ull;
        me->splat_iter = me->splat_iter + 1ull;
    }
    if (tid < W && !alone) {
        st_release_sys(&pa.box[tid]->flag_splat[R], seq + 1ull);
        const unsigned long long* f = &me->flag_splat[tid];
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(f) < seq + 1ull) {
            if (global_timer_ns() - t0 > 4000000000ull) break;
        }
    }
}

}  // namespace
}  // namespace xyzb

extern "C" int xyz_adam_step_individual_peer(const xyz_peer_group* group, const xyz_peer_splat_buffers* buffers,
                                             xyz_adam_state* adam, int num_gaussians, const float lr_host[5], float beta1,
                                             float beta2, float epsilon, int iteration, float* total_loss, void* stream) {
    using namespace xyzb;
    if (!group || !buffers || !lr_host || num_gaussians < 0 || iteration < 0) return XYZ_ERR_INVALID_ARGUMENT;
    if (group->world < 1 || group->world > XYZ_PEER_MAX_WORLD || group->rank < 0 || group->rank >= group->world)
        return XYZ_ERR_INVALID_ARGUMENT;
    if (num_gaussians > 0 && !adam) return XYZ_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    PeerArgs pa{};
    PeerAdamArgs a{};
    for (int r = 0; r < group->world; ++r) {
        if (!group->mailbox[r] || (num_gaussians > 0 && (!buffers->params[r] || !buffers->grads[r]))) return XYZ_ERR_INVALID_ARGUMENT;
        pa.box[r] = static_cast<PeerMailbox*>(group->mailbox[r]);
        a.params[r] = reinterpret_cast<float*>(buffers->params[r]);
        a.grads[r] = reinterpret_cast<float*>(buffers->grads[r]);
        if (!aligned16(a.params[r]) || !aligned16(a.grads[r])) return XYZ_ERR_INVALID_ARGUMENT;
    }
    pa.rank = group->rank;
    pa.world = group->world;
    pa.seq = 0;  // unused: the sequence lives in the mailbox
    a.beta1 = beta1; a.beta2 = beta2; a.eps = epsilon;
    const long long n_floats = static_cast<long long>(num_gaussians) * 9;
    const long long mine4 = n_floats / 4 / group->world + 1;
    long long want = (mine4 + kPeerAdamThreads - 1) / kPeerAdamThreads;
    const long long cap = 4LL * sm_count();
    const int grid = static_cast<int>(want < 1 ? 1 : (want > cap ? cap : want));
    peer_adam_kernel<<<grid, kPeerAdamThreads, 0, st>>>(pa, a, reinterpret_cast<float*>(adam), n_floats, total_loss, iteration,
                                                        lr_host[0], lr_host[1], lr_host[2], lr_host[3], lr_host[4]);
    count_launch();
    return last_error();
}
