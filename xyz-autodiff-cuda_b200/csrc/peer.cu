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
