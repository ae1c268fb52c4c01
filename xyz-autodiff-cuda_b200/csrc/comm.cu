// csrc/comm.cu -- NCCL exchange of the shared-parameter gradients (C4 on G GPUs, C5), SURVEY 8b / 8e.
//
// The reference has no communication layer (single GPU, gaussian_splatting_training.cu:209).  What a sharded splat
// iteration needs is the sum over ranks of the N x 9 gradient buffer (108 MB at C5) -- or, without redundant optimiser
// work, its reduce-scatter by Gaussian range, Adam on the range, and an all-gather of the parameters.
//
// libnccl.so.2 is bound at RUN time: a process that already carries NCCL (PyTorch bundles its own copy) keeps using that
// one copy, a plain C++ host gets the system library, and libxyz_b200.so has no link-time dependency on either.  Only
// the types of <nccl.h> are used at compile time.
#include "common.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <mutex>
#include <vector>

struct xyz_comm {
    ncclComm_t comm;
    int rank, world;
    bool owned;
};

namespace xyzb {
int adam_launch_range(xyz_gaussian_params* params, const xyz_gaussian_grads* grads, xyz_adam_state* adam, int g_begin,
                      int g_end, const float lr[5], float beta1, float beta2, float eps, int iteration, void* stream);

namespace {
struct Nccl {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    bool ok = false;
};
Nccl g_nccl;
std::once_flag g_nccl_once;

template <class F>
bool bind(void* h, const char* name, F& fn) {
    fn = reinterpret_cast<F>(dlsym(h, name));
    return fn != nullptr;
}

const Nccl& nccl() {
    std::call_once(g_nccl_once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // the copy the process already has
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        Nccl& n = g_nccl;
        n.handle = h;
        n.ok = bind(h, "ncclGetUniqueId", n.GetUniqueId) && bind(h, "ncclCommInitRank", n.CommInitRank) &&
               bind(h, "ncclCommInitAll", n.CommInitAll) && bind(h, "ncclCommDestroy", n.CommDestroy) &&
               bind(h, "ncclAllReduce", n.AllReduce) && bind(h, "ncclReduce", n.Reduce) &&
               bind(h, "ncclBroadcast", n.Broadcast) && bind(h, "ncclGroupStart", n.GroupStart) &&
               bind(h, "ncclGroupEnd", n.GroupEnd);
    });
    return g_nccl;
}

inline int nccl_code(ncclResult_t r) { return r == ncclSuccess ? 0 : XYZ_ERR_COMM; }

// Gaussian range of rank r: contiguous, balanced
inline void gaussian_range(int n, int rank, int world, int& g0, int& g1) {
    g0 = static_cast<int>(static_cast<long long>(n) * rank / world);
    g1 = static_cast<int>(static_cast<long long>(n) * (rank + 1) / world);
}
}  // namespace
}  // namespace xyzb

using xyzb::nccl;
using xyzb::nccl_code;

extern "C" int xyz_comm_unique_id(unsigned char id_out[128]) {
    static_assert(sizeof(ncclUniqueId) == 128, "xyz_comm_unique_id passes the NCCL id as 128 bytes");
    if (!id_out) return XYZ_ERR_INVALID_ARGUMENT;
    if (!nccl().ok) return XYZ_ERR_COMM;
    ncclUniqueId id;
    const int e = nccl_code(nccl().GetUniqueId(&id));
    if (e) return e;
    memcpy(id_out, &id, 128);
    return 0;
}

extern "C" int xyz_comm_init_rank(xyz_comm** comm_out, const unsigned char id[128], int rank, int world) {
    if (!comm_out || !id || world < 1 || rank < 0 || rank >= world) return XYZ_ERR_INVALID_ARGUMENT;
    if (!nccl().ok) return XYZ_ERR_COMM;
    ncclUniqueId uid;
    memcpy(&uid, id, 128);
    ncclComm_t c = nullptr;
    const int e = nccl_code(nccl().CommInitRank(&c, world, uid, rank));
    if (e) return e;
    *comm_out = new xyz_comm{c, rank, world, true};
    return 0;
}

extern "C" int xyz_comm_init(xyz_comm** comm_out, void* nccl_comm, int rank, int world) {
    if (!comm_out || !nccl_comm || world < 1 || rank < 0 || rank >= world) return XYZ_ERR_INVALID_ARGUMENT;
    if (!nccl().ok) return XYZ_ERR_COMM;
    *comm_out = new xyz_comm{static_cast<ncclComm_t>(nccl_comm), rank, world, false};
    return 0;
}

extern "C" int xyz_comm_init_all(xyz_comm** comms_out, int ndev, const int* devices) {
    if (!comms_out || ndev < 1 || ndev > 64) return XYZ_ERR_INVALID_ARGUMENT;
    if (!nccl().ok) return XYZ_ERR_COMM;
    std::vector<ncclComm_t> c(ndev);
    const int e = nccl_code(nccl().CommInitAll(c.data(), ndev, devices));
    if (e) return e;
    for (int i = 0; i < ndev; ++i) comms_out[i] = new xyz_comm{c[i], i, ndev, true};
    return 0;
}

extern "C" int xyz_comm_destroy(xyz_comm* comm) {
    if (!comm) return 0;
    int e = 0;
    if (comm->owned && nccl().ok) e = nccl_code(nccl().CommDestroy(comm->comm));
    delete comm;
    return e;
}

extern "C" int xyz_comm_rank(const xyz_comm* comm) { return comm ? comm->rank : XYZ_ERR_INVALID_ARGUMENT; }
extern "C" int xyz_comm_world(const xyz_comm* comm) { return comm ? comm->world : XYZ_ERR_INVALID_ARGUMENT; }

extern "C" int xyz_comm_group_start(void) { return nccl().ok ? nccl_code(nccl().GroupStart()) : XYZ_ERR_COMM; }
extern "C" int xyz_comm_group_end(void) { return nccl().ok ? nccl_code(nccl().GroupEnd()) : XYZ_ERR_COMM; }

extern "C" int xyz_allreduce_grads(xyz_comm* comm, float* grads, long long n, void* stream) {
    if (!comm || n < 0 || (n > 0 && !grads)) return XYZ_ERR_INVALID_ARGUMENT;
    if (n == 0 || comm->world == 1) return 0;
    return nccl_code(nccl().AllReduce(grads, grads, static_cast<size_t>(n), ncclFloat32, ncclSum, comm->comm,
                                      static_cast<cudaStream_t>(stream)));
}

extern "C" int xyz_allreduce_f64(xyz_comm* comm, double* values, long long n, void* stream) {
    if (!comm || n < 0 || (n > 0 && !values)) return XYZ_ERR_INVALID_ARGUMENT;
    if (n == 0 || comm->world == 1) return 0;
    return nccl_code(nccl().AllReduce(values, values, static_cast<size_t>(n), ncclFloat64, ncclSum, comm->comm,
                                      static_cast<cudaStream_t>(stream)));
}

// reduce-scatter by Gaussian range (ranges differ by at most one Gaussian, so: one ncclReduce per range inside a group)
// -> adam_step_individual on this rank's range -> zero the whole local gradient buffer -> all-gather (one ncclBroadcast
// per range inside a group).
extern "C" int xyz_adam_step_individual_sharded(xyz_comm* comm, xyz_gaussian_params* params, xyz_gaussian_grads* grads,
                                                xyz_adam_state* adam, int num_gaussians, const float lr_host[5],
                                                float beta1, float beta2, float epsilon, int iteration,
                                                float* total_loss, void* stream) {
    using namespace xyzb;
    if (!comm || !lr_host || num_gaussians < 0 || iteration < 1) return XYZ_ERR_INVALID_ARGUMENT;
    if (num_gaussians > 0 && (!params || !grads || !adam)) return XYZ_ERR_INVALID_ARGUMENT;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int world = comm->world, rank = comm->rank;
    float* g = reinterpret_cast<float*>(grads);
    float* p = reinterpret_cast<float*>(params);
    int e = 0;
    if (world > 1) {
        if ((e = nccl_code(nccl().GroupStart()))) return e;
        for (int r = 0; r < world; ++r) {
            int g0, g1;
            gaussian_range(num_gaussians, r, world, g0, g1);
            if (g1 > g0)
                e |= nccl_code(nccl().Reduce(g + 9LL * g0, g + 9LL * g0, 9ULL * (g1 - g0), ncclFloat32, ncclSum, r, comm->comm, st));
        }
        if (total_loss) e |= nccl_code(nccl().AllReduce(total_loss, total_loss, 1, ncclFloat32, ncclSum, comm->comm, st));
        e |= nccl_code(nccl().GroupEnd());
        if (e) return XYZ_ERR_COMM;
    }
    int g0, g1;
    gaussian_range(num_gaussians, rank, world, g0, g1);
    e = adam_launch_range(params, grads, adam, g0, g1, lr_host, beta1, beta2, epsilon, iteration, stream);
    if (e) return e;
    e = xyz_zero_gradients(grads, num_gaussians, stream);
    if (e) return e;
    if (world > 1) {
        if ((e = nccl_code(nccl().GroupStart()))) return e;
        for (int r = 0; r < world; ++r) {
            int r0, r1;
            gaussian_range(num_gaussians, r, world, r0, r1);
            if (r1 > r0)
                e |= nccl_code(nccl().Broadcast(p + 9LL * r0, p + 9LL * r0, 9ULL * (r1 - r0), ncclFloat32, r, comm->comm, st));
        }
        e |= nccl_code(nccl().GroupEnd());
        if (e) return XYZ_ERR_COMM;
    }
    return 0;
}
