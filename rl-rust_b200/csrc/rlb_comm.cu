// rlb_comm.cu — the path's one multi-GPU exchange behind the C ABI (include/rlb.h, "multi-GPU").
//
// Agents are independent, so nothing is exchanged while they train.  What a job gathers is the per-episode metric
// sums each engine reduces on its own GPU ([n_episodes][4] f64: sum of lengths, returns, td, |td| — the curves of
// bin/taxi.rs:170-223): one gather to the root rank, done with grouped ncclSend / ncclRecv (the system NCCL, 2.27, has
// no ncclGather).  NCCL is resolved at run time with dlopen("libnccl.so.2"): librlb.so itself links only the static
// CUDA runtime, loads on boxes without NCCL, and inside a process that already holds a NCCL (PyTorch's) the loader
// hands back that same copy.
#include <dlfcn.h>

#include <cstring>
#include <mutex>
#include <vector>

#include <cuda_runtime.h>
#include <nccl.h>   // types and prototypes only; no symbol of it is linked

#include "rlb_host.h"

using namespace rlb;

namespace {

struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    bool ok = false;
    std::string why;
};

NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) { api.why = std::string("dlopen(libnccl.so.2): ") + (dlerror() ? dlerror() : "not found"); return; }
        bool all = true;
#define RLB_SYM(field, name)                                                          \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.handle, name));       \
    if (!api.field) { all = false; api.why = std::string("libnccl lacks ") + name; }
        RLB_SYM(GetUniqueId, "ncclGetUniqueId")
        RLB_SYM(CommInitRank, "ncclCommInitRank")
        RLB_SYM(CommInitAll, "ncclCommInitAll")
        RLB_SYM(CommDestroy, "ncclCommDestroy")
        RLB_SYM(GetErrorString, "ncclGetErrorString")
        RLB_SYM(Send, "ncclSend")
        RLB_SYM(Recv, "ncclRecv")
        RLB_SYM(AllReduce, "ncclAllReduce")
        RLB_SYM(GroupStart, "ncclGroupStart")
        RLB_SYM(GroupEnd, "ncclGroupEnd")
        RLB_SYM(GetVersion, "ncclGetVersion")
#undef RLB_SYM
        api.ok = all;
    });
    return api;
}

rlb_status nccl_missing() {
    set_error("NCCL unavailable: %s", nccl().why.c_str());
    return RLB_ERR_NCCL;
}
rlb_status nccl_fail(ncclResult_t r, const char* what) {
    set_error("%s: %s", what, nccl().GetErrorString(r));
    return RLB_ERR_NCCL;
}
rlb_status cuda_fail(cudaError_t err, const char* what) {
    set_error("%s: %s", what, cudaGetErrorString(err));
    return RLB_ERR_CUDA;
}

#define NK(call)                                                     \
    do {                                                             \
        ncclResult_t r__ = (call);                                   \
        if (r__ != ncclSuccess) return nccl_fail(r__, #call);        \
    } while (0)
#define CK(call)                                                     \
    do {                                                             \
        cudaError_t e__ = (call);                                    \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call);        \
    } while (0)

}   // namespace

struct rlb_comm {
    ncclComm_t comm = nullptr;
    int32_t rank = 0, world = 1, device = 0;
    cudaStream_t stream = nullptr;   // used when the caller passes no stream
};

extern "C" {

rlb_status rlb_comm_unique_id(uint8_t id_out[RLB_COMM_ID_BYTES]) {
    static_assert(RLB_COMM_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "rlb.h carries NCCL's unique id verbatim");
    if (!id_out) { set_error("id_out is NULL"); return RLB_ERR_INVALID_ARG; }
    if (!nccl().ok) return nccl_missing();
    ncclUniqueId id;
    NK(nccl().GetUniqueId(&id));
    std::memcpy(id_out, id.internal, RLB_COMM_ID_BYTES);
    return RLB_OK;
}

rlb_status rlb_comm_init_rank(const uint8_t id[RLB_COMM_ID_BYTES], int32_t world_size, int32_t rank, int32_t device, rlb_comm** out) {
    if (!out) { set_error("out is NULL"); return RLB_ERR_INVALID_ARG; }
    *out = nullptr;
    if (!id || world_size < 1 || rank < 0 || rank >= world_size) { set_error("bad id / world_size / rank"); return RLB_ERR_INVALID_ARG; }
    if (!nccl().ok) return nccl_missing();
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        set_error("device %d out of range (have %d)", device, ndev);
        return ndev == 0 ? RLB_ERR_CUDA : RLB_ERR_INVALID_ARG;
    }
    CK(cudaSetDevice(device));
    rlb_comm* c = new rlb_comm();
    c->rank = rank; c->world = world_size; c->device = device;
    ncclUniqueId uid;
    std::memcpy(uid.internal, id, RLB_COMM_ID_BYTES);
    ncclResult_t r = nccl().CommInitRank(&c->comm, world_size, uid, rank);
    if (r != ncclSuccess) { delete c; return nccl_fail(r, "ncclCommInitRank"); }
    cudaError_t err = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (err != cudaSuccess) { nccl().CommDestroy(c->comm); delete c; return cuda_fail(err, "cudaStreamCreateWithFlags"); }
    *out = c;
    return RLB_OK;
}

rlb_status rlb_comm_init_all(const int32_t* devices, int32_t n_devices, rlb_comm** comms_out) {
    if (!devices || !comms_out || n_devices < 1) { set_error("bad devices / comms_out"); return RLB_ERR_INVALID_ARG; }
    for (int i = 0; i < n_devices; ++i) comms_out[i] = nullptr;
    if (!nccl().ok) return nccl_missing();
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); set_error("no CUDA device"); return RLB_ERR_CUDA; }
    std::vector<int> devs(devices, devices + n_devices);
    for (int d : devs) if (d < 0 || d >= ndev) { set_error("device %d out of range (have %d)", d, ndev); return RLB_ERR_INVALID_ARG; }
    std::vector<ncclComm_t> comms(n_devices);
    NK(nccl().CommInitAll(comms.data(), n_devices, devs.data()));
    for (int i = 0; i < n_devices; ++i) {
        rlb_comm* c = new rlb_comm();
        c->comm = comms[i]; c->rank = i; c->world = n_devices; c->device = devs[i];
        CK(cudaSetDevice(devs[i]));
        CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        comms_out[i] = c;
    }
    return RLB_OK;
}

void rlb_comm_destroy(rlb_comm* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    if (c->comm && nccl().ok) nccl().CommDestroy(c->comm);
    cudaGetLastError();
    delete c;
}
int32_t rlb_comm_rank(const rlb_comm* c) { return c ? c->rank : -1; }
int32_t rlb_comm_world_size(const rlb_comm* c) { return c ? c->world : 0; }

rlb_status rlb_comm_group_begin(void) {
    if (!nccl().ok) return nccl_missing();
    NK(nccl().GroupStart());
    return RLB_OK;
}
rlb_status rlb_comm_group_end(void) {
    if (!nccl().ok) return nccl_missing();
    NK(nccl().GroupEnd());
    return RLB_OK;
}

rlb_status rlb_comm_gather_episode_sums(rlb_comm* c, const double* local_sums, uint64_t n_episodes, double* gathered_out,
                                        int32_t root, void* cuda_stream) {
    if (!c || !local_sums) { set_error("NULL argument"); return RLB_ERR_INVALID_ARG; }
    if (root < 0 || root >= c->world) { set_error("root %d out of range (world %d)", root, c->world); return RLB_ERR_INVALID_ARG; }
    if (c->rank == root && !gathered_out) { set_error("gathered_out is NULL on the root rank"); return RLB_ERR_INVALID_ARG; }
    CK(cudaSetDevice(c->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : c->stream;
    const size_t count = (size_t)n_episodes * 4;
    if (count) {
        // the root's own block is a device copy; every other block is one send / recv pair, all in one group so
        // that the root's world-1 receives progress together
        if (c->rank == root) CK(cudaMemcpyAsync(gathered_out + (size_t)root * count, local_sums, count * sizeof(double), cudaMemcpyDeviceToDevice, s));
        if (c->world > 1) {
            NK(nccl().GroupStart());
            if (c->rank == root) {
                for (int r = 0; r < c->world; ++r) {
                    if (r == root) continue;
                    ncclResult_t res = nccl().Recv(gathered_out + (size_t)r * count, count, ncclDouble, r, c->comm, s);
                    if (res != ncclSuccess) { nccl().GroupEnd(); return nccl_fail(res, "ncclRecv"); }
                }
            } else {
                ncclResult_t res = nccl().Send(local_sums, count, ncclDouble, root, c->comm, s);
                if (res != ncclSuccess) { nccl().GroupEnd(); return nccl_fail(res, "ncclSend"); }
            }
            NK(nccl().GroupEnd());
        }
    }
    if (!cuda_stream) CK(cudaStreamSynchronize(s));
    return RLB_OK;
}

rlb_status rlb_comm_allreduce_sum(rlb_comm* c, double* values, uint64_t count, void* cuda_stream) {
    if (!c || !values) { set_error("NULL argument"); return RLB_ERR_INVALID_ARG; }
    CK(cudaSetDevice(c->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : c->stream;
    if (count && c->world > 1) NK(nccl().AllReduce(values, values, (size_t)count, ncclDouble, ncclSum, c->comm, s));
    if (!cuda_stream) CK(cudaStreamSynchronize(s));
    return RLB_OK;
}

}   // extern "C"
