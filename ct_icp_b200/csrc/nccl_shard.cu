// nccl_shard.cu — keypoint-sharded multi-GPU mode.
//
// One process per GPU. Every rank registers the same scan against its own replica of the map (sampling and map
// update are deterministic, so the replicas stay identical); the K keypoints of a frame are split into contiguous
// chunks [K r / G, K (r+1) / G) and each Gauss-Newton iteration performs ONE collective: an in-place ncclAllReduce
// (sum) of the 96-double accumulator (78 JTJ + 12 JTr + counters, 768 bytes) over NVLink/NVSwitch, after which every
// rank solves the same 12x12 system and applies the same pose update — no broadcast needed.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): the process usually already holds torch's bundled NCCL, and a
// link-time dependency on the system copy would put two NCCL versions behind the same symbols.
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>

#include "engine.h"
#include "icp.h"

namespace cticp {

struct NcclError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

namespace {
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi &Api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) return;
        api.GetUniqueId = (decltype(api.GetUniqueId)) dlsym(api.handle, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank)) dlsym(api.handle, "ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy)) dlsym(api.handle, "ncclCommDestroy");
        api.AllReduce = (decltype(api.AllReduce)) dlsym(api.handle, "ncclAllReduce");
        api.GetErrorString = (decltype(api.GetErrorString)) dlsym(api.handle, "ncclGetErrorString");
    });
    if (!api.handle || !api.GetUniqueId || !api.CommInitRank || !api.AllReduce)
        throw NcclError("NCCL: libnccl.so.2 could not be loaded");
    return api;
}
void Check(ncclResult_t r, const char *what) {
    if (r != ncclSuccess) {
        const char *msg = Api().GetErrorString ? Api().GetErrorString(r) : "?";
        throw NcclError(std::string("NCCL: ") + what + ": " + msg);
    }
}
}  // namespace

void IcpSolver::AllReduceAccumulator(void *nccl_comm) {
    Check(Api().AllReduce(d_acc_, d_acc_, kAcc, ncclDouble, ncclSum, (ncclComm_t) nccl_comm, stream_), "ncclAllReduce");
}

void Engine::EnableSharding(const void *unique_id, int rank, int world) {
    if (!unique_id || world < 1 || rank < 0 || rank >= world) throw std::invalid_argument("EnableSharding arguments");
    if (cudaSetDevice(device_) != cudaSuccess) throw CudaError("cudaSetDevice");
    if (nccl_comm_) {
        Api().CommDestroy((ncclComm_t) nccl_comm_);
        nccl_comm_ = nullptr;
    }
    shard_rank_ = rank;
    shard_world_ = world;
    if (world == 1) return;
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(&id, unique_id, sizeof(id));
    ncclComm_t comm;
    Check(Api().CommInitRank(&comm, world, id, rank), "ncclCommInitRank");
    nccl_comm_ = comm;
}

void Engine::DestroySharding() {
    if (nccl_comm_) {
        try {
            Api().CommDestroy((ncclComm_t) nccl_comm_);
        } catch (...) {
        }
        nccl_comm_ = nullptr;
    }
}

}  // namespace cticp

extern "C" int cticp_nccl_unique_id(void *out_128_bytes) {
    try {
        ncclUniqueId id;
        cticp::Check(cticp::Api().GetUniqueId(&id), "ncclGetUniqueId");
        memcpy(out_128_bytes, &id, sizeof(id));
        return CTICP_OK;
    } catch (const std::exception &) {
        return CTICP_ERR_NCCL;
    }
}
