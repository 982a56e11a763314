// nccl_dyn.h — NCCL bound at run time with dlopen, so liblamcg.so has no link-time NCCL
// dependency: inside a PyTorch process the already-loaded bundled libnccl.so.2 is reused (one NCCL
// per process), the stand-alone CLI picks up the system library.  Only the handful of entry points
// the CG loop needs are bound.  Types come from <nccl.h> (compile time only).
#pragma once

#include <dlfcn.h>
#include <nccl.h>

namespace lamcgk {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;

    bool load(const char **why)
    {
        if (handle) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) {
            *why = "libnccl.so.2 not found (dlopen)";
            return false;
        }
#define LAMCG_BIND(field, sym)                                        \
    field = reinterpret_cast<decltype(field)>(dlsym(handle, sym));   \
    if (!field) { *why = "missing NCCL symbol " sym; return false; }
        LAMCG_BIND(GetUniqueId, "ncclGetUniqueId")
        LAMCG_BIND(CommInitRank, "ncclCommInitRank")
        LAMCG_BIND(CommDestroy, "ncclCommDestroy")
        LAMCG_BIND(AllReduce, "ncclAllReduce")
        LAMCG_BIND(AllGather, "ncclAllGather")
        LAMCG_BIND(Broadcast, "ncclBroadcast")
        LAMCG_BIND(GetErrorString, "ncclGetErrorString")
        LAMCG_BIND(GetVersion, "ncclGetVersion")
#undef LAMCG_BIND
        return true;
    }
};

inline NcclApi &nccl_api()
{
    static NcclApi api;
    return api;
}

} // namespace lamcgk
