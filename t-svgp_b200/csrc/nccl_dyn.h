// Minimal run-time binding of NCCL (dlopen), so that libtsvgp.so has no link-time dependency on it: a single-GPU
// context never touches NCCL, and a multi-GPU one binds whichever libnccl.so.2 the process already holds (the
// torch-bundled 2.28 when bench.py imported torch for its rendezvous, else the system 2.27).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>

namespace tsvgp {

struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm* NcclComm;
enum { NCCL_INT32 = 2, NCCL_FLOAT64 = 8, NCCL_SUM = 0 };   // ncclInt / ncclDouble / ncclSum in nccl.h (stable since NCCL 2.0)

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*ReduceScatter)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;   // recvcount per rank
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;            // sendcount per rank
    int (*Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;   // send, recv, count, type, root
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool bcast() const { return Broadcast && GroupStart && GroupEnd; }
    bool ok() const { return handle && GetUniqueId && CommInitRank && CommDestroy && AllReduce; }
};

inline NcclApi& nccl_api() {
    static NcclApi api;
    if (api.handle) return api;
    const char* names[] = {"libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
    for (const char* n : names) {
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) return api;
    api.GetUniqueId = (int (*)(NcclUniqueId*))dlsym(api.handle, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))dlsym(api.handle, "ncclCommInitRank");
    api.CommDestroy = (int (*)(NcclComm))dlsym(api.handle, "ncclCommDestroy");
    api.AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(api.handle, "ncclAllReduce");
    api.ReduceScatter = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(api.handle, "ncclReduceScatter");
    api.AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t))dlsym(api.handle, "ncclAllGather");
    api.Broadcast = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(api.handle, "ncclBroadcast");
    api.GroupStart = (int (*)())dlsym(api.handle, "ncclGroupStart");
    api.GroupEnd = (int (*)())dlsym(api.handle, "ncclGroupEnd");
    api.GetErrorString = (const char* (*)(int))dlsym(api.handle, "ncclGetErrorString");
    return api;
}

}  // namespace tsvgp
