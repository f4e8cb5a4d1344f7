// jw_nccl.cuh -- NCCL through dlopen: the library has no link-time NCCL dependency, and inside a
// process that already loaded an NCCL (e.g. torch's bundled one) the same instance is reused
// (same soname).  Used only by the row-sharded multi-GPU sweep: set-up all-reduces of the integer marker / pair counts, one
// all-gather of the ycorr shards per sweep, and (engine 0 only) one all-reduce of the exact int64 block rhs per block.
#pragma once
#include <dlfcn.h>
#include "jw_common.cuh"

#define JW_NCCL_UNIQUE_ID_BYTES 128
struct jw_nccl_id { char internal[JW_NCCL_UNIQUE_ID_BYTES]; };
typedef void* jw_nccl_comm;

struct jw_nccl_api {
    void* lib = nullptr;
    int (*GetUniqueId)(jw_nccl_id*) = nullptr;
    int (*CommInitRank)(jw_nccl_comm*, int, jw_nccl_id, int) = nullptr;
    int (*CommDestroy)(jw_nccl_comm) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, jw_nccl_comm, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, jw_nccl_comm, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, jw_nccl_comm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

static jw_nccl_api* jw_nccl() {
    static jw_nccl_api api;
    static bool tried = false;
    if (tried) return api.lib ? &api : nullptr;
    tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    for (int i = 0; names[i] && !api.lib; ++i) api.lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) return nullptr;
#define JW_SYM(field, name) *(void**)(&api.field) = dlsym(api.lib, name); if (!api.field) { api.lib = nullptr; return nullptr; }
    JW_SYM(GetUniqueId, "ncclGetUniqueId")
    JW_SYM(CommInitRank, "ncclCommInitRank")
    JW_SYM(CommDestroy, "ncclCommDestroy")
    JW_SYM(AllReduce, "ncclAllReduce")
    JW_SYM(Broadcast, "ncclBroadcast")
    JW_SYM(AllGather, "ncclAllGather")
    JW_SYM(GroupStart, "ncclGroupStart")
    JW_SYM(GroupEnd, "ncclGroupEnd")
    JW_SYM(GetErrorString, "ncclGetErrorString")
#undef JW_SYM
    return &api;
}

#define JW_NCCL_INT32 2
#define JW_NCCL_INT64 4
#define JW_NCCL_FLOAT32 7
#define JW_NCCL_SUM 0

#define JW_NCCL(call)                                                                   \
    do {                                                                                \
        int r__ = (call);                                                               \
        if (r__ != 0) {                                                                 \
            jw_set_error(std::string(#call) + ": " + jw_nccl()->GetErrorString(r__));   \
            return 13;                                                                  \
        }                                                                               \
    } while (0)
