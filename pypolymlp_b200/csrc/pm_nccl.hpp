// pm_nccl.hpp -- NCCL entry points resolved at run time (dlopen of libnccl.so.2), so that the library has no link-time
// dependency on NCCL: single-GPU users never load it, multi-GPU users get whatever NCCL the box provides (>= 2.4).
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <stdexcept>
#include <string>

namespace pm {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

inline NcclApi& nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* names[] = {getenv("POLYMLP_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (api.handle) {
            auto sym = [&](const char* s) { return dlsym(api.handle, s); };
            api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
            api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
            api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
            api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
            api.Reduce = (decltype(api.Reduce))sym("ncclReduce");
            api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
            api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
            api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
            api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        }
    }
    if (!api.handle || !api.GetUniqueId || !api.CommInitRank || !api.CommInitAll || !api.Reduce || !api.AllReduce ||
        !api.GroupStart || !api.GroupEnd || !api.CommDestroy)
        throw std::runtime_error("NCCL is not available (libnccl.so.2 could not be loaded): multi-GPU reduction needs it");
    return api;
}

inline void nccl_check(ncclResult_t r, const char* what) {
    if (r == ncclSuccess) return;
    NcclApi& a = nccl_api();
    throw std::runtime_error(std::string(what) + ": " + (a.GetErrorString ? a.GetErrorString(r) : "NCCL error"));
}

}  // namespace pm
