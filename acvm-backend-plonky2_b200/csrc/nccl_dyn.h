// NCCL entry points resolved at run time (dlopen), so libp2g.so has no load-time dependency on libnccl: single-GPU users never
// need it, and inside a PyTorch process the library torch already loaded (its bundled libnccl.so.2) is the one that is used.
// Only the handful of calls the coset-sharded prover makes (include/p2g.h p2g_circuit_create_sharded_nccl).
#pragma once
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>

#include <mutex>
#include <string>

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
    ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t*) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    std::string origin;   // which file was opened
    bool ok = false;
};

inline const NcclApi& nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = nullptr;
        // 1. a copy that is already in the process (PyTorch's bundled libnccl.so.2, or one the host preloaded): two NCCL builds
        //    under one SONAME cannot coexist, so never load a second one next to it; 2. P2G_NCCL_LIB; 3. the system library.
        const char* env = getenv("P2G_NCCL_LIB");
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (h) api.origin = "libnccl.so.2 (already loaded)";
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            if (h) break;
            if (!nm) continue;
            h = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
            if (h) api.origin = nm;
        }
        if (!h) return;
        bool all = true;
        auto sym = [&](const char* s) {
            void* p = dlsym(h, s);
            if (!p) all = false;
            return p;
        };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.CommAbort = (decltype(api.CommAbort))sym("ncclCommAbort");
        api.CommGetAsyncError = (decltype(api.CommGetAsyncError))sym("ncclCommGetAsyncError");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
        api.ok = all;
    });
    return api;
}
