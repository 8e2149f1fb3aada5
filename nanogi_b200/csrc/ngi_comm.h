// ngi_comm.h — NCCL, loaded at run time, for the one exchange step of the path.
//
// The reference gathers its per-thread films with `film[i] += ctx.film[i] * (W*H/N)` (src/nanogi.cpp:429-437). Across GPUs the
// same gather is ONE ncclReduce(SUM) of the per-GPU films over NVLink; a scene built once on the first device is
// handed to the others with ncclBroadcast. Nothing else of the path communicates.
//
// libnccl.so.2 is opened with dlopen on first use, so that libnanogi_gpu.so has no link-time dependency on it: a single-GPU host
// without NCCL still loads the module, and inside a PyTorch process the copy PyTorch already mapped (same soname) is the one used.
// NGI_NCCL_LIB names another library file.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <mutex>
#include <string>

struct NgiNccl {
    void* handle = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

inline NgiNccl* ngi_nccl() {
    static NgiNccl api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[3] = {getenv("NGI_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            if (!nm || !*nm) continue;
            api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) { api.error = std::string("NCCL not available: ") + (dlerror() ? dlerror() : "libnccl.so.2 not found"); return; }
        bool ok = true;
        auto sym = [&](const char* s) { void* p = dlsym(api.handle, s); if (!p) { ok = false; api.error = std::string("NCCL symbol missing: ") + s; } return p; };
        api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.Reduce = (decltype(api.Reduce))sym("ncclReduce");
        api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        if (!ok) { dlclose(api.handle); api.handle = nullptr; }
    });
    return &api;
}
