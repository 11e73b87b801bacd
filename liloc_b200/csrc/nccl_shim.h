// liloc_b200/csrc/nccl_shim.h -- the five NCCL entry points the path needs, bound at run time.
//
// The only collective of the path is the final all-gather of the packed 6-DoF results: 8 floats per work
// item.  libliloc_gpu.so does not link NCCL: liloc_comm_init() dlopen()s libnccl.so.2 -- in a process that already loaded
// one (torch ships its own) the loader hands back that very library -- and binds the prototypes below, which are the stable
// NCCL 2.x C ABI (nccl.h: ncclGetUniqueId, ncclCommInitRank, ncclAllGather, ncclCommDestroy, ncclGetErrorString).
// Single-GPU callers never touch this file's code path.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <string>

namespace liloc {

struct NcclShim {
    typedef struct ncclComm* comm_t;
    struct unique_id { char internal[128]; };            // NCCL_UNIQUE_ID_BYTES
    enum { kFloat = 7 };                                 // ncclFloat32

    int (*GetUniqueId)(unique_id*) = nullptr;
    int (*CommInitRank)(comm_t*, int, unique_id, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, comm_t, cudaStream_t) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    void* handle = nullptr;

    // empty string on success, otherwise what went wrong
    std::string load() {
        if (handle) return "";
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        std::string err;
        for (const char* n : names) {
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
            const char* e = dlerror();
            err += std::string(n) + ": " + (e ? e : "not found") + "; ";
        }
        if (!handle) return "cannot load NCCL (" + err + ")";
        auto sym = [&](const char* s) { return dlsym(handle, s); };
        GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(sym("ncclGetUniqueId"));
        CommInitRank = reinterpret_cast<decltype(CommInitRank)>(sym("ncclCommInitRank"));
        AllGather = reinterpret_cast<decltype(AllGather)>(sym("ncclAllGather"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
        if (!GetUniqueId || !CommInitRank || !AllGather || !CommDestroy || !GetErrorString) { handle = nullptr; return "libnccl lacks an expected symbol"; }
        return "";
    }
};

inline NcclShim& nccl() { static NcclShim s; return s; }

}  // namespace liloc
