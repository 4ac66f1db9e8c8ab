// NCCL resolved at run time (dlopen/dlsym) so that the library binds to the NCCL already loaded in the process
// (torch bundles its own libnccl.so.2) instead of pulling in a second copy. Only the handful of entry points the
// Schur-complement all-reduce needs (reduceKKT -> ncclAllReduce, Lsolve b0 -> ncclAllReduce).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstring>
#include <string>

namespace pipsgpu {

struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;

struct NcclApi {
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
    std::string err;

    static NcclApi& get() {
        static NcclApi api = load();
        return api;
    }

   private:
    static NcclApi load() {
        NcclApi a;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) {
            a.err = "cannot dlopen libnccl.so.2";
            return a;
        }
        a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
        a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
        a.AllReduce = (decltype(a.AllReduce))dlsym(h, "ncclAllReduce");
        a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.GetErrorString;
        if (!a.ok) a.err = "libnccl.so.2 lacks required symbols";
        return a;
    }
};

constexpr int kNcclFloat64 = 8;   // ncclDouble
constexpr int kNcclSum = 0;

}  // namespace pipsgpu
