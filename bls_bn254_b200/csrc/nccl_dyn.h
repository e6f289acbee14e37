// nccl_dyn.h -- NCCL bound at run time (dlopen of libnccl.so.2), host code only.
//
// The aggregate-verification path has the library's single exchange step: one all-gather of a 400-byte record per
// rank.  libnccl is resolved with dlopen instead of at link time so that libbls_bn254_b200.so loads
// next to whatever NCCL the process already carries (PyTorch ships its own libnccl.so.2; a second copy with another
// version in the same process is what a link-time dependency risks) and so that single-GPU users need no NCCL at all.
// Only the handful of entry points below are used; their signatures are the stable NCCL 2.x C API.
#pragma once
#include <dlfcn.h>
#include <stddef.h>

#include <cuda_runtime.h>

namespace bn_nccl {

typedef struct ncclComm *comm_t;
typedef struct { char internal[128]; } unique_id;   // NCCL_UNIQUE_ID_BYTES
enum { kSuccess = 0, kUint8 = 1 /* ncclUint8 */ };

struct Api {
    void *handle = nullptr;
    int (*GetVersion)(int *) = nullptr;
    int (*GetUniqueId)(unique_id *) = nullptr;
    int (*CommInitRank)(comm_t *, int, unique_id, int) = nullptr;
    int (*CommInitAll)(comm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    const char *why = nullptr;   // set when the library could not be bound
    bool ok() const { return handle != nullptr && why == nullptr; }
};

inline const Api &api() {
    static Api a = [] {
        Api x;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            x.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (x.handle) break;
        }
        if (!x.handle) {
            x.why = "libnccl.so.2 not found (dlopen)";
            return x;
        }
#define BN_NCCL_SYM(field, sym)                                    \
    x.field = reinterpret_cast<decltype(x.field)>(dlsym(x.handle, sym)); \
    if (!x.field) x.why = "libnccl lacks " sym;
        BN_NCCL_SYM(GetVersion, "ncclGetVersion")
        BN_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        BN_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        BN_NCCL_SYM(CommInitAll, "ncclCommInitAll")
        BN_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        BN_NCCL_SYM(AllGather, "ncclAllGather")
        BN_NCCL_SYM(GroupStart, "ncclGroupStart")
        BN_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        BN_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef BN_NCCL_SYM
        return x;
    }();
    return a;
}

}  // namespace bn_nccl
