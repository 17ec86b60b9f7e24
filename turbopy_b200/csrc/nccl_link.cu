// The one collective of the path: an in-place SUM all-reduce of the diagnostic's <= 8 doubles over
// NCCL (NVLink / NVSwitch), enqueued ON THE CALLER'S STREAM -- the compute stream the reduction kernel ran on -- so
// that `[K x gather_push, reduce_moments, all-reduce]` is one stream-ordered sequence that a CUDA graph can capture
// and replay.  torch.distributed.all_reduce cannot give that: it runs on ProcessGroupNCCL's own stream behind
// event waits.
//
// libnccl is resolved at run time (dlopen): the process normally has torch's bundled libnccl.so.2 loaded already
// and this file binds to that copy (RTLD_NOLOAD first), so no second NCCL is initialised and the library itself
// carries no link-time dependency on NCCL (single-GPU users never touch it).  Only the five entry points below
// are used; the few NCCL types are restated here (stable ABI since NCCL 2.0).
#include "common.cuh"

#include <dlfcn.h>

#include <cstring>
#include <mutex>

namespace {

typedef struct { char internal[128]; } nccl_unique_id;            // NCCL_UNIQUE_ID_BYTES = 128
typedef void* nccl_comm;
constexpr int kNcclFloat64 = 8;                                  // ncclDataType_t: ncclFloat64 / ncclDouble
constexpr int kNcclSum = 0;                                      // ncclRedOp_t: ncclSum

struct NcclApi {
    void* handle = nullptr;
    int (*get_unique_id)(nccl_unique_id*) = nullptr;
    int (*comm_init_rank)(nccl_comm*, int, nccl_unique_id, int) = nullptr;
    int (*all_reduce)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
    int (*comm_destroy)(nccl_comm) = nullptr;
    int (*get_version)(int*) = nullptr;
    const char* (*get_error_string)(int) = nullptr;
    bool ok = false;
};

NcclApi& api() {
    static NcclApi a;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {                            // the copy already in the process (torch's), if any
            a.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD);
            if (a.handle) break;
        }
        if (!a.handle)
            for (const char* n : names) {
                a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
                if (a.handle) break;
            }
        if (!a.handle) return;
        a.get_unique_id = reinterpret_cast<decltype(a.get_unique_id)>(dlsym(a.handle, "ncclGetUniqueId"));
        a.comm_init_rank = reinterpret_cast<decltype(a.comm_init_rank)>(dlsym(a.handle, "ncclCommInitRank"));
        a.all_reduce = reinterpret_cast<decltype(a.all_reduce)>(dlsym(a.handle, "ncclAllReduce"));
        a.comm_destroy = reinterpret_cast<decltype(a.comm_destroy)>(dlsym(a.handle, "ncclCommDestroy"));
        a.get_version = reinterpret_cast<decltype(a.get_version)>(dlsym(a.handle, "ncclGetVersion"));
        a.get_error_string = reinterpret_cast<decltype(a.get_error_string)>(dlsym(a.handle, "ncclGetErrorString"));
        a.ok = a.get_unique_id && a.comm_init_rank && a.all_reduce && a.comm_destroy;
    });
    return a;
}

}  // namespace

// NCCL's own result codes are positive small integers like cudaError_t; they are passed through shifted into a
// private range so that tp_error_string can tell them apart.
#define TP_NCCL_TRY(expr)                                   \
    do {                                                    \
        const int _r = (expr);                              \
        if (_r != 0) return TP_E_NCCL - _r;                 \
    } while (0)

extern "C" int tp_nccl_version(int* version) {
    NcclApi& a = api();
    if (!a.ok) return TP_E_NCCL;
    if (!version) return TP_E_ARG;
    *version = 0;
    if (a.get_version) TP_NCCL_TRY(a.get_version(version));
    return TP_OK;
}

extern "C" int tp_nccl_unique_id(void* id128) {
    NcclApi& a = api();
    if (!a.ok) return TP_E_NCCL;
    if (!id128) return TP_E_ARG;
    nccl_unique_id id;
    TP_NCCL_TRY(a.get_unique_id(&id));
    std::memcpy(id128, id.internal, sizeof(id.internal));
    return TP_OK;
}

extern "C" int tp_nccl_comm_init(const void* id128, int world, int rank, void** comm) {
    NcclApi& a = api();
    if (!a.ok) return TP_E_NCCL;
    if (!id128 || !comm || world < 1 || rank < 0 || rank >= world) return TP_E_ARG;
    const tp::DeviceProps& d = tp::device_props();
    if (d.status != TP_OK) return d.status;
    nccl_unique_id id;
    std::memcpy(id.internal, id128, sizeof(id.internal));
    nccl_comm c = nullptr;
    TP_NCCL_TRY(a.comm_init_rank(&c, world, id, rank));
    *comm = c;
    return TP_OK;
}

extern "C" int tp_nccl_allreduce_sum_f64(void* comm, double* buf, int64_t n, tp_stream_t stream) {
    NcclApi& a = api();
    if (!a.ok) return TP_E_NCCL;
    if (!comm || !buf || n < 0) return TP_E_ARG;
    if (n == 0) return TP_OK;
    TP_NCCL_TRY(a.all_reduce(buf, buf, static_cast<size_t>(n), kNcclFloat64, kNcclSum, comm,
                             static_cast<cudaStream_t>(stream)));
    return TP_OK;
}

extern "C" int tp_nccl_comm_destroy(void* comm) {
    NcclApi& a = api();
    if (!a.ok) return TP_E_NCCL;
    if (!comm) return TP_OK;
    TP_NCCL_TRY(a.comm_destroy(comm));
    return TP_OK;
}
