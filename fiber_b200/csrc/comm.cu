// comm.cu -- engine-level collectives for the one-process-per-GPU mode, behind the C ABI.
//
// The map itself needs no data-path collective (tasks are independent, every rank owns a contiguous
// block).  The exchange steps either side of it are real collectives and belong to the engine, not to
// the caller's framework:
//   * shared arguments (the parzen sample block, initargs) resident on one rank   -> ncclBroadcast
//   * a map's input array resident on one rank                                    -> grouped ncclSend/ncclRecv (scatter)
//   * the ordered result blocks of all ranks                                      -> ncclAllGather, or grouped send/recv to a root
//   * scalar folds (the pi count)                                                 -> ncclAllReduce(sum, int64)
//   * experimental.Ring's collective (fiber/experimental/ring.py:44-129 bootstraps a ring for
//     examples/ring.py:81-86's all_reduce)                                        -> ncclAllReduce(sum, float32)
// The bootstrap handle (ncclUniqueId, 128 bytes) is what a ring node publishes in the member table instead
// of the reference's ip/port pair.
//
// NCCL is bound at run time (dlopen): libfiber_b200 has no link-time dependency on it, and a pool that never
// builds a communicator never loads it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <string>

#include "../../include/fiber_b200.h"

namespace {

thread_local std::string g_cerr;
int cfail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_cerr = buf;
    return code;
}

// the slice of nccl.h this file uses (NCCL 2.x ABI)
typedef struct { char internal[128]; } ncclUniqueId;
typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;
enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5, ncclFloat16 = 6, ncclFloat32 = 7, ncclFloat64 = 8 };
enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 };

struct Nccl {
    void* handle = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
};
std::mutex g_nccl_mu;
Nccl g_nccl;

int load_nccl(const char* path) {
    std::lock_guard<std::mutex> g(g_nccl_mu);
    if (g_nccl.handle) return FBR_OK;
    const char* cands[] = {path, "libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    std::string tried;
    for (const char* c : cands) {
        if (!c || !*c) continue;
        h = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
        tried += std::string(c) + ": " + dlerror() + "; ";
    }
    if (!h) return cfail(FBR_ENOENT, "cannot load NCCL (%s)", tried.c_str());
#define SYM(field, name)                                                                   \
    *(void**)(&g_nccl.field) = dlsym(h, name);                                             \
    if (!g_nccl.field) { dlclose(h); return cfail(FBR_ENOENT, "NCCL symbol %s missing", name); }
    SYM(GetVersion, "ncclGetVersion")
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(CommAbort, "ncclCommAbort")
    SYM(GetErrorString, "ncclGetErrorString")
    SYM(Broadcast, "ncclBroadcast")
    SYM(AllReduce, "ncclAllReduce")
    SYM(AllGather, "ncclAllGather")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
#undef SYM
    g_nccl.handle = h;
    return FBR_OK;
}

#define NCK(call)                                                                                          \
    do {                                                                                                   \
        ncclResult_t r_ = (call);                                                                          \
        if (r_ != 0) return cfail(FBR_ECUDA, "%s failed: %s", #call, g_nccl.GetErrorString(r_));           \
    } while (0)
#define CCK(call)                                                                                          \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) return cfail(FBR_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_));    \
    } while (0)

int nccl_type(int dtype, size_t* elem) {
    switch (dtype) {
        case FBR_DT_U8: *elem = 1; return ncclUint8;
        case FBR_DT_I32: *elem = 4; return ncclInt32;
        case FBR_DT_I64: *elem = 8; return ncclInt64;
        case FBR_DT_F32: *elem = 4; return ncclFloat32;
        case FBR_DT_F64: *elem = 8; return ncclFloat64;
        default: *elem = 0; return -1;
    }
}

}  // namespace

struct fbr_comm {
    int device = 0, rank = 0, nranks = 1;
    ncclComm_t comm = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    long long* d_scalar = nullptr;      // device scratch for scalar folds
    long long* h_scalar = nullptr;      // pinned: [0] value in, [1] global sum out (truly asynchronous copies)
    cudaEvent_t ev_scalar = nullptr;
    bool scalar_pending = false;
};

extern "C" {

const char* fbr_comm_last_error(void) { return g_cerr.c_str(); }

int fbr_comm_load(const char* libnccl_path, int* version) {
    int rc = load_nccl(libnccl_path);
    if (rc != FBR_OK) return rc;
    if (version) NCK(g_nccl.GetVersion(version));
    return FBR_OK;
}

int fbr_comm_unique_id(void* id128) {
    if (!id128) return cfail(FBR_EINVAL, "NULL argument");
    int rc = load_nccl(nullptr);
    if (rc != FBR_OK) return rc;
    ncclUniqueId id;
    NCK(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return FBR_OK;
}

int fbr_comm_create(int device_id, int nranks, int rank, const void* id128, fbr_comm_t** out) {
    if (!out || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return cfail(FBR_EINVAL, "bad arguments");
    int rc = load_nccl(nullptr);
    if (rc != FBR_OK) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return cfail(FBR_ENODEV, "no CUDA device visible; collectives have no CPU fallback");
    }
    if (device_id < 0 || device_id >= ndev) return cfail(FBR_EINVAL, "device id %d out of range", device_id);
    fbr_comm* c = new fbr_comm();
    c->device = device_id;
    c->rank = rank;
    c->nranks = nranks;
    CCK(cudaSetDevice(device_id));
    CCK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CCK(cudaEventCreate(&c->ev0));
    CCK(cudaEventCreate(&c->ev1));
    CCK(cudaMalloc((void**)&c->d_scalar, 64));
    CCK(cudaHostAlloc((void**)&c->h_scalar, 64, cudaHostAllocPortable));
    CCK(cudaEventCreateWithFlags(&c->ev_scalar, cudaEventDisableTiming));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    NCK(g_nccl.CommInitRank(&c->comm, nranks, id, rank));
    *out = c;
    return FBR_OK;
}

int fbr_comm_info(fbr_comm_t* c, int* rank, int* nranks, int* device) {
    if (!c) return cfail(FBR_EINVAL, "NULL comm");
    if (rank) *rank = c->rank;
    if (nranks) *nranks = c->nranks;
    if (device) *device = c->device;
    return FBR_OK;
}

int fbr_comm_sync(fbr_comm_t* c) {
    if (!c) return cfail(FBR_EINVAL, "NULL comm");
    CCK(cudaSetDevice(c->device));
    CCK(cudaStreamSynchronize(c->stream));
    return FBR_OK;
}

/* shared arguments resident on `root` -> every rank (ncclBroadcast, in place) */
int fbr_comm_broadcast(fbr_comm_t* c, void* dptr, uint64_t bytes, int root) {
    if (!c || (!dptr && bytes)) return cfail(FBR_EINVAL, "bad arguments");
    CCK(cudaSetDevice(c->device));
    NCK(g_nccl.Broadcast(dptr, dptr, bytes, ncclUint8, root, c->comm, c->stream));
    return FBR_OK;
}

/* equal ordered result blocks of every rank -> the full ordered result on every rank */
int fbr_comm_allgather(fbr_comm_t* c, const void* send, void* recv, uint64_t bytes_per_rank) {
    if (!c || !send || !recv) return cfail(FBR_EINVAL, "bad arguments");
    CCK(cudaSetDevice(c->device));
    NCK(g_nccl.AllGather(send, recv, bytes_per_rank, ncclUint8, c->comm, c->stream));
    return FBR_OK;
}

/* every rank's block -> recv + r * bytes_per_rank on `root` (grouped send/recv: the fan-in of fiber/pool.py:916-920) */
int fbr_comm_gather(fbr_comm_t* c, const void* send, void* recv_on_root, uint64_t bytes_per_rank, int root) {
    if (!c || !send || (c->rank == root && !recv_on_root)) return cfail(FBR_EINVAL, "bad arguments");
    CCK(cudaSetDevice(c->device));
    NCK(g_nccl.GroupStart());
    if (c->rank == root)
        for (int r = 0; r < c->nranks; ++r)
            NCK(g_nccl.Recv((uint8_t*)recv_on_root + (uint64_t)r * bytes_per_rank, bytes_per_rank, ncclUint8, r, c->comm, c->stream));
    NCK(g_nccl.Send(send, bytes_per_rank, ncclUint8, root, c->comm, c->stream));
    NCK(g_nccl.GroupEnd());
    return FBR_OK;
}

/* send + r * bytes_per_rank on `root` -> every rank's block (the fan-out of fiber/pool.py:910-914) */
int fbr_comm_scatter(fbr_comm_t* c, const void* send_on_root, void* recv, uint64_t bytes_per_rank, int root) {
    if (!c || !recv || (c->rank == root && !send_on_root)) return cfail(FBR_EINVAL, "bad arguments");
    CCK(cudaSetDevice(c->device));
    NCK(g_nccl.GroupStart());
    if (c->rank == root)
        for (int r = 0; r < c->nranks; ++r)
            NCK(g_nccl.Send((const uint8_t*)send_on_root + (uint64_t)r * bytes_per_rank, bytes_per_rank, ncclUint8, r, c->comm, c->stream));
    NCK(g_nccl.Recv(recv, bytes_per_rank, ncclUint8, root, c->comm, c->stream));
    NCK(g_nccl.GroupEnd());
    return FBR_OK;
}

int fbr_comm_allreduce(fbr_comm_t* c, const void* send, void* recv, uint64_t count, int dtype, int op) {
    if (!c || !send || !recv) return cfail(FBR_EINVAL, "bad arguments");
    size_t elem = 0;
    const int nt = nccl_type(dtype, &elem);
    if (nt < 0 || op < FBR_OP_SUM || op > FBR_OP_MIN) return cfail(FBR_EINVAL, "unsupported dtype %d / op %d", dtype, op);
    CCK(cudaSetDevice(c->device));
    NCK(g_nccl.AllReduce(send, recv, count, nt, op, c->comm, c->stream));
    return FBR_OK;
}

/* `iters` back-to-back all-reduces timed with CUDA events on the communicator's stream (ms per call) */
int fbr_comm_allreduce_timed(fbr_comm_t* c, void* buf, uint64_t count, int dtype, int op, int iters, float* ms_per_call) {
    if (!c || !buf || iters < 1 || !ms_per_call) return cfail(FBR_EINVAL, "bad arguments");
    size_t elem = 0;
    const int nt = nccl_type(dtype, &elem);
    if (nt < 0) return cfail(FBR_EINVAL, "unsupported dtype %d", dtype);
    CCK(cudaSetDevice(c->device));
    CCK(cudaEventRecord(c->ev0, c->stream));
    for (int i = 0; i < iters; ++i) NCK(g_nccl.AllReduce(buf, buf, count, nt, op, c->comm, c->stream));
    CCK(cudaEventRecord(c->ev1, c->stream));
    CCK(cudaStreamSynchronize(c->stream));
    float ms = 0;
    CCK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    *ms_per_call = ms / iters;
    return FBR_OK;
}

/* Scalar fold of one int64 per rank (the pi count), split in two so that it overlaps the next map:
 * _begin enqueues H2D (pinned) + ncclAllReduce(sum, int64) + D2H (pinned) on the communicator's stream and returns;
 * _end waits for it and hands out the global sum.  One fold may be pending per communicator. */
int fbr_comm_allreduce_i64_begin(fbr_comm_t* c, int64_t value) {
    if (!c) return cfail(FBR_EINVAL, "NULL comm");
    if (c->scalar_pending) return cfail(FBR_ESTATE, "a scalar fold is already pending on this communicator");
    CCK(cudaSetDevice(c->device));
    c->h_scalar[0] = value;
    CCK(cudaMemcpyAsync(c->d_scalar, &c->h_scalar[0], sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
    NCK(g_nccl.AllReduce(c->d_scalar, c->d_scalar, 1, ncclInt64, ncclSum, c->comm, c->stream));
    CCK(cudaMemcpyAsync(&c->h_scalar[1], c->d_scalar, sizeof(int64_t), cudaMemcpyDeviceToHost, c->stream));
    CCK(cudaEventRecord(c->ev_scalar, c->stream));
    c->scalar_pending = true;
    return FBR_OK;
}
int fbr_comm_allreduce_i64_end(fbr_comm_t* c, int64_t* sum) {
    if (!c || !sum) return cfail(FBR_EINVAL, "bad arguments");
    if (!c->scalar_pending) return cfail(FBR_ESTATE, "no scalar fold pending");
    CCK(cudaSetDevice(c->device));
    CCK(cudaEventSynchronize(c->ev_scalar));
    c->scalar_pending = false;
    *sum = c->h_scalar[1];
    return FBR_OK;
}
/* host value in, global sum out */
int fbr_comm_allreduce_i64(fbr_comm_t* c, int64_t* value) {
    if (!c || !value) return cfail(FBR_EINVAL, "bad arguments");
    int rc = fbr_comm_allreduce_i64_begin(c, *value);
    if (rc != FBR_OK) return rc;
    return fbr_comm_allreduce_i64_end(c, value);
}

/* device buffers for callers that have no pool (ring nodes): plain cudaMalloc on the communicator's device */
int fbr_comm_device_alloc(fbr_comm_t* c, uint64_t bytes, void** dptr) {
    if (!c || !dptr) return cfail(FBR_EINVAL, "bad arguments");
    CCK(cudaSetDevice(c->device));
    CCK(cudaMalloc(dptr, bytes ? bytes : 1));
    return FBR_OK;
}
int fbr_comm_device_free(fbr_comm_t* c, void* dptr) {
    if (!c) return cfail(FBR_EINVAL, "NULL comm");
    CCK(cudaSetDevice(c->device));
    CCK(cudaFree(dptr));
    return FBR_OK;
}
int fbr_comm_memcpy_h2d(fbr_comm_t* c, void* dptr, const void* src, uint64_t bytes) {
    if (!c) return cfail(FBR_EINVAL, "NULL comm");
    CCK(cudaSetDevice(c->device));
    CCK(cudaMemcpyAsync(dptr, src, bytes, cudaMemcpyHostToDevice, c->stream));
    CCK(cudaStreamSynchronize(c->stream));
    return FBR_OK;
}
int fbr_comm_memcpy_d2h(fbr_comm_t* c, void* dst, const void* dptr, uint64_t bytes) {
    if (!c) return cfail(FBR_EINVAL, "NULL comm");
    CCK(cudaSetDevice(c->device));
    CCK(cudaMemcpyAsync(dst, dptr, bytes, cudaMemcpyDeviceToHost, c->stream));
    CCK(cudaStreamSynchronize(c->stream));
    return FBR_OK;
}

int fbr_comm_destroy(fbr_comm_t* c) {
    if (!c) return FBR_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    cudaFree(c->d_scalar);
    cudaFreeHost(c->h_scalar);
    if (c->ev_scalar) cudaEventDestroy(c->ev_scalar);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return FBR_OK;
}

}  // extern "C"
