// express.cu -- doorbell path for single-task submissions (apply / apply_async).
//
// The reference pays one TCP round trip per apply_async (task chunk out, one result message back:
// fiber/pool.py:1089-1116, 760-825; tests/test_pool.py:247-270 does 5000 of them).  The wave
// pipeline of engine.cu costs two kernel launches and three small copies per map, ~30-50 us for a
// one-task map.  For bodies whose argument and result fit one 64-byte record this file keeps a
// *resident* one-warp kernel per worker that polls a pinned, device-mapped request lane (the
// doorbell), runs the body, and writes the result record back into a pinned response lane the host
// polls: no launch, no cudaMemcpy, no event on the round trip.
//
// The kernel never outlives its usefulness: after `idle_ns` without a request it exits (so a
// device-wide synchronisation elsewhere in the process, e.g. torch.cuda.synchronize(), waits at
// most that long) and the next request relaunches it.  The exit/relaunch race is closed with a
// three-state handshake (RUNNING -> EXITING -> re-check lane -> EXITED).
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <chrono>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>

#include "../../include/fiber_b200.h"
#include "bodies.cuh"

using namespace fbr;

static thread_local std::string x_err;
extern "C" const char* fbr_express_last_error(void) { return x_err.c_str(); }
static int xfail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    x_err = buf;
    return code;
}
#define XCK(call)                                                                                    \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) return xfail(FBR_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

constexpr uint32_t kXCap = 256;   // requests in flight (power of two)

struct alignas(64) XRequest {     // fixed-layout task record of one apply
    unsigned long long ticket;
    uint32_t func_id;
    uint32_t attempt;
    uint8_t arg[48];
};
struct alignas(64) XResponse {
    unsigned long long ticket;
    uint32_t err;                 // TaskError
    uint32_t result_bytes;
    uint8_t result[48];
};
enum { X_EXITED = 0, X_RUNNING = 1, X_EXITING = 2 };

struct alignas(64) XLanes {       // pinned + mapped
    volatile unsigned long long req_head;   // host
    char p0[56];
    volatile unsigned long long req_tail;   // device
    char p1[56];
    volatile unsigned long long rsp_head;   // device
    char p2[56];
    volatile int state;                     // X_*
    volatile int kill;
    volatile unsigned long long served;
    char p3[40];
    XRequest req[kXCap];
    XResponse rsp[kXCap];
};

__device__ __forceinline__ unsigned long long xtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <class B>
__device__ __forceinline__ void run_body(const XRequest& rq, XResponse& rs, unsigned long long* err_word, int* fault) {
    typename B::Arg a;
    memcpy(&a, rq.arg, sizeof a);
    const ErrSink es{err_word, fault};
    const typename B::Res r = B::run(a, 0, es, rq.attempt);
    memcpy(rs.result, &r, sizeof r);
    rs.result_bytes = sizeof r;
}

__global__ void __launch_bounds__(32) express_kernel(XLanes* L, unsigned long long* err_scratch, unsigned long long idle_ns) {
    if (threadIdx.x != 0) return;
    __shared__ int s_fault;
    unsigned long long last = xtime();
    for (;;) {
        const unsigned long long t = L->req_tail;
        const unsigned long long h = *(volatile unsigned long long*)&L->req_head;
        if (h != t) {
            __threadfence_system();
            XRequest rq;
            {
                const volatile uint4* src = (const volatile uint4*)&L->req[t & (kXCap - 1)];
                uint4* dst = (uint4*)&rq;
                for (int i = 0; i < 4; ++i) { uint4 v; v.x = src[i].x; v.y = src[i].y; v.z = src[i].z; v.w = src[i].w; dst[i] = v; }
            }
            XResponse rs;
            memset(&rs, 0, sizeof rs);
            rs.ticket = rq.ticket;
            *err_scratch = ~0ull;
            s_fault = 0;
            switch (rq.func_id) {
            case F_SQUARE_I64: run_body<SquareI64>(rq, rs, err_scratch, &s_fault); break;
            case F_MUL2_I64: run_body<Mul2I64>(rq, rs, err_scratch, &s_fault); break;
            case F_SQUARE_SCALE_I64: run_body<SquareScaleI64>(rq, rs, err_scratch, &s_fault); break;
            case F_IDENTITY_I64: run_body<IdentityI64>(rq, rs, err_scratch, &s_fault); break;
            case F_PI_INSIDE_DET: run_body<PiInsideDet>(rq, rs, err_scratch, &s_fault); break;
            case F_SLEEP_F64: run_body<SleepF64>(rq, rs, err_scratch, &s_fault); break;
            default: rs.err = TASK_BADARG; break;
            }
            if (*err_scratch != ~0ull) rs.err = (uint32_t)(*err_scratch & 0xff);
            if (s_fault) rs.err = TASK_FAULT;
            {
                volatile uint4* dst = (volatile uint4*)&L->rsp[t & (kXCap - 1)];
                const uint4* src = (const uint4*)&rs;
                for (int i = 0; i < 4; ++i) { dst[i].x = src[i].x; dst[i].y = src[i].y; dst[i].z = src[i].z; dst[i].w = src[i].w; }
            }
            __threadfence_system();
            L->rsp_head = t + 1;
            L->req_tail = t + 1;
            L->served = L->served + 1;
            last = xtime();
            continue;
        }
        if (L->kill || xtime() - last > idle_ns) {
            // exit handshake: announce, re-check the doorbell once, then leave
            L->state = X_EXITING;
            __threadfence_system();
            if (*(volatile unsigned long long*)&L->req_head != L->req_tail && !L->kill) {
                L->state = X_RUNNING;
                __threadfence_system();
                last = xtime();
                continue;
            }
            L->state = X_EXITED;
            __threadfence_system();
            return;
        }
        __nanosleep(500);
    }
}

struct fbr_express {
    int device = 0;
    cudaStream_t stream = nullptr;
    XLanes* lanes = nullptr;                 // pinned mapped
    unsigned long long* d_err = nullptr;     // device scratch for ErrSink
    std::mutex mu;
    unsigned long long next_ticket = 0;      // == req_head
    unsigned long long collected = 0;        // responses consumed from the lane
    std::unordered_map<unsigned long long, XResponse> parked;
    std::unordered_set<unsigned long long> abandoned;   // tickets whose handle was dropped before the response arrived
    unsigned long long idle_ns = 2000000ull; // 2 ms
    uint64_t launches = 0;
};

extern "C" int fbr_internal_preload(int device);   // engine.cu

static int ensure_running(fbr_express* x) {
    // called with x->mu held, after a request was published
    for (int spin = 0;; ++spin) {
        __atomic_thread_fence(__ATOMIC_SEQ_CST);
        const int st = x->lanes->state;
        if (st == X_RUNNING) return FBR_OK;
        if (st == X_EXITED) {
            XCK(cudaSetDevice(x->device));
            x->lanes->state = X_RUNNING;
            x->lanes->kill = 0;
            __atomic_thread_fence(__ATOMIC_SEQ_CST);
            express_kernel<<<1, 32, 0, x->stream>>>(x->lanes, x->d_err, x->idle_ns);
            XCK(cudaGetLastError());
            x->launches++;
            return FBR_OK;
        }
        // X_EXITING: the kernel is deciding; it resolves to RUNNING or EXITED within microseconds
        if (spin > 1000000) return xfail(FBR_ETIMEOUT, "express kernel stuck in exit handshake");
    }
}

extern "C" {

int fbr_express_create(int device_id, int idle_timeout_us, fbr_express_t** out) {
    if (!out) return xfail(FBR_EINVAL, "NULL argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return xfail(FBR_ENODEV, "no CUDA device visible; the express lane has no CPU fallback");
    }
    if (device_id < 0 || device_id >= ndev) return xfail(FBR_EINVAL, "device id %d out of range", device_id);
    fbr_express* x = new fbr_express();
    x->device = device_id;
    XCK(cudaSetDevice(device_id));
    fbr_internal_preload(device_id);
    {
        cudaFuncAttributes at;
        cudaFuncGetAttributes(&at, (const void*)express_kernel);
    }
    XCK(cudaStreamCreateWithFlags(&x->stream, cudaStreamNonBlocking));
    XCK(cudaHostAlloc((void**)&x->lanes, sizeof(XLanes), cudaHostAllocPortable | cudaHostAllocMapped));
    memset((void*)x->lanes, 0, sizeof(XLanes));
    XCK(cudaMalloc((void**)&x->d_err, sizeof(unsigned long long)));
    if (idle_timeout_us > 0) x->idle_ns = (unsigned long long)idle_timeout_us * 1000ull;
    *out = x;
    return FBR_OK;
}

/* Publish one apply; returns its ticket.  arg_bytes <= 48. */
int fbr_express_submit(fbr_express_t* x, int func_id, const void* arg, uint32_t arg_bytes, uint64_t* ticket) {
    if (!x || !ticket || (arg_bytes && !arg) || arg_bytes > 48) return xfail(FBR_EINVAL, "bad arguments");
    std::lock_guard<std::mutex> g(x->mu);
    XLanes* L = x->lanes;
    const auto t0 = std::chrono::steady_clock::now();
    while (x->next_ticket - x->collected >= kXCap) {
        // ring full of unconsumed responses: drain into the parked map
        const unsigned long long rh = __atomic_load_n(&L->rsp_head, __ATOMIC_ACQUIRE);
        if (x->collected < rh) {
            XResponse r;
            memcpy(&r, (const void*)&L->rsp[x->collected & (kXCap - 1)], sizeof r);
            if (!x->abandoned.erase(r.ticket)) x->parked[r.ticket] = r;
            x->collected++;
            continue;
        }
        if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(30)) return xfail(FBR_ETIMEOUT, "express lane stalled");
    }
    const unsigned long long t = x->next_ticket;
    XRequest rq;
    memset(&rq, 0, sizeof rq);
    rq.ticket = t;
    rq.func_id = (uint32_t)func_id;
    memcpy(rq.arg, arg, arg_bytes);
    memcpy((void*)&L->req[t & (kXCap - 1)], &rq, sizeof rq);
    __atomic_store_n(&L->req_head, t + 1, __ATOMIC_RELEASE);
    x->next_ticket = t + 1;
    *ticket = t;
    return ensure_running(x);
}

/* Wait for the response of `ticket`; copies result_bytes (<= 48) into `result`. */
int fbr_express_wait(fbr_express_t* x, uint64_t ticket, void* result, uint32_t* result_bytes, uint32_t* err, int timeout_ms) {
    if (!x) return xfail(FBR_EINVAL, "NULL argument");
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(timeout_ms < 0 ? 0 : timeout_ms);
    XLanes* L = x->lanes;
    int spins = 0;
    for (;;) {
        {
            std::lock_guard<std::mutex> g(x->mu);
            auto it = x->parked.find(ticket);
            if (it == x->parked.end()) {
                const unsigned long long rh = __atomic_load_n(&L->rsp_head, __ATOMIC_ACQUIRE);
                while (x->collected < rh) {
                    XResponse r;
                    memcpy(&r, (const void*)&L->rsp[x->collected & (kXCap - 1)], sizeof r);
                    x->collected++;
                    if (!x->abandoned.erase(r.ticket)) x->parked[r.ticket] = r;
                }
                it = x->parked.find(ticket);
            }
            if (it != x->parked.end()) {
                const XResponse& r = it->second;
                if (result && r.result_bytes) memcpy(result, r.result, r.result_bytes);
                if (result_bytes) *result_bytes = r.result_bytes;
                if (err) *err = r.err;
                const uint32_t e = r.err;
                x->parked.erase(it);
                if (e) return xfail(FBR_ETASK, "express task %llu failed with code %u", (unsigned long long)ticket, e);
                return FBR_OK;
            }
            if (ticket >= x->next_ticket) return xfail(FBR_ENOENT, "unknown express ticket");
            // a request is pending but the kernel may have exited between submit and now
            int rc = ensure_running(x);
            if (rc) return rc;
        }
        if (timeout_ms >= 0 && std::chrono::steady_clock::now() >= deadline) return xfail(FBR_ETIMEOUT, "express response timeout");
        if (++spins > 20000) std::this_thread::sleep_for(std::chrono::microseconds(20));
    }
}

/* The caller dropped its handle on `ticket` without waiting: forget the response (now or when it arrives). */
int fbr_express_discard(fbr_express_t* x, uint64_t ticket) {
    if (!x) return xfail(FBR_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> g(x->mu);
    if (x->parked.erase(ticket) == 0 && ticket < x->next_ticket) x->abandoned.insert(ticket);
    return FBR_OK;
}

int fbr_express_stats(fbr_express_t* x, uint64_t* served, uint64_t* launches, int* resident) {
    if (!x) return xfail(FBR_EINVAL, "NULL argument");
    if (served) *served = x->lanes->served;
    if (launches) *launches = x->launches;
    if (resident) *resident = x->lanes->state != X_EXITED;
    return FBR_OK;
}

int fbr_express_destroy(fbr_express_t* x) {
    if (!x) return FBR_OK;
    {
        std::lock_guard<std::mutex> g(x->mu);
        x->lanes->kill = 1;
        __atomic_thread_fence(__ATOMIC_SEQ_CST);
    }
    cudaSetDevice(x->device);
    cudaStreamSynchronize(x->stream);
    cudaStreamDestroy(x->stream);
    cudaFree(x->d_err);
    // the pinned lanes are left to process exit (cudaFreeHost may synchronise other resident kernels)
    delete x;
    return FBR_OK;
}

}  // extern "C"
