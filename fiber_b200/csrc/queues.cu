// queues.cu -- SimpleQueue / Pipe / device Process on pinned rings.
//
// Reference: fiber/queues.py:262-352 (Pipe, SimpleQueuePush) over fiber/socket.py:297-366,416-425
// (a forwarder "device" thread running nn_device between a PULL and a PUSH socket: writers are
// fair-queued in, readers are load-balanced round-robin out -- tests/test_queue.py:218-250 pins
// exactly 600 of 2400 messages per reader).
//
// Here every endpoint owns one SPSC *lane*: a ring of 64-byte records in pinned, device-mapped host
// memory with a producer-written head and a consumer-written tail.  A queue is a set of writer lanes
// and reader lanes plus the forwarder (one host thread for all queues) that moves records from
// writer lanes (fair) to reader lanes (strict round-robin, skipping only full lanes).  An endpoint
// is either the host (fbr_lane_send/recv) or a *device process*: a resident one-warp kernel bound to
// a GPU that polls its lanes through the mapped pointers -- the GPU analogue of a job-backed
// fiber.Process (fiber/process.py:83-323) running one of the reference tests' target functions.
// Device processes carry a kill flag and an idle watchdog so a forgotten reader can never hang the
// GPU.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/fiber_b200.h"

static thread_local std::string q_err;
extern "C" const char* fbr_queue_last_error(void) { return q_err.c_str(); }
static int qfail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    q_err = buf;
    return code;
}
#define QCK(call)                                                                                    \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) return qfail(FBR_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

// ------------------------------------------------------------------------------------------------
// lanes
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kLaneCapacity = 1024;   // records (power of two)

struct alignas(64) Lane {
    volatile unsigned long long head;      // written by the producer
    char pad0[56];
    volatile unsigned long long tail;      // written by the consumer
    char pad1[56];
    fbr_record_t rec[kLaneCapacity];
};

// host side (the other end may be a GPU: publish/observe with full fences)
static bool lane_push_host(Lane* l, const fbr_record_t* r) {
    const unsigned long long h = l->head, t = __atomic_load_n(&l->tail, __ATOMIC_ACQUIRE);
    if (h - t >= kLaneCapacity) return false;
    memcpy((void*)&l->rec[h & (kLaneCapacity - 1)], r, sizeof *r);
    __atomic_store_n(&l->head, h + 1, __ATOMIC_RELEASE);
    return true;
}
static bool lane_pop_host(Lane* l, fbr_record_t* r) {
    const unsigned long long t = l->tail, h = __atomic_load_n(&l->head, __ATOMIC_ACQUIRE);
    if (t == h) return false;
    memcpy(r, (const void*)&l->rec[t & (kLaneCapacity - 1)], sizeof *r);
    __atomic_store_n(&l->tail, t + 1, __ATOMIC_RELEASE);
    return true;
}
static bool lane_has_space(Lane* l) { return l->head - __atomic_load_n(&l->tail, __ATOMIC_ACQUIRE) < kLaneCapacity; }

// ------------------------------------------------------------------------------------------------
// queues + the forwarder thread (the reference's ProcessDevice / nn_device)
// ------------------------------------------------------------------------------------------------
struct fbr_lane {
    Lane* ring = nullptr;     // pinned, device-mapped (UVA: same pointer on the device)
    struct fbr_queue* q = nullptr;
    bool is_writer = false;
    std::mutex mu;            // several host threads may share one host endpoint
};

struct fbr_queue {
    std::mutex mu;
    std::vector<fbr_lane*> writers, readers;
    size_t rr_in = 0, rr_out = 0;
    fbr_lane* host_writer = nullptr;
    fbr_lane* host_reader = nullptr;
    uint64_t forwarded = 0;
    bool closed = false;
};

struct Hub {
    std::mutex mu;
    std::vector<fbr_queue*> queues;
    std::thread thr;
    std::atomic<bool> stop{false};
    bool started = false;

    void run() {
        while (!stop.load(std::memory_order_relaxed)) {
            bool moved = false;
            std::vector<fbr_queue*> qs;
            {
                std::lock_guard<std::mutex> g(mu);
                qs = queues;
            }
            for (fbr_queue* q : qs) {
                std::lock_guard<std::mutex> g(q->mu);
                if (q->closed || q->readers.empty()) continue;
                const size_t nw = q->writers.size();
                for (size_t k = 0; k < nw; ++k) {
                    fbr_lane* w = q->writers[(q->rr_in + k) % nw];
                    for (int burst = 0; burst < 64; ++burst) {       // fair queueing: bounded burst per writer
                        if (w->ring->tail == __atomic_load_n(&w->ring->head, __ATOMIC_ACQUIRE)) break;
                        // strict round-robin over the readers, skipping only full lanes
                        fbr_lane* dst = nullptr;
                        const size_t nr = q->readers.size();
                        for (size_t j = 0; j < nr; ++j) {
                            fbr_lane* r = q->readers[(q->rr_out + j) % nr];
                            if (lane_has_space(r->ring)) { dst = r; q->rr_out = (q->rr_out + j + 1) % nr; break; }
                        }
                        if (!dst) break;
                        fbr_record_t rec;
                        lane_pop_host(w->ring, &rec);
                        lane_push_host(dst->ring, &rec);
                        q->forwarded++;
                        moved = true;
                    }
                }
                if (nw) q->rr_in = (q->rr_in + 1) % nw;
            }
            if (!moved) std::this_thread::sleep_for(std::chrono::microseconds(20));
        }
    }
    void ensure_started() {
        std::lock_guard<std::mutex> g(mu);
        if (!started) {
            started = true;
            thr = std::thread([this] { run(); });
            thr.detach();
        }
    }
};
static Hub& hub() {
    static Hub* h = new Hub();   // intentionally leaked: the detached forwarder may outlive static dtors
    return *h;
}

static int lane_create(fbr_queue* q, bool writer, fbr_lane** out) {
    std::unique_ptr<fbr_lane> l(new fbr_lane());
    // pinned + device-mapped so a GPU endpoint can poll it; on a host without a CUDA device the lane
    // is ordinary memory (host<->host endpoints still work; device processes fail with FBR_ENODEV)
    if (cudaHostAlloc((void**)&l->ring, sizeof(Lane), cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
        cudaGetLastError();
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0) return qfail(FBR_ENOMEM, "cudaHostAlloc of a queue lane failed");
        cudaGetLastError();
        if (posix_memalign((void**)&l->ring, 64, sizeof(Lane)) != 0) return qfail(FBR_ENOMEM, "out of memory");
    }
    l->ring->head = 0;
    l->ring->tail = 0;
    l->q = q;
    l->is_writer = writer;
    {
        std::lock_guard<std::mutex> g(q->mu);
        (writer ? q->writers : q->readers).push_back(l.get());
    }
    *out = l.release();
    return FBR_OK;
}

// ------------------------------------------------------------------------------------------------
// device side
// ------------------------------------------------------------------------------------------------
struct ProcCtrl {                 // pinned, mapped: host <-> device process control block
    volatile int kill;            // host sets 1: terminate()
    volatile int state;           // 0 not started, 1 running, 2 exited
    volatile int exitcode;        // 0 ok, -15 terminated (SIGTERM-like), 3 idle watchdog
    volatile unsigned long long handled;
};

struct ProcArgs {
    int kind;
    Lane* in;                     // lane this process reads (may be null)
    Lane* out;                    // lane this process writes (may be null)
    long long ident;              // queue_worker: what to put; get_queue: n; put_queue: count of values
    fbr_record_t msg;             // write_pipe / put_queue(single value)
    const fbr_record_t* list;     // put_queue(list): device-mapped pinned array
    unsigned long long idle_ns;   // watchdog
    ProcCtrl* ctrl;
};

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// returns 0 ok, 1 killed, 2 watchdog
__device__ int dev_recv(const ProcArgs& a, Lane* l, fbr_record_t* r) {
    const unsigned long long t0 = gtime();
    for (;;) {
        const unsigned long long t = l->tail;
        const unsigned long long h = *(volatile unsigned long long*)&l->head;
        if (h != t) {
            __threadfence_system();
            const volatile uint4* src = (const volatile uint4*)&l->rec[t & (kLaneCapacity - 1)];
            uint4* dst = (uint4*)r;
            for (int i = 0; i < 4; ++i) { uint4 v; v.x = src[i].x; v.y = src[i].y; v.z = src[i].z; v.w = src[i].w; dst[i] = v; }
            __threadfence_system();
            l->tail = t + 1;
            return 0;
        }
        if (a.ctrl->kill) return 1;
        if (gtime() - t0 > a.idle_ns) return 2;
        __nanosleep(2000);
    }
}
__device__ int dev_send(const ProcArgs& a, Lane* l, const fbr_record_t* r) {
    const unsigned long long t0 = gtime();
    for (;;) {
        const unsigned long long h = l->head;
        const unsigned long long t = *(volatile unsigned long long*)&l->tail;
        if (h - t < kLaneCapacity) {
            volatile uint4* dst = (volatile uint4*)&l->rec[h & (kLaneCapacity - 1)];
            const uint4* src = (const uint4*)r;
            for (int i = 0; i < 4; ++i) { dst[i].x = src[i].x; dst[i].y = src[i].y; dst[i].z = src[i].z; dst[i].w = src[i].w; }
            __threadfence_system();
            l->head = h + 1;
            return 0;
        }
        if (a.ctrl->kill) return 1;
        if (gtime() - t0 > a.idle_ns) return 2;
        __nanosleep(2000);
    }
}
__device__ bool is_str(const fbr_record_t& r, const char* s, uint32_t n) {
    if (r.tag != FBR_REC_STR || r.len != n) return false;
    for (uint32_t i = 0; i < n; ++i)
        if (r.payload[i] != (uint8_t)s[i]) return false;
    return true;
}
__device__ fbr_record_t make_int(long long v) {
    fbr_record_t r;
    memset(&r, 0, sizeof r);
    r.tag = FBR_REC_INT;
    r.len = 8;
    memcpy(r.payload, &v, 8);
    return r;
}
__device__ fbr_record_t make_bytes(const char* s, uint32_t n) {
    fbr_record_t r;
    memset(&r, 0, sizeof r);
    r.tag = FBR_REC_BYTES;
    r.len = n;
    for (uint32_t i = 0; i < n; ++i) r.payload[i] = (uint8_t)s[i];
    return r;
}

// One resident warp per device process; lane 0 runs the target function.
__global__ void __launch_bounds__(32) device_process_kernel(const ProcArgs a) {
    if (threadIdx.x != 0) return;
    a.ctrl->state = 1;
    __threadfence_system();
    int rc = 0;
    unsigned long long handled = 0;
    fbr_record_t rec;
    switch (a.kind) {
    case FBR_PROC_QUEUE_WORKER:      // tests/test_queue.py:44-50  worker(q_in, q_out, ident)
        for (;;) {
            rc = dev_recv(a, a.in, &rec);
            if (rc) break;
            if (is_str(rec, "quit", 4)) break;
            const fbr_record_t id = make_int(a.ident);
            rc = dev_send(a, a.out, &id);
            if (rc) break;
            ++handled;
        }
        break;
    case FBR_PROC_PUT_QUEUE:         // tests/test_queue.py:23-33  put_queue(q, data)
        if (a.list == nullptr) {
            rc = dev_send(a, a.out, &a.msg);
            handled = 1;
        } else {
            for (long long i = 0; i < a.ident && !rc; ++i) {
                rec = a.list[i];
                rc = dev_send(a, a.out, &rec);
                ++handled;
            }
        }
        break;
    case FBR_PROC_GET_QUEUE:         // tests/test_queue.py:36-42  get_queue(q_in, q_out, n)
        for (long long i = 0; i < a.ident; ++i) {
            rc = dev_recv(a, a.in, &rec);
            if (rc) break;
            rc = dev_send(a, a.out, &rec);
            if (rc) break;
            ++handled;
        }
        break;
    case FBR_PROC_WRITE_PIPE:        // tests/test_queue.py:19-20  write_pipe(pipe, msg)
        rc = dev_send(a, a.out, &a.msg);
        handled = 1;
        break;
    case FBR_PROC_PIPE_WORKER: {     // tests/test_queue.py:53-57  pipe_worker(conn)
        rc = dev_recv(a, a.in, &rec);
        if (!rc) {
            const fbr_record_t ack = make_bytes("ack", 3);
            rc = dev_send(a, a.out, &ack);
            handled = 1;
        }
        break;
    }
    default:
        rc = 4;
    }
    a.ctrl->handled = handled;
    a.ctrl->exitcode = rc == 0 ? 0 : (rc == 1 ? -15 : (rc == 2 ? 3 : 4));
    __threadfence_system();
    a.ctrl->state = 2;
    __threadfence_system();
}

struct fbr_process {
    int device = 0;
    cudaStream_t stream = nullptr;
    ProcCtrl* ctrl = nullptr;          // pinned mapped
    fbr_record_t* list = nullptr;      // pinned mapped copy of put_queue's list
    cudaEvent_t done = nullptr;
    bool started = false;
};

// ------------------------------------------------------------------------------------------------
// extern "C"
// ------------------------------------------------------------------------------------------------
extern "C" int fbr_internal_preload(int device);   // engine.cu

extern "C" {

int fbr_queue_create(fbr_queue_t** out) {
    if (!out) return qfail(FBR_EINVAL, "NULL argument");
    fbr_queue* q = new fbr_queue();
    {
        std::lock_guard<std::mutex> g(hub().mu);
        hub().queues.push_back(q);
    }
    hub().ensure_started();
    *out = q;
    return FBR_OK;
}

int fbr_queue_open_writer(fbr_queue_t* q, fbr_lane_t** lane) {
    if (!q || !lane) return qfail(FBR_EINVAL, "NULL argument");
    return lane_create(q, true, lane);
}

int fbr_queue_open_reader(fbr_queue_t* q, fbr_lane_t** lane) {
    if (!q || !lane) return qfail(FBR_EINVAL, "NULL argument");
    return lane_create(q, false, lane);
}

int fbr_lane_send(fbr_lane_t* l, const fbr_record_t* rec, int timeout_ms) {
    if (!l || !rec || !l->is_writer) return qfail(FBR_EINVAL, "not a writer lane");
    std::lock_guard<std::mutex> g(l->mu);
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(timeout_ms);
    while (!lane_push_host(l->ring, rec)) {
        if (timeout_ms >= 0 && std::chrono::steady_clock::now() >= deadline) return qfail(FBR_ETIMEOUT, "queue full");
        std::this_thread::sleep_for(std::chrono::microseconds(20));
    }
    return FBR_OK;
}

int fbr_lane_recv(fbr_lane_t* l, fbr_record_t* rec, int timeout_ms) {
    if (!l || !rec || l->is_writer) return qfail(FBR_EINVAL, "not a reader lane");
    std::lock_guard<std::mutex> g(l->mu);
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(timeout_ms);
    int spins = 0;
    while (!lane_pop_host(l->ring, rec)) {
        if (timeout_ms >= 0 && std::chrono::steady_clock::now() >= deadline) return qfail(FBR_ETIMEOUT, "queue empty");
        if (++spins > 200) std::this_thread::sleep_for(std::chrono::microseconds(20));
    }
    return FBR_OK;
}

int fbr_lane_poll(fbr_lane_t* l, int* ready) {
    if (!l || !ready) return qfail(FBR_EINVAL, "NULL argument");
    *ready = l->ring->tail != __atomic_load_n(&l->ring->head, __ATOMIC_ACQUIRE);
    return FBR_OK;
}

/* SimpleQueue.put / get on the host's own lazily opened endpoints (LazyZConnection,
 * fiber/queues.py:190-249: the reader connects on first use). */
static std::mutex g_host_lane_mu;   // serialises the lazy creation of a queue's own host endpoints

int fbr_queue_put(fbr_queue_t* q, const fbr_record_t* rec, int timeout_ms) {
    if (!q || !rec) return qfail(FBR_EINVAL, "NULL argument");
    {
        std::lock_guard<std::mutex> g(g_host_lane_mu);
        if (!q->host_writer) {
            fbr_lane* l = nullptr;
            int rc = lane_create(q, true, &l);
            if (rc) return rc;
            q->host_writer = l;
        }
    }
    return fbr_lane_send(q->host_writer, rec, timeout_ms);
}

int fbr_queue_get(fbr_queue_t* q, fbr_record_t* rec, int timeout_ms) {
    if (!q || !rec) return qfail(FBR_EINVAL, "NULL argument");
    {
        std::lock_guard<std::mutex> g(g_host_lane_mu);
        if (!q->host_reader) {
            fbr_lane* l = nullptr;
            int rc = lane_create(q, false, &l);
            if (rc) return rc;
            q->host_reader = l;
        }
    }
    return fbr_lane_recv(q->host_reader, rec, timeout_ms);
}

int fbr_queue_stats(fbr_queue_t* q, uint64_t* forwarded, uint32_t* n_writers, uint32_t* n_readers) {
    if (!q) return qfail(FBR_EINVAL, "NULL argument");
    std::lock_guard<std::mutex> g(q->mu);
    if (forwarded) *forwarded = q->forwarded;
    if (n_writers) *n_writers = (uint32_t)q->writers.size();
    if (n_readers) *n_readers = (uint32_t)q->readers.size();
    return FBR_OK;
}

int fbr_queue_destroy(fbr_queue_t* q) {
    if (!q) return FBR_OK;
    {
        std::lock_guard<std::mutex> g(hub().mu);
        auto& v = hub().queues;
        for (size_t i = 0; i < v.size(); ++i)
            if (v[i] == q) { v.erase(v.begin() + i); break; }
    }
    std::lock_guard<std::mutex> g(q->mu);   // the forwarder is not inside this queue any more
    q->closed = true;
    // lanes stay allocated (a device process may still hold the mapped pointers); they are small
    // and are reclaimed at process exit.
    return FBR_OK;
}

/* ---- device processes ------------------------------------------------------------------------- */
int fbr_process_start(int device_id, int kind, fbr_lane_t* in, fbr_lane_t* out, int64_t ident,
                      const fbr_record_t* msg, const fbr_record_t* list, uint32_t list_len, int idle_timeout_ms,
                      fbr_process_t** proc) {
    if (!proc) return qfail(FBR_EINVAL, "NULL argument");
    if (in && in->is_writer) return qfail(FBR_EINVAL, "`in` must be a reader lane");
    if (out && !out->is_writer) return qfail(FBR_EINVAL, "`out` must be a writer lane");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return qfail(FBR_ENODEV, "no CUDA device visible; device processes have no CPU fallback");
    }
    if (device_id < 0 || device_id >= ndev) return qfail(FBR_EINVAL, "device id %d out of range", device_id);
    std::unique_ptr<fbr_process> p(new fbr_process());
    p->device = device_id;
    QCK(cudaSetDevice(device_id));
    // load every kernel of the library now: a lazy module load later would synchronise with the
    // resident kernel started below (which may be waiting for this very host thread)
    fbr_internal_preload(device_id);
    {
        cudaFuncAttributes at;
        cudaFuncGetAttributes(&at, (const void*)device_process_kernel);
    }
    QCK(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    QCK(cudaHostAlloc((void**)&p->ctrl, sizeof(ProcCtrl), cudaHostAllocPortable | cudaHostAllocMapped));
    memset((void*)p->ctrl, 0, sizeof(ProcCtrl));
    ProcArgs a;
    memset(&a, 0, sizeof a);
    a.kind = kind;
    a.in = in ? in->ring : nullptr;
    a.out = out ? out->ring : nullptr;
    a.ident = ident;
    if (msg) a.msg = *msg;
    if (list && list_len) {
        QCK(cudaHostAlloc((void**)&p->list, sizeof(fbr_record_t) * list_len, cudaHostAllocPortable | cudaHostAllocMapped));
        memcpy(p->list, list, sizeof(fbr_record_t) * list_len);
        a.list = p->list;
        a.ident = list_len;
    }
    a.idle_ns = (unsigned long long)(idle_timeout_ms > 0 ? idle_timeout_ms : 30000) * 1000000ull;
    a.ctrl = p->ctrl;
    QCK(cudaEventCreateWithFlags(&p->done, cudaEventDisableTiming));
    device_process_kernel<<<1, 32, 0, p->stream>>>(a);
    QCK(cudaGetLastError());
    QCK(cudaEventRecord(p->done, p->stream));
    p->started = true;
    *proc = p.release();
    return FBR_OK;
}

int fbr_process_poll(fbr_process_t* p, int* alive, int* exitcode) {
    if (!p) return qfail(FBR_EINVAL, "NULL argument");
    const int st = p->ctrl->state;
    bool finished = st == 2;
    if (finished) {
        cudaSetDevice(p->device);
        finished = cudaEventQuery(p->done) == cudaSuccess;
        cudaGetLastError();
    }
    if (alive) *alive = finished ? 0 : 1;
    if (exitcode) *exitcode = finished ? p->ctrl->exitcode : 0;
    return FBR_OK;
}

int fbr_process_join(fbr_process_t* p, int timeout_ms) {
    if (!p) return qfail(FBR_EINVAL, "NULL argument");
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(timeout_ms);
    for (;;) {
        int alive = 1;
        fbr_process_poll(p, &alive, nullptr);
        if (!alive) return FBR_OK;
        if (timeout_ms >= 0 && std::chrono::steady_clock::now() >= deadline) return qfail(FBR_ETIMEOUT, "process still running");
        std::this_thread::sleep_for(std::chrono::microseconds(50));
    }
}

int fbr_process_terminate(fbr_process_t* p) {
    if (!p) return qfail(FBR_EINVAL, "NULL argument");
    p->ctrl->kill = 1;
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
    return FBR_OK;
}

int fbr_process_handled(fbr_process_t* p, uint64_t* handled) {
    if (!p || !handled) return qfail(FBR_EINVAL, "NULL argument");
    *handled = p->ctrl->handled;
    return FBR_OK;
}

int fbr_process_destroy(fbr_process_t* p) {
    if (!p) return FBR_OK;
    fbr_process_terminate(p);
    fbr_process_join(p, 5000);
    cudaSetDevice(p->device);
    cudaStreamSynchronize(p->stream);
    cudaStreamDestroy(p->stream);
    cudaEventDestroy(p->done);
    // ctrl/list are tiny pinned blocks; freeing pinned memory can synchronise with other resident
    // device processes, so they are left to process exit.
    delete p;
    return FBR_OK;
}

}  // extern "C"
