// NCCL external-network plugin ("libnccl-net-bagua.so") on top of the multi-stream TCP transport in net_engine.*.
//
// Exports the v6 table (`ncclNetPlugin_v6`): NCCL 2.13 … 2.28 all accept it (2.28 probes v11 → v6 in turn), and v6 is
// the last revision whose calls carry only host-side objects, which is all a TCP transport needs.  The reference ships
// v3 + v4 tables (rust/bagua-net/cc/v3, cc/v4) for NCCL 2.6 … 2.10; those ABIs are no longer loaded by current NCCL.
//
// Activate with  NCCL_NET_PLUGIN=bagua  and the directory of this library on LD_LIBRARY_PATH
// (`bagua_b200.net.enable()` / the launchers' --enable_bagua_net do both).
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#include "net_engine.h"

#define BAGUA_EXPORT __attribute__((visibility("default")))

extern "C" {

// --- the slice of nccl.h / nccl_net.h this file needs (declarations only; layout fixed by NCCL's plugin ABI) ---------
typedef enum {
    ncclSuccess = 0,
    ncclUnhandledCudaError = 1,
    ncclSystemError = 2,
    ncclInternalError = 3,
    ncclInvalidArgument = 4,
    ncclInvalidUsage = 5,
    ncclRemoteError = 6,
    ncclInProgress = 7
} ncclResult_t;

#define NCCL_NET_HANDLE_MAXSIZE 128
#define NCCL_PTR_HOST 0x1
#define NCCL_PTR_CUDA 0x2
#define NCCL_PTR_DMABUF 0x4

typedef enum { NCCL_LOG_NONE = 0, NCCL_LOG_VERSION = 1, NCCL_LOG_WARN = 2, NCCL_LOG_INFO = 3, NCCL_LOG_ABORT = 4, NCCL_LOG_TRACE = 5 } ncclDebugLogLevel;
typedef enum { NCCL_INIT = 1, NCCL_COLL = 2, NCCL_P2P = 4, NCCL_SHM = 8, NCCL_NET = 16 } ncclDebugLogSubSys;
typedef void (*ncclDebugLogger_t)(ncclDebugLogLevel level, unsigned long flags, const char* file, int line, const char* fmt, ...);

typedef struct {
    char* name;      // used mostly for logging
    char* pciPath;   // path to the PCI device in /sys
    uint64_t guid;   // unique identifier of the NIC chip
    int ptrSupport;  // NCCL_PTR_HOST [| NCCL_PTR_CUDA | NCCL_PTR_DMABUF]
    int speed;       // port speed in Mbps
    int port;        // port number
    float latency;   // network latency in microseconds
    int maxComms;    // maximum number of comms we can create
    int maxRecvs;    // maximum number of grouped receives
} ncclNetProperties_v6_t;

typedef struct {
    const char* name;
    ncclResult_t (*init)(ncclDebugLogger_t logFunction);
    ncclResult_t (*devices)(int* ndev);
    ncclResult_t (*getProperties)(int dev, ncclNetProperties_v6_t* props);
    ncclResult_t (*listen)(int dev, void* handle, void** listenComm);
    ncclResult_t (*connect)(int dev, void* handle, void** sendComm);
    ncclResult_t (*accept)(void* listenComm, void** recvComm);
    ncclResult_t (*regMr)(void* comm, void* data, int size, int type, void** mhandle);
    ncclResult_t (*regMrDmaBuf)(void* comm, void* data, size_t size, int type, uint64_t offset, int fd, void** mhandle);
    ncclResult_t (*deregMr)(void* comm, void* mhandle);
    ncclResult_t (*isend)(void* sendComm, void* data, int size, int tag, void* mhandle, void** request);
    ncclResult_t (*irecv)(void* recvComm, int n, void** data, int* sizes, int* tags, void** mhandles, void** request);
    ncclResult_t (*iflush)(void* recvComm, int n, void** data, int* sizes, void** mhandles, void** request);
    ncclResult_t (*test)(void* request, int* done, int* sizes);
    ncclResult_t (*closeSend)(void* sendComm);
    ncclResult_t (*closeRecv)(void* recvComm);
    ncclResult_t (*closeListen)(void* listenComm);
} ncclNet_v6_t;

}  // extern "C"

namespace {
using namespace bagua_net;

ncclDebugLogger_t g_log = nullptr;
Config g_cfg;
std::vector<NetDevice> g_devs;

#define NET_INFO(...)                                                              \
    do {                                                                           \
        if (g_log) g_log(NCCL_LOG_INFO, NCCL_INIT | NCCL_NET, __FILE__, __LINE__, __VA_ARGS__); \
    } while (0)
#define NET_WARN(...)                                                   \
    do {                                                                \
        if (g_log) g_log(NCCL_LOG_WARN, NCCL_NET, __FILE__, __LINE__, __VA_ARGS__); \
    } while (0)

ncclResult_t net_init(ncclDebugLogger_t logger) {
    g_log = logger;
    g_cfg = Config::from_env();
    g_devs = discover_devices();
    start_metrics_push_if_configured();
    NET_INFO("BaguaNet-B200: %d interface(s), %d data streams per connection, min chunk %zu bytes", static_cast<int>(g_devs.size()),
             g_cfg.nstreams, g_cfg.min_chunk);
    for (auto& d : g_devs) NET_INFO("BaguaNet-B200: using %s (%d Mb/s)", d.name.c_str(), d.speed_mbps);
    return g_devs.empty() ? ncclSystemError : ncclSuccess;
}

ncclResult_t net_devices(int* ndev) {
    *ndev = static_cast<int>(g_devs.size());
    return ncclSuccess;
}

ncclResult_t net_get_properties(int dev, ncclNetProperties_v6_t* props) {
    if (dev < 0 || dev >= static_cast<int>(g_devs.size())) return ncclInvalidArgument;
    NetDevice& d = g_devs[dev];
    props->name = const_cast<char*>(d.name.c_str());
    props->pciPath = d.pci_path.empty() ? nullptr : const_cast<char*>(d.pci_path.c_str());
    props->guid = static_cast<uint64_t>(dev);
    props->ptrSupport = NCCL_PTR_HOST;  // NCCL stages GPU buffers through pinned host memory for us
    props->speed = d.speed_mbps;
    props->port = 0;
    props->latency = 0.f;
    props->maxComms = 65536;
    props->maxRecvs = 1;
    return ncclSuccess;
}

ncclResult_t net_listen(int dev, void* handle, void** listen_comm) {
    if (dev < 0 || dev >= static_cast<int>(g_devs.size())) return ncclInvalidArgument;
    static_assert(sizeof(Handle) <= NCCL_NET_HANDLE_MAXSIZE, "handle too large");
    Listener* l = Listener::create(g_devs[dev], g_cfg, static_cast<Handle*>(handle));
    if (!l) {
        NET_WARN("BaguaNet-B200: listen on %s failed", g_devs[dev].name.c_str());
        return ncclSystemError;
    }
    *listen_comm = l;
    return ncclSuccess;
}

ncclResult_t net_connect(int /*dev*/, void* handle, void** send_comm) {
    Connection* c = connect_to(*static_cast<Handle*>(handle), g_cfg);
    if (!c) {
        NET_WARN("BaguaNet-B200: connect failed");
        return ncclSystemError;
    }
    *send_comm = c;
    return ncclSuccess;
}

ncclResult_t net_accept(void* listen_comm, void** recv_comm) {
    *recv_comm = static_cast<Listener*>(listen_comm)->try_accept();  // nullptr → NCCL polls again
    return ncclSuccess;
}

ncclResult_t net_reg_mr(void* /*comm*/, void* /*data*/, int /*size*/, int type, void** mhandle) {
    *mhandle = nullptr;
    return type == NCCL_PTR_HOST ? ncclSuccess : ncclInternalError;
}
ncclResult_t net_reg_mr_dmabuf(void*, void*, size_t, int, uint64_t, int, void**) { return ncclInternalError; }
ncclResult_t net_dereg_mr(void*, void*) { return ncclSuccess; }

ncclResult_t net_isend(void* send_comm, void* data, int size, int /*tag*/, void* /*mhandle*/, void** request) {
    *request = static_cast<Connection*>(send_comm)->post_send(data, static_cast<size_t>(size));
    if (*request) stats().isend_count.fetch_add(1, std::memory_order_relaxed);
    return ncclSuccess;
}

ncclResult_t net_irecv(void* recv_comm, int n, void** data, int* sizes, int* /*tags*/, void** /*mhandles*/, void** request) {
    if (n != 1) return ncclInternalError;  // maxRecvs == 1
    *request = static_cast<Connection*>(recv_comm)->post_recv(data[0], static_cast<size_t>(sizes[0]));
    if (*request) stats().irecv_count.fetch_add(1, std::memory_order_relaxed);
    return ncclSuccess;
}

ncclResult_t net_iflush(void*, int, void**, int*, void**, void** request) {
    *request = nullptr;  // host memory only: nothing to flush
    return ncclSuccess;
}

ncclResult_t net_test(void* request, int* done, int* sizes) {
    auto* r = static_cast<Request*>(request);
    *done = 0;
    if (r->pending.load(std::memory_order_acquire) > 0) return ncclSuccess;
    const bool failed = r->error.load(std::memory_order_relaxed) != 0;
    const uint64_t t_done = now_ns();
    const uint64_t dt = t_done - r->t_post_ns;
    trace_span(r->is_send, r->t_post_ns, t_done, r->size, failed);
    if (r->is_send) {
        stats().bytes_sent.fetch_add(r->size, std::memory_order_relaxed);
        stats().isend_ns.fetch_add(dt, std::memory_order_relaxed);
    } else {
        stats().bytes_received.fetch_add(r->size, std::memory_order_relaxed);
        stats().irecv_ns.fetch_add(dt, std::memory_order_relaxed);
    }
    if (sizes) *sizes = static_cast<int>(r->size);
    *done = 1;
    r->in_use.store(false, std::memory_order_release);
    return failed ? ncclSystemError : ncclSuccess;
}

ncclResult_t net_close_conn(void* comm) {
    delete static_cast<Connection*>(comm);
    return ncclSuccess;
}
ncclResult_t net_close_listen(void* comm) {
    delete static_cast<Listener*>(comm);
    return ncclSuccess;
}
}  // namespace

extern "C" {

BAGUA_EXPORT ncclNet_v6_t ncclNetPlugin_v6 = {
    "BaguaNet-B200", net_init,  net_devices, net_get_properties, net_listen,     net_connect,    net_accept,      net_reg_mr,
    net_reg_mr_dmabuf, net_dereg_mr, net_isend,   net_irecv,          net_iflush,     net_test,       net_close_conn,  net_close_conn,
    net_close_listen,
};

// ---- helpers for tests / tooling (not part of NCCL's ABI) ----------------------------------------------------------
// Writes up to `cap` (offset, bytes, stream) triples; returns the number of chunks of the plan.
BAGUA_EXPORT int bagua_net_plan_chunks(uint64_t size, int nstreams, uint64_t min_chunk, unsigned cursor, uint64_t* out, int cap) {
    auto plan = bagua_net::plan_chunks(size, nstreams, min_chunk, cursor);
    int n = 0;
    for (auto& c : plan) {
        if (n < cap) {
            out[3 * n] = c.offset;
            out[3 * n + 1] = c.bytes;
            out[3 * n + 2] = static_cast<uint64_t>(c.stream);
        }
        ++n;
    }
    return n;
}

BAGUA_EXPORT int bagua_net_stats_json(char* buf, int cap) {
    std::string s = bagua_net::stats().json();
    if (cap > 0) {
        strncpy(buf, s.c_str(), static_cast<size_t>(cap) - 1);
        buf[cap - 1] = 0;
    }
    return static_cast<int>(s.size());
}

BAGUA_EXPORT void bagua_net_trace_flush() { bagua_net::trace_flush(); }

BAGUA_EXPORT int bagua_net_device_count() { return static_cast<int>(bagua_net::discover_devices().size()); }
}
