// bagua_b200 multi-stream TCP transport for NCCL's external-network plugin interface.
//
// Role of the reference's Bagua-Net (rust/bagua-net/src/implement/nthread_per_socket_backend.rs, cc/v4/nccl_net_v4.cc):
// inter-node NCCL traffic over plain Ethernet is spread over several TCP streams so that one flow's congestion window
// does not cap the whole transfer.  Inside one NVSwitch domain this code is never reached (the peer-memory kernels and
// NCCL's P2P/NVLS transports carry everything); it exists for multi-node jobs on TCP fabrics.
//
// Design (not the reference's):
//   * one control connection + S data connections per (sender, receiver) pair; every connection has a dedicated
//     worker thread that sleeps on a condition variable (no busy polling next to the training process);
//   * messages <= kInlineBytes travel on the control connection right behind their 8-byte length header (one syscall,
//     no fan-out latency); larger ones are cut into at most S chunks of >= min_chunk bytes by plan_chunks(), which
//     both sides evaluate identically, starting at a rotating stream cursor so consecutive messages load all streams;
//   * requests live in a fixed per-connection pool; completion is one atomic counter per request.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace bagua_net {

constexpr size_t kInlineBytes = 16 * 1024;
constexpr int kMaxStreams = 16;
constexpr int kMaxRequests = 64;
constexpr uint32_t kMagic = 0xBA60A200u;

struct Config {
    int nstreams = 4;                 // BAGUA_NET_NSTREAMS
    size_t min_chunk = 1u << 20;      // BAGUA_NET_MIN_CHUNKSIZE
    int sock_buf = 8 << 20;           // BAGUA_NET_SOCKBUF (SO_SNDBUF / SO_RCVBUF)
    static Config from_env();
};

struct Chunk {
    size_t offset;
    size_t bytes;
    int stream;
};
// Split a message of `size` bytes into n = clamp(size / min_chunk, 1, nstreams) nearly equal chunks (16-byte aligned cuts),
// assigned to streams (cursor + i) % nstreams.  Deterministic: sender and receiver derive the same plan from `size`.
std::vector<Chunk> plan_chunks(size_t size, int nstreams, size_t min_chunk, unsigned cursor);

struct NetDevice {
    std::string name;        // interface name
    std::string pci_path;    // /sys/class/net/<if>/device resolved, or empty
    int speed_mbps = 10000;  // /sys/class/net/<if>/speed
    unsigned char addr[128]; // sockaddr_storage image
    unsigned addr_len = 0;
};
// Interfaces usable for traffic.  Honours NCCL_SOCKET_IFNAME ("eth,ib" prefixes, "=eth0" exact, "^docker,lo" exclusion);
// without it loopback/docker are skipped unless nothing else exists.
std::vector<NetDevice> discover_devices();

struct Stats {
    std::atomic<uint64_t> isend_count{0}, irecv_count{0}, bytes_sent{0}, bytes_received{0};
    std::atomic<uint64_t> isend_ns{0}, irecv_ns{0};  // post→completion latency sums
    std::atomic<uint64_t> chunks{0}, inline_msgs{0}, errors{0};
    std::string json() const;
    std::string prometheus(int rank) const;
};
Stats& stats();
// Starts (once) a thread pushing stats().prometheus() to BAGUA_NET_PROMETHEUS_ADDRESS ("host:port") every 5 s.
void start_metrics_push_if_configured();
// Span tracing (the reference exports isend/irecv spans to Jaeger for ranks < 8, nthread_per_socket_backend.rs:113-137,226).
// BAGUA_NET_TRACE_FILE=<prefix> writes one Chrome/Perfetto trace-event file per rank (<prefix>.<rank>.json, "ph":"X" spans
// named isend/irecv with byte counts); BAGUA_NET_JAEGER_ADDRESS alone selects the same exporter with the file under
// $TMPDIR (no thrift/UDP Jaeger client is built in).  Off by default: trace_span() is one branch.
void trace_span(bool is_send, uint64_t start_ns, uint64_t end_ns, size_t bytes, bool failed);
void trace_flush();

struct Request {
    std::atomic<int> pending{0};   // outstanding pieces (control header + chunks)
    std::atomic<int> error{0};
    std::atomic<bool> in_use{false};
    size_t size = 0;               // bytes actually transferred (receiver learns it from the header)
    char* data = nullptr;
    size_t capacity = 0;
    bool is_send = false;
    uint64_t t_post_ns = 0;
};

// A blocking FIFO of work items served by one thread that owns one socket.
class Lane {
public:
    struct Item {
        Request* req;
        char* ptr;
        size_t bytes;
        bool header;  // control lane only: 8-byte length header precedes (send) / is read before (recv) the payload
    };
    Lane() = default;
    ~Lane();
    void start(int fd, bool sender, class Connection* owner, bool control);
    void push(const Item& it);
    void stop();
    int fd() const { return fd_; }

private:
    void run_sender();
    void run_receiver();
    bool next(Item& out);
    int fd_ = -1;
    bool control_ = false;
    class Connection* owner_ = nullptr;
    std::thread th_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<Item> q_;
    bool stopping_ = false;
};

class Connection {
public:
    Connection(bool sender, const Config& cfg) : sender_(sender), cfg_(cfg) {}
    ~Connection();
    // takes ownership of the descriptors; data_fds.size() == cfg.nstreams
    void adopt(int ctrl_fd, const std::vector<int>& data_fds);
    Request* post_send(void* data, size_t size);   // nullptr = no free request slot, try again
    Request* post_recv(void* data, size_t capacity);
    // receiver control lane: a header announced `size` bytes for `req`; fan the payload out over the data lanes
    void fan_out_recv(Request* req, size_t size);
    bool is_sender() const { return sender_; }
    const Config& config() const { return cfg_; }
    void fail(Request* req);
    static void complete_piece(Request* req);

private:
    Request* grab();
    bool sender_;
    Config cfg_;
    Lane ctrl_;
    Lane data_[kMaxStreams];
    int n_data_ = 0;
    unsigned cursor_ = 0;
    std::mutex mu_;
    Request pool_[kMaxRequests];
};

// What listen() publishes to the connecting side through NCCL's handle exchange.
struct Handle {
    uint32_t magic;
    uint32_t nstreams;
    uint64_t listen_id;
    unsigned addr_len;
    unsigned char addr[64];
};
static_assert(sizeof(Handle) <= 128, "NCCL_NET_HANDLE_MAXSIZE");

class Listener {
public:
    ~Listener();
    static Listener* create(const NetDevice& dev, const Config& cfg, Handle* out);
    // Non-blocking: collects incoming sockets; returns a receiver Connection once one peer's control + data streams
    // have all arrived, else nullptr (NCCL calls accept() again).
    Connection* try_accept();

private:
    struct Pending {
        uint64_t conn_id;
        int ctrl = -1;
        std::vector<int> data;
        int have = 0;
    };
    int fd_ = -1;
    Config cfg_;
    std::vector<Pending> pending_;
};

// Blocking connect of 1 + nstreams sockets to the listener described by `h` (a TCP connect completes against the
// listen backlog, so this cannot deadlock with the peer's own connect()).
Connection* connect_to(const Handle& h, const Config& cfg);

uint64_t now_ns();

}  // namespace bagua_net
