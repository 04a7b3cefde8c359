// Multi-stream TCP transport — see net_engine.h for the design.
#include "net_engine.h"

#include <strings.h>

#include <arpa/inet.h>
#include <cerrno>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <ifaddrs.h>
#include <net/if.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <time.h>
#include <unistd.h>

namespace bagua_net {

uint64_t now_ns() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return static_cast<uint64_t>(ts.tv_sec) * 1000000000ull + ts.tv_nsec;
}

// ---- configuration ---------------------------------------------------------------------------------------
static long env_long(const char* name, long dflt) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    char* end = nullptr;
    long x = strtol(v, &end, 10);
    return (end && end != v) ? x : dflt;
}

Config Config::from_env() {
    Config c;
    long s = env_long("BAGUA_NET_NSTREAMS", c.nstreams);
    c.nstreams = static_cast<int>(s < 1 ? 1 : (s > kMaxStreams ? kMaxStreams : s));
    long m = env_long("BAGUA_NET_MIN_CHUNKSIZE", static_cast<long>(c.min_chunk));
    c.min_chunk = static_cast<size_t>(m < 4096 ? 4096 : m);
    c.sock_buf = static_cast<int>(env_long("BAGUA_NET_SOCKBUF", c.sock_buf));
    // The reference selects between two transports with BAGUA_NET_IMPLEMENT=BASIC|TOKIO (rust/bagua-net/src/lib.rs:19-33); this
    // plugin has one (thread per socket, condition-variable driven), so both names are accepted and anything else is reported.
    const char* impl = getenv("BAGUA_NET_IMPLEMENT");
    if (impl && *impl && strcasecmp(impl, "BASIC") != 0 && strcasecmp(impl, "TOKIO") != 0) {
        static std::once_flag warned;
        std::call_once(warned, [impl] { fprintf(stderr, "bagua-net: unknown BAGUA_NET_IMPLEMENT=%s (BASIC and TOKIO select the same transport here)\n", impl); });
    }
    return c;
}

std::vector<Chunk> plan_chunks(size_t size, int nstreams, size_t min_chunk, unsigned cursor) {
    std::vector<Chunk> out;
    if (size == 0) return out;
    size_t n = size / min_chunk;
    if (n < 1) n = 1;
    if (n > static_cast<size_t>(nstreams)) n = nstreams;
    // equal shares rounded up to 16 bytes; the last chunk takes the remainder
    size_t share = ((size + n - 1) / n + 15) & ~static_cast<size_t>(15);
    size_t off = 0;
    for (size_t i = 0; i < n && off < size; ++i) {
        size_t len = (i + 1 == n || off + share > size) ? size - off : share;
        out.push_back(Chunk{off, len, static_cast<int>((cursor + i) % nstreams)});
        off += len;
    }
    return out;
}

// ---- statistics ---------------------------------------------------------------------------------------------
Stats& stats() {
    static Stats s;
    return s;
}

// (no iostreams in this library: it is dlopen()ed into processes that may carry a different libstdc++)
std::string Stats::json() const {
    char buf[512];
    snprintf(buf, sizeof(buf),
             "{\"isend_count\":%llu,\"irecv_count\":%llu,\"bytes_sent\":%llu,\"bytes_received\":%llu,\"isend_ns\":%llu,\"irecv_ns\":%llu,"
             "\"chunks\":%llu,\"inline_msgs\":%llu,\"errors\":%llu}",
             (unsigned long long)isend_count.load(), (unsigned long long)irecv_count.load(), (unsigned long long)bytes_sent.load(),
             (unsigned long long)bytes_received.load(), (unsigned long long)isend_ns.load(), (unsigned long long)irecv_ns.load(),
             (unsigned long long)chunks.load(), (unsigned long long)inline_msgs.load(), (unsigned long long)errors.load());
    return buf;
}

std::string Stats::prometheus(int rank) const {
    std::string o;
    auto line = [&](const char* name, const char* help, uint64_t v) {
        char buf[512];
        snprintf(buf, sizeof(buf), "# HELP %s %s\n# TYPE %s counter\n%s{rank=\"%d\"} %llu\n", name, help, name, name, rank, (unsigned long long)v);
        o += buf;
    };
    line("bagua_net_isend_total", "messages posted with isend", isend_count.load());
    line("bagua_net_irecv_total", "messages posted with irecv", irecv_count.load());
    line("bagua_net_sent_bytes_total", "payload bytes sent", bytes_sent.load());
    line("bagua_net_received_bytes_total", "payload bytes received", bytes_received.load());
    line("bagua_net_isend_nanoseconds_total", "sum of isend post-to-completion latencies", isend_ns.load());
    line("bagua_net_irecv_nanoseconds_total", "sum of irecv post-to-completion latencies", irecv_ns.load());
    line("bagua_net_errors_total", "failed requests", errors.load());
    return o;
}

// ---- socket helpers ------------------------------------------------------------------------------------------
static bool write_all(int fd, const char* p, size_t n) {
    while (n) {
        ssize_t w = ::send(fd, p, n > (1u << 30) ? (1u << 30) : n, MSG_NOSIGNAL);
        if (w < 0) {
            if (errno == EINTR) continue;
            return false;
        }
        p += w;
        n -= static_cast<size_t>(w);
    }
    return true;
}

static bool read_all(int fd, char* p, size_t n) {
    while (n) {
        ssize_t r = ::recv(fd, p, n > (1u << 30) ? (1u << 30) : n, 0);
        if (r < 0) {
            if (errno == EINTR) continue;
            return false;
        }
        if (r == 0) return false;  // peer closed
        p += r;
        n -= static_cast<size_t>(r);
    }
    return true;
}

static void tune_socket(int fd, const Config& cfg) {
    int one = 1;
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
    if (cfg.sock_buf > 0) {
        setsockopt(fd, SOL_SOCKET, SO_SNDBUF, &cfg.sock_buf, sizeof(int));
        setsockopt(fd, SOL_SOCKET, SO_RCVBUF, &cfg.sock_buf, sizeof(int));
    }
}

static void set_blocking(int fd, bool blocking) {
    int fl = fcntl(fd, F_GETFL, 0);
    if (fl < 0) return;
    fcntl(fd, F_SETFL, blocking ? (fl & ~O_NONBLOCK) : (fl | O_NONBLOCK));
}

struct Hello {
    uint32_t magic;
    uint32_t lane;  // 0 = control, 1 + i = data stream i
    uint64_t conn_id;
    uint32_t nstreams;
    uint32_t reserved;
};

// ---- device discovery ------------------------------------------------------------------------------------------
namespace {
struct IfFilter {
    bool exclude = false, exact = false;
    std::vector<std::string> names;
    bool active = false;
    bool match(const std::string& ifn) const {
        bool hit = false;
        for (auto& n : names) {
            if (exact ? ifn == n : ifn.compare(0, n.size(), n) == 0) {
                hit = true;
                break;
            }
        }
        return exclude ? !hit : hit;
    }
};

IfFilter parse_ifname_filter() {
    IfFilter f;
    const char* v = getenv("NCCL_SOCKET_IFNAME");
    if (!v || !*v) return f;
    std::string s(v);
    if (s[0] == '^') {
        f.exclude = true;
        s.erase(0, 1);
    } else if (s[0] == '=') {
        f.exact = true;
        s.erase(0, 1);
    }
    size_t pos = 0;
    while (pos <= s.size()) {
        size_t comma = s.find(',', pos);
        if (comma == std::string::npos) comma = s.size();
        if (comma > pos) f.names.push_back(s.substr(pos, comma - pos));
        pos = comma + 1;
    }
    f.active = !f.names.empty();
    return f;
}

// NCCL_SOCKET_FAMILY restricts the address family of the interfaces considered: NCCL's own spelling (AF_INET / AF_INET6), the
// reference's numeric sa_family value (rust/bagua-net/src/utils.rs:33-36,101: 2 / 10) or plain 4 / 6; anything else = no restriction.
int parse_socket_family() {
    const char* v = getenv("NCCL_SOCKET_FAMILY");
    if (!v || !*v) return -1;
    std::string s(v);
    if (s == "AF_INET" || s == "4" || s == std::to_string(AF_INET)) return AF_INET;
    if (s == "AF_INET6" || s == "6" || s == std::to_string(AF_INET6)) return AF_INET6;
    return -1;
}

int read_speed(const std::string& ifn) {
    FILE* fp = fopen(("/sys/class/net/" + ifn + "/speed").c_str(), "r");
    int sp = -1;
    if (fp) {
        if (fscanf(fp, "%d", &sp) != 1) sp = -1;
        fclose(fp);
    }
    return sp > 0 ? sp : 10000;
}

std::string read_pci_path(const std::string& ifn) {
    char buf[PATH_MAX];
    std::string link = "/sys/class/net/" + ifn + "/device";
    return realpath(link.c_str(), buf) ? std::string(buf) : std::string();
}
}  // namespace

std::vector<NetDevice> discover_devices() {
    std::vector<NetDevice> devs, fallback;
    IfFilter filt = parse_ifname_filter();
    const int want_family = parse_socket_family();
    ifaddrs* list = nullptr;
    if (getifaddrs(&list) != 0) return devs;
    for (ifaddrs* it = list; it; it = it->ifa_next) {
        if (!it->ifa_addr || !(it->ifa_flags & IFF_UP) || !(it->ifa_flags & IFF_RUNNING)) continue;
        const int fam = it->ifa_addr->sa_family;
        if (fam != AF_INET && fam != AF_INET6) continue;
        if (want_family != -1 && fam != want_family) continue;
        if (fam == AF_INET6) {
            auto* a6 = reinterpret_cast<sockaddr_in6*>(it->ifa_addr);
            if (IN6_IS_ADDR_LINKLOCAL(&a6->sin6_addr)) continue;
        }
        std::string ifn(it->ifa_name);
        bool dup = false;
        for (auto& d : devs) dup |= (d.name == ifn);
        for (auto& d : fallback) dup |= (d.name == ifn);
        if (dup) continue;  // first address of an interface wins (IPv4 is listed first by the kernel)
        NetDevice d;
        d.name = ifn;
        d.addr_len = fam == AF_INET ? sizeof(sockaddr_in) : sizeof(sockaddr_in6);
        memcpy(d.addr, it->ifa_addr, d.addr_len);
        d.speed_mbps = read_speed(ifn);
        d.pci_path = read_pci_path(ifn);
        if (filt.active) {
            if (filt.match(ifn)) devs.push_back(d);
        } else if ((it->ifa_flags & IFF_LOOPBACK) || ifn.compare(0, 6, "docker") == 0) {
            fallback.push_back(d);
        } else {
            devs.push_back(d);
        }
    }
    freeifaddrs(list);
    if (devs.empty()) devs = fallback;
    return devs;
}

// ---- Lane ---------------------------------------------------------------------------------------------------------
Lane::~Lane() { stop(); }

void Lane::start(int fd, bool sender, Connection* owner, bool control) {
    fd_ = fd;
    owner_ = owner;
    control_ = control;
    th_ = std::thread(sender ? &Lane::run_sender : &Lane::run_receiver, this);
}

void Lane::push(const Item& it) {
    {
        std::lock_guard<std::mutex> g(mu_);
        q_.push_back(it);
    }
    cv_.notify_one();
}

bool Lane::next(Item& out) {
    std::unique_lock<std::mutex> g(mu_);
    cv_.wait(g, [&] { return stopping_ || !q_.empty(); });
    if (q_.empty()) return false;
    out = q_.front();
    q_.pop_front();
    return true;
}

void Lane::stop() {
    {
        std::lock_guard<std::mutex> g(mu_);
        if (stopping_ && !th_.joinable()) return;
        stopping_ = true;
    }
    cv_.notify_all();
    if (fd_ >= 0) ::shutdown(fd_, SHUT_RDWR);  // unblocks a thread parked in send()/recv()
    if (th_.joinable()) th_.join();
    if (fd_ >= 0) {
        ::close(fd_);
        fd_ = -1;
    }
    for (auto& it : q_)
        if (it.req) owner_->fail(it.req);
    q_.clear();
}

void Lane::run_sender() {
    bool dead = false;
    Item it;
    while (next(it)) {
        bool ok = !dead;
        if (ok && it.header) {
            uint64_t len = it.bytes;
            ok = write_all(fd_, reinterpret_cast<const char*>(&len), sizeof(len));
            if (ok && it.ptr && len) ok = write_all(fd_, it.ptr, len);  // inline payload
        } else if (ok) {
            ok = write_all(fd_, it.ptr, it.bytes);
        }
        if (ok) {
            Connection::complete_piece(it.req);
        } else {
            dead = true;  // the byte stream is out of step with the peer: everything after this fails too
            owner_->fail(it.req);
        }
    }
}

void Lane::run_receiver() {
    bool dead = false;
    Item it;
    while (next(it)) {
        bool ok = !dead;
        if (ok && it.header) {
            uint64_t len = 0;
            ok = read_all(fd_, reinterpret_cast<char*>(&len), sizeof(len));
            if (ok && len > it.req->capacity) ok = false;
            if (ok) {
                it.req->size = len;
                if (len <= kInlineBytes) {
                    if (len) ok = read_all(fd_, it.req->data, len);
                    if (ok) stats().inline_msgs.fetch_add(1, std::memory_order_relaxed);
                } else {
                    owner_->fan_out_recv(it.req, len);  // adds the chunk pieces before this piece completes
                }
            }
        } else if (ok) {
            ok = read_all(fd_, it.ptr, it.bytes);
        }
        if (ok) {
            Connection::complete_piece(it.req);
        } else {
            dead = true;
            owner_->fail(it.req);
        }
    }
}

// ---- Connection -----------------------------------------------------------------------------------------------------
Connection::~Connection() {
    ctrl_.stop();
    for (int i = 0; i < n_data_; ++i) data_[i].stop();
}

void Connection::adopt(int ctrl_fd, const std::vector<int>& data_fds) {
    n_data_ = static_cast<int>(data_fds.size());
    ctrl_.start(ctrl_fd, sender_, this, true);
    for (int i = 0; i < n_data_; ++i) data_[i].start(data_fds[i], sender_, this, false);
}

Request* Connection::grab() {
    for (auto& r : pool_) {
        bool expected = false;
        if (r.in_use.compare_exchange_strong(expected, true, std::memory_order_acq_rel)) {
            r.error.store(0, std::memory_order_relaxed);
            return &r;
        }
    }
    return nullptr;
}

void Connection::complete_piece(Request* req) { req->pending.fetch_sub(1, std::memory_order_release); }

void Connection::fail(Request* req) {
    stats().errors.fetch_add(1, std::memory_order_relaxed);
    req->error.store(1, std::memory_order_relaxed);
    req->pending.fetch_sub(1, std::memory_order_release);
}

Request* Connection::post_send(void* data, size_t size) {
    std::lock_guard<std::mutex> g(mu_);
    Request* r = grab();
    if (!r) return nullptr;
    r->is_send = true;
    r->data = static_cast<char*>(data);
    r->size = size;
    r->capacity = size;
    r->t_post_ns = now_ns();
    if (size <= kInlineBytes) {
        r->pending.store(1, std::memory_order_relaxed);
        stats().inline_msgs.fetch_add(1, std::memory_order_relaxed);
        ctrl_.push(Lane::Item{r, r->data, size, true});
        return r;
    }
    auto plan = plan_chunks(size, n_data_, cfg_.min_chunk, cursor_);
    cursor_ += static_cast<unsigned>(plan.size());
    r->pending.store(1 + static_cast<int>(plan.size()), std::memory_order_relaxed);
    stats().chunks.fetch_add(plan.size(), std::memory_order_relaxed);
    ctrl_.push(Lane::Item{r, nullptr, size, true});
    for (auto& c : plan) data_[c.stream].push(Lane::Item{r, r->data + c.offset, c.bytes, false});
    return r;
}

Request* Connection::post_recv(void* data, size_t capacity) {
    std::lock_guard<std::mutex> g(mu_);
    Request* r = grab();
    if (!r) return nullptr;
    r->is_send = false;
    r->data = static_cast<char*>(data);
    r->size = 0;
    r->capacity = capacity;
    r->t_post_ns = now_ns();
    r->pending.store(1, std::memory_order_relaxed);
    ctrl_.push(Lane::Item{r, nullptr, 0, true});
    return r;
}

void Connection::fan_out_recv(Request* req, size_t size) {
    // Runs on the control lane's thread only, in message order, so cursor_ evolves exactly like the sender's.
    auto plan = plan_chunks(size, n_data_, cfg_.min_chunk, cursor_);
    cursor_ += static_cast<unsigned>(plan.size());
    req->pending.fetch_add(static_cast<int>(plan.size()), std::memory_order_relaxed);
    stats().chunks.fetch_add(plan.size(), std::memory_order_relaxed);
    for (auto& c : plan) data_[c.stream].push(Lane::Item{req, req->data + c.offset, c.bytes, false});
}

// ---- Listener / connect ------------------------------------------------------------------------------------------------
Listener::~Listener() {
    if (fd_ >= 0) ::close(fd_);
    for (auto& p : pending_) {
        if (p.ctrl >= 0) ::close(p.ctrl);
        for (int fd : p.data)
            if (fd >= 0) ::close(fd);
    }
}

Listener* Listener::create(const NetDevice& dev, const Config& cfg, Handle* out) {
    sockaddr_storage ss;
    memset(&ss, 0, sizeof(ss));
    memcpy(&ss, dev.addr, dev.addr_len);
    const int fam = reinterpret_cast<sockaddr*>(&ss)->sa_family;
    if (fam == AF_INET)
        reinterpret_cast<sockaddr_in*>(&ss)->sin_port = 0;
    else
        reinterpret_cast<sockaddr_in6*>(&ss)->sin6_port = 0;
    int fd = ::socket(fam, SOCK_STREAM, 0);
    if (fd < 0) return nullptr;
    int one = 1;
    setsockopt(fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
    if (::bind(fd, reinterpret_cast<sockaddr*>(&ss), dev.addr_len) != 0 || ::listen(fd, 16384) != 0) {
        ::close(fd);
        return nullptr;
    }
    socklen_t len = sizeof(ss);
    getsockname(fd, reinterpret_cast<sockaddr*>(&ss), &len);
    set_blocking(fd, false);
    auto* l = new Listener();
    l->fd_ = fd;
    l->cfg_ = cfg;
    static std::atomic<uint64_t> next_id{1};
    memset(out, 0, sizeof(*out));
    out->magic = kMagic;
    out->nstreams = static_cast<uint32_t>(cfg.nstreams);
    out->listen_id = next_id.fetch_add(1);
    out->addr_len = len;
    memcpy(out->addr, &ss, len <= sizeof(out->addr) ? len : sizeof(out->addr));
    return l;
}

Connection* Listener::try_accept() {
    for (;;) {
        int s = ::accept(fd_, nullptr, nullptr);
        if (s < 0) break;  // EAGAIN: nothing more queued right now
        set_blocking(s, true);
        timeval tv{10, 0};
        setsockopt(s, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));  // the hello follows the connect immediately
        Hello h;
        if (!read_all(s, reinterpret_cast<char*>(&h), sizeof(h)) || h.magic != kMagic || h.nstreams != static_cast<uint32_t>(cfg_.nstreams) ||
            h.lane > h.nstreams) {
            ::close(s);
            continue;
        }
        timeval none{0, 0};
        setsockopt(s, SOL_SOCKET, SO_RCVTIMEO, &none, sizeof(none));
        tune_socket(s, cfg_);
        Pending* p = nullptr;
        for (auto& q : pending_)
            if (q.conn_id == h.conn_id) p = &q;
        if (!p) {
            pending_.push_back(Pending{h.conn_id, -1, std::vector<int>(cfg_.nstreams, -1), 0});
            p = &pending_.back();
        }
        int& slot = h.lane == 0 ? p->ctrl : p->data[h.lane - 1];
        if (slot >= 0) {
            ::close(s);
            continue;
        }
        slot = s;
        p->have++;
    }
    for (size_t i = 0; i < pending_.size(); ++i) {
        if (pending_[i].have == cfg_.nstreams + 1) {
            auto* c = new Connection(false, cfg_);
            c->adopt(pending_[i].ctrl, pending_[i].data);
            pending_.erase(pending_.begin() + static_cast<long>(i));
            return c;
        }
    }
    return nullptr;
}

Connection* connect_to(const Handle& h, const Config& cfg_in) {
    if (h.magic != kMagic) return nullptr;
    Config cfg = cfg_in;
    cfg.nstreams = static_cast<int>(h.nstreams);  // the listener's stream count is authoritative
    static std::atomic<uint64_t> counter{1};
    const uint64_t conn_id = (static_cast<uint64_t>(getpid()) << 40) ^ (now_ns() << 8) ^ counter.fetch_add(1);
    std::vector<int> fds;
    auto cleanup = [&] {
        for (int fd : fds) ::close(fd);
    };
    for (uint32_t lane = 0; lane <= h.nstreams; ++lane) {
        const int fam = reinterpret_cast<const sockaddr*>(h.addr)->sa_family;
        int fd = ::socket(fam, SOCK_STREAM, 0);
        if (fd < 0) {
            cleanup();
            return nullptr;
        }
        fds.push_back(fd);
        tune_socket(fd, cfg);
        int rc;
        do {
            rc = ::connect(fd, reinterpret_cast<const sockaddr*>(h.addr), h.addr_len);
        } while (rc != 0 && errno == EINTR);
        Hello hello{kMagic, lane, conn_id, h.nstreams, 0};
        if (rc != 0 || !write_all(fd, reinterpret_cast<const char*>(&hello), sizeof(hello))) {
            cleanup();
            return nullptr;
        }
    }
    auto* c = new Connection(true, cfg);
    c->adopt(fds[0], std::vector<int>(fds.begin() + 1, fds.end()));
    return c;
}

// ---- metrics push ---------------------------------------------------------------------------------------------------------
static void push_once(const std::string& host, const std::string& port, int rank) {
    addrinfo hints{}, *res = nullptr;
    hints.ai_socktype = SOCK_STREAM;
    if (getaddrinfo(host.c_str(), port.c_str(), &hints, &res) != 0 || !res) return;
    int fd = ::socket(res->ai_family, SOCK_STREAM, 0);
    if (fd >= 0) {
        timeval tv{2, 0};
        setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof(tv));
        setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
        if (::connect(fd, res->ai_addr, res->ai_addrlen) == 0) {
            std::string body = stats().prometheus(rank);
            char head[512];
            snprintf(head, sizeof(head),
                     "POST /metrics/job/bagua_net/rank/%d HTTP/1.1\r\nHost: %s\r\nContent-Type: text/plain\r\nContent-Length: %zu\r\nConnection: close\r\n\r\n",
                     rank, host.c_str(), body.size());
            std::string s = std::string(head) + body;
            if (write_all(fd, s.data(), s.size())) {
                char sink[256];
                while (::recv(fd, sink, sizeof(sink), 0) > 0) {
                }
            }
        }
        ::close(fd);
    }
    freeaddrinfo(res);
}

namespace {
struct Tracer {
    FILE* f = nullptr;
    std::mutex mu;
    bool first = true;
    int rank = 0;
    unsigned since_flush = 0;
    Tracer() {
        const char* prefix = getenv("BAGUA_NET_TRACE_FILE");
        const char* jaeger = getenv("BAGUA_NET_JAEGER_ADDRESS");
        rank = static_cast<int>(env_long("RANK", 0));
        std::string path;
        if (prefix && *prefix) {
            path = std::string(prefix) + "." + std::to_string(rank) + ".json";
        } else if (jaeger && *jaeger) {
            if (rank >= 8) return;  // the reference traces the first eight ranks only
            const char* tmp = getenv("TMPDIR");
            path = std::string(tmp && *tmp ? tmp : "/tmp") + "/bagua_net_trace." + std::to_string(rank) + ".json";
            fprintf(stderr, "bagua-net: BAGUA_NET_JAEGER_ADDRESS is set; spans are written as a trace-event file to %s\n", path.c_str());
        } else {
            return;
        }
        f = fopen(path.c_str(), "w");
        if (f) fputs("[\n", f);
    }
    ~Tracer() {
        if (f) {
            fputs("\n]\n", f);
            fclose(f);
        }
    }
};
Tracer& tracer() {
    static Tracer t;
    return t;
}
}  // namespace

void trace_span(bool is_send, uint64_t start_ns, uint64_t end_ns, size_t bytes, bool failed) {
    Tracer& t = tracer();
    if (!t.f) return;
    std::lock_guard<std::mutex> lk(t.mu);
    fprintf(t.f, "%s{\"name\":\"%s\",\"cat\":\"bagua_net\",\"ph\":\"X\",\"ts\":%.3f,\"dur\":%.3f,\"pid\":%d,\"tid\":%d,\"args\":{\"bytes\":%zu,\"ok\":%s}}",
            t.first ? "" : ",\n", is_send ? "isend" : "irecv", start_ns / 1e3, (end_ns - start_ns) / 1e3, t.rank, is_send ? 0 : 1, bytes,
            failed ? "false" : "true");
    t.first = false;
    if (++t.since_flush >= 256) {
        fflush(t.f);
        t.since_flush = 0;
    }
}

void trace_flush() {
    Tracer& t = tracer();
    if (!t.f) return;
    std::lock_guard<std::mutex> lk(t.mu);
    fflush(t.f);
    t.since_flush = 0;
}

void start_metrics_push_if_configured() {
    static std::once_flag once;
    std::call_once(once, [] {
        const char* addr = getenv("BAGUA_NET_PROMETHEUS_ADDRESS");
        if (!addr || !*addr) return;
        std::string a(addr);
        auto colon = a.rfind(':');
        if (colon == std::string::npos) return;
        std::string host = a.substr(0, colon), port = a.substr(colon + 1);
        int rank = static_cast<int>(env_long("RANK", 0));
        long period = env_long("BAGUA_NET_PROMETHEUS_PERIOD_S", 5);
        std::thread([host, port, rank, period] {
            for (;;) {
                push_once(host, port, rank);
                ::sleep(static_cast<unsigned>(period < 1 ? 1 : period));
            }
        }).detach();
    });
}

}  // namespace bagua_net
