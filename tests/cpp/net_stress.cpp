// Stress driver for the multi-stream TCP transport (csrc/net/net_engine.*), meant to run under -fsanitize=thread/address.
// Two connection pairs over loopback, each driven by its own sender and receiver thread the way NCCL's proxy threads
// would: random message sizes across the inline / single-chunk / multi-chunk classes, several requests in flight,
// payload verified by checksum, then an abrupt close with requests still pending.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

#include "net_engine.h"

using namespace bagua_net;

static int failures = 0;
#define EXPECT(c)                                                    \
    do {                                                             \
        if (!(c)) {                                                  \
            ++failures;                                              \
            fprintf(stderr, "FAILED %s:%d %s\n", __FILE__, __LINE__, #c); \
        }                                                            \
    } while (0)

static void fill(std::vector<char>& b, unsigned seed) {
    std::mt19937 g(seed);
    for (auto& c : b) c = static_cast<char>(g());
}

static void wait_done(Request* r) {
    while (r->pending.load(std::memory_order_acquire) > 0) std::this_thread::yield();
}

static void run_pair(Connection* tx, Connection* rx, unsigned seed, int n_msgs, std::atomic<int>* bad) {
    std::vector<size_t> sizes;
    std::mt19937 g(seed);
    const size_t classes[] = {0, 1, 4000, kInlineBytes, kInlineBytes + 1, 70000, 300000, 2000000};
    for (int i = 0; i < n_msgs; ++i) sizes.push_back(classes[g() % 8] + (g() % 3 == 0 ? g() % 1000 : 0));
    std::thread sender([&] {
        std::vector<std::vector<char>> bufs(n_msgs);
        std::vector<Request*> reqs;
        for (int i = 0; i < n_msgs; ++i) {
            bufs[i].resize(sizes[i]);
            fill(bufs[i], seed * 1000 + i);
            Request* r = nullptr;
            while (!(r = tx->post_send(bufs[i].data(), sizes[i]))) {  // pool exhausted: retire the oldest
                wait_done(reqs.front());
                reqs.front()->in_use.store(false, std::memory_order_release);
                reqs.erase(reqs.begin());
            }
            reqs.push_back(r);
        }
        for (auto* r : reqs) {
            wait_done(r);
            if (r->error.load()) bad->fetch_add(1);
            r->in_use.store(false, std::memory_order_release);
        }
    });
    std::thread receiver([&] {
        for (int i = 0; i < n_msgs; ++i) {
            std::vector<char> got(sizes[i] + 32, 0x5A), want(sizes[i]);
            fill(want, seed * 1000 + i);
            Request* r = nullptr;
            while (!(r = rx->post_recv(got.data(), got.size()))) std::this_thread::yield();
            wait_done(r);
            if (r->error.load() || r->size != sizes[i] || memcmp(got.data(), want.data(), sizes[i]) != 0 || got[sizes[i]] != 0x5A) bad->fetch_add(1);
            r->in_use.store(false, std::memory_order_release);
        }
    });
    sender.join();
    receiver.join();
}

int main() {
    setenv("NCCL_SOCKET_IFNAME", "lo", 1);
    Config cfg;
    cfg.nstreams = 3;
    cfg.min_chunk = 65536;
    auto devs = discover_devices();
    EXPECT(!devs.empty());
    if (devs.empty()) return 1;
    Handle h;
    Listener* l = Listener::create(devs[0], cfg, &h);
    EXPECT(l != nullptr);
    Connection* tx[2] = {connect_to(h, cfg), connect_to(h, cfg)};
    Connection* rx[2] = {nullptr, nullptr};
    for (int i = 0; i < 2; ++i)
        while (!(rx[i] = l->try_accept())) std::this_thread::yield();
    // accept order is arbitrary: find out which receiver belongs to tx[0] with a probe message
    char probe = 42, sink0 = 0, sink1 = 0;
    Request* ps = tx[0]->post_send(&probe, 1);
    Request* r0 = rx[0]->post_recv(&sink0, 1);
    Request* r1 = rx[1]->post_recv(&sink1, 1);
    wait_done(ps);
    ps->in_use = false;
    while (r0->pending.load() > 0 && r1->pending.load() > 0) std::this_thread::yield();
    const bool straight = r0->pending.load() == 0;
    Request* ps2 = tx[1]->post_send(&probe, 1);  // completes the other probe receive
    wait_done(ps2);
    ps2->in_use = false;
    wait_done(r0);
    wait_done(r1);
    r0->in_use = false;
    r1->in_use = false;
    EXPECT(sink0 == 42 && sink1 == 42);
    if (!straight) std::swap(rx[0], rx[1]);

    std::atomic<int> bad{0};
    std::thread a(run_pair, tx[0], rx[0], 1u, 150, &bad);
    std::thread b(run_pair, tx[1], rx[1], 2u, 150, &bad);
    a.join();
    b.join();
    EXPECT(bad.load() == 0);

    // abrupt shutdown with work still queued: receives that will never be matched + a large send nobody reads
    std::vector<char> big(8 << 20, 1), tmp(1024);
    rx[0]->post_recv(tmp.data(), tmp.size());
    tx[1]->post_send(big.data(), big.size());
    delete tx[0];
    delete rx[0];
    delete rx[1];
    delete tx[1];
    delete l;
    printf("net stress: %d failures\n", failures);
    return failures ? 1 : 0;
}
