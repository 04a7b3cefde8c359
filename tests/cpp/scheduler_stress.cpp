// ThreadSanitizer stress of the C++ CommScheduler on the CPU backend: several producer threads mark tensors of many
// buckets ready in random order while the worker executes callback ops and a consumer waits; checks ordering and counts.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <memory>
#include <random>
#include <thread>
#include <vector>

#include "scheduler.h"

using namespace bagua;

// a native (non-blocking) op: eligible for inline issue on the marking thread
struct RecordOp final : CommOp {
    std::vector<int>* order;
    std::mutex* mu;
    int b;
    RecordOp(std::vector<int>* o, std::mutex* m, int bucket) : order(o), mu(m), b(bucket) {}
    const char* kind() const override { return "record"; }
    void run(Bucket&, StreamHandle, int) override {
        std::lock_guard<std::mutex> lk(*mu);
        order->push_back(b);
    }
};

int main() {
    constexpr int kBuckets = 16, kTensorsPerBucket = 8, kIters = 200;
    Backend be(4, -1, nullptr, 30.0);
    be.set_watchdog_fatal(false);
    be.set_record_spans(true);
    be.set_profile(true);  // per-bucket statistics are read concurrently with the worker below
    be.set_timeline(true); // ... and so is the per-bucket timeline (samples appended while the worker issues)
    std::vector<std::shared_ptr<Bucket>> buckets;
    std::vector<std::shared_ptr<Tensor>> tensors;
    std::vector<int> order;
    std::mutex order_mu;
    for (int b = 0; b < kBuckets; ++b) {
        std::vector<std::shared_ptr<Tensor>> ts;
        for (int t = 0; t < kTensorsPerBucket; ++t) {
            auto x = std::make_shared<Tensor>("b" + std::to_string(b) + "t" + std::to_string(t), 0x10000ull * (b + 1) + 64ull * t, 16, F32, -1);
            ts.push_back(x);
            tensors.push_back(x);
        }
        auto bk = std::make_shared<Bucket>("bucket" + std::to_string(b), ts);
        if (b % 3 == 0)  // "python" ops always go through the worker; the native ones are issued inline when the switch is on
            bk->append_op(std::make_shared<CallbackOp>([&order, &order_mu, b](const std::string&) {
                std::lock_guard<std::mutex> lk(order_mu);
                order.push_back(b);
            }));
        else
            bk->append_op(std::make_shared<RecordOp>(&order, &order_mu, b));
        buckets.push_back(bk);
    }
    be.register_ordered_buckets(buckets);
    int failures = 0;
    for (int it = 0; it < kIters; ++it) {
        be.set_inline((it & 1) != 0);  // alternate worker-only and inline issue; the order contract is the same
        std::vector<std::thread> producers;
        constexpr int kProducers = 4;
        for (int p = 0; p < kProducers; ++p) {
            producers.emplace_back([&, p, it] {
                std::mt19937 rng(1234 + 17 * it + p);
                std::vector<int> idx;
                for (size_t i = p; i < tensors.size(); i += kProducers) idx.push_back(static_cast<int>(i));
                std::shuffle(idx.begin(), idx.end(), rng);
                for (int i : idx) be.mark_communication_ready(tensors[i], nullptr);
            });
        }
        std::thread reader([&] { (void)be.bucket_stats(false); });
        for (auto& t : producers) t.join();
        size_t n = be.wait_pending_comm_ops(nullptr, true);
        reader.join();
        if (n != kBuckets) {
            std::fprintf(stderr, "iteration %d: waited for %zu buckets, expected %d\n", it, n, kBuckets);
            ++failures;
        }
        std::lock_guard<std::mutex> lk(order_mu);
        for (int b = 0; b < kBuckets; ++b)
            if (order[static_cast<size_t>(it) * kBuckets + b] != b) ++failures;  // strictly in registration order, every iteration
    }
    auto spans = be.pop_ready_spans();
    if (spans.empty()) ++failures;
    // ready spans and timeline samples are tagged with the pass over the registered order they belong to
    uint64_t max_iter = 0;
    for (auto& sp : spans) max_iter = std::max<uint64_t>(max_iter, sp.iteration);
    auto samples = be.pop_bucket_timeline();
    std::vector<int> per_iter(kIters, 0);
    for (auto& smp : samples) {
        if (smp.iteration >= static_cast<uint64_t>(kIters) || smp.start_ms < 0 || smp.device_ms < 0) ++failures;
        else per_iter[smp.iteration]++;
    }
    for (int it = 0; it < kIters; ++it)
        if (per_iter[it] != kBuckets) ++failures;
    if (samples.size() != static_cast<size_t>(kIters) * kBuckets || max_iter != static_cast<uint64_t>(kIters) - 1) {
        std::fprintf(stderr, "timeline: %zu samples (expected %d), last span iteration %llu (expected %d)\n", samples.size(), kIters * kBuckets,
                     static_cast<unsigned long long>(max_iter), kIters - 1);
        ++failures;
    }
    be.set_timeline(false);
    uint64_t launches = 0;
    for (auto& st : be.bucket_stats(false)) launches += st.count;
    if (launches != static_cast<uint64_t>(kIters) * kBuckets) {
        std::fprintf(stderr, "profile counted %llu launches, expected %d\n", static_cast<unsigned long long>(launches), kIters * kBuckets);
        ++failures;
    }
    if (be.inline_total() == 0) {
        std::fprintf(stderr, "inline issue never happened\n");
        ++failures;
    }
    be.shutdown();
    std::printf("scheduler stress: %d iterations, %zu ops, %d failures\n", kIters, order.size(), failures);
    return failures == 0 ? 0 : 1;
}
