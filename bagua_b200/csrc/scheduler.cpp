#include "scheduler.h"
#include "log.h"

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cstdlib>
#include <set>
#include <sstream>

namespace bagua {

namespace {
inline cudaStream_t S(StreamHandle s) { return reinterpret_cast<cudaStream_t>(s); }
inline cudaEvent_t E(EventHandle e) { return reinterpret_cast<cudaEvent_t>(e); }
inline int64_t now_ns() {
    return std::chrono::duration_cast<std::chrono::nanoseconds>(
               std::chrono::steady_clock::now().time_since_epoch())
        .count();
}
}  // namespace

void CallbackOp::run(Bucket& bucket, StreamHandle, int) { fn_(bucket.name()); }

Bucket::Bucket(std::string name, std::vector<std::shared_ptr<Tensor>> tensors)
    : name_(std::move(name)), tensors_(std::move(tensors)) {
    if (tensors_.empty()) throw std::invalid_argument("bagua: bucket '" + name_ + "' has no tensors");
    recompute_layout();
}

void Bucket::recompute_layout() {
    dtype_ = tensors_.front()->dtype();
    device_ = tensors_.front()->device();
    numel_ = 0;
    bytes_ = 0;
    contiguous_ = true;
    flat_ptr_ = tensors_.front()->data_ptr();
    uint64_t expect = flat_ptr_;
    for (auto& t : tensors_) {
        // A bucket never mixes dtypes or devices (reference: datatypes/mod.rs:1135-1147).
        if (t->dtype() != dtype_)
            throw std::invalid_argument("bagua: bucket '" + name_ + "' mixes dtypes (tensor " + t->name() + ")");
        if (t->device() != device_)
            throw std::invalid_argument("bagua: bucket '" + name_ + "' mixes devices (tensor " + t->name() + ")");
        if (t->data_ptr() != expect) contiguous_ = false;
        expect = t->data_ptr() + t->bytes();
        numel_ += t->numel();
        bytes_ += t->bytes();
    }
    if (!contiguous_) flat_ptr_ = 0;
}

std::string Bucket::describe_ops() const {
    std::ostringstream os;
    os << "bucket " << name_ << " [";
    for (size_t i = 0; i < ops_.size(); ++i) os << (i ? ", " : "") << ops_[i]->kind();
    os << "]";
    return os.str();
}

bool Bucket::ready_for_comm() const {
    for (auto& t : tensors_)
        if (!t->always_ready_ && !t->ready()) return false;
    return true;
}

void Bucket::reset_comm_ready() {
    for (auto& t : tensors_) t->ready_.store(false, std::memory_order_release);
}

Backend::Backend(size_t channel_cap, int device_id, StreamHandle comm_stream, double watchdog_timeout_s)
    : cap_(std::max<size_t>(channel_cap, 1)), device_(device_id), stream_(comm_stream), timeout_s_(watchdog_timeout_s) {
    BAGUA_LOG(INFO, "comm backend created: device %d, channel capacity %zu, watchdog %.0f s", device_, cap_, timeout_s_);
    worker_ = std::thread([this] { worker_loop(); });
    watchdog_ = std::thread([this] { watchdog_loop(); });
}

Backend::~Backend() { shutdown(); }

void Backend::shutdown() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        if (stop_) return;
        stop_ = true;
    }
    cv_worker_.notify_all();
    cv_watch_.notify_all();
    cv_space_.notify_all();
    cv_done_.notify_all();
    if (worker_.joinable()) worker_.join();
    if (watchdog_.joinable()) watchdog_.join();
    // Events are intentionally leaked at process teardown if the context is already gone.
    if (device_ >= 0) {
        std::lock_guard<std::mutex> lk(pool_mu_);
        for (auto e : event_pool_) cudaEventDestroy(E(e));
        event_pool_.clear();
        std::lock_guard<std::mutex> plk(prof_mu_);
        for (auto e : timing_pool_) cudaEventDestroy(E(e));
        timing_pool_.clear();
        if (timeline_ref_) cudaEventDestroy(E(timeline_ref_));
        timeline_ref_ = nullptr;
        for (auto& smp : prof_pending_) {
            if (smp.start) cudaEventDestroy(E(smp.start));
            if (smp.stop) cudaEventDestroy(E(smp.stop));
        }
        prof_pending_.clear();
    }
}

EventHandle Backend::acquire_event() {
    {
        std::lock_guard<std::mutex> lk(pool_mu_);
        if (!event_pool_.empty()) {
            auto e = event_pool_.back();
            event_pool_.pop_back();
            return e;
        }
    }
    cudaEvent_t e;
    BAGUA_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return reinterpret_cast<EventHandle>(e);
}

EventHandle Backend::acquire_timing_event() {
    {
        std::lock_guard<std::mutex> lk(prof_mu_);
        if (!timing_pool_.empty()) {
            auto e = timing_pool_.back();
            timing_pool_.pop_back();
            return e;
        }
    }
    cudaEvent_t e;
    BAGUA_CUDA_CHECK(cudaEventCreate(&e));
    return reinterpret_cast<EventHandle>(e);
}

std::vector<BucketStat> Backend::bucket_stats(bool reset) {
    std::lock_guard<std::mutex> lk(prof_mu_);
    while (!prof_pending_.empty()) {
        ProfSample& smp = prof_pending_.front();
        double ms = smp.host_ms;
        if (smp.stop) {
            if (cudaEventQuery(E(smp.stop)) != cudaSuccess) {
                (void)cudaGetLastError();  // cudaErrorNotReady is expected: the kernels of this bucket are still running
                break;
            }
            float f = 0.f;
            if (cudaEventElapsedTime(&f, E(smp.start), E(smp.stop)) != cudaSuccess) {
                (void)cudaGetLastError();
                f = 0.f;
            }
            ms = f;
            if (timeline_.load(std::memory_order_relaxed) && timeline_ref_ && timeline_samples_.size() < kTimelineCap) {
                float off = 0.f;
                if (cudaEventElapsedTime(&off, E(timeline_ref_), E(smp.start)) != cudaSuccess) {
                    (void)cudaGetLastError();   // the reference event belongs to a later set_timeline(true): sample predates it
                    off = -1.f;
                }
                if (off >= 0.f) timeline_samples_.push_back(BucketSample{smp.name, smp.iteration, smp.issue_ns, off, ms, smp.queue_ms});
            }
            timing_pool_.push_back(smp.start);
            timing_pool_.push_back(smp.stop);
        } else if (timeline_.load(std::memory_order_relaxed) && timeline_samples_.size() < kTimelineCap && smp.issue_ns >= timeline_ref_ns_) {
            timeline_samples_.push_back(BucketSample{smp.name, smp.iteration, smp.issue_ns, (smp.issue_ns - timeline_ref_ns_) / 1e6, ms, smp.queue_ms});
        }
        auto it = prof_stats_.find(smp.name);
        if (it != prof_stats_.end()) {
            BucketStat& st = it->second;
            st.count++;
            st.total_ms += ms;
            st.queue_ms += smp.queue_ms;
            if (ms > st.max_ms) st.max_ms = ms;
        }
        prof_pending_.pop_front();
    }
    std::vector<BucketStat> out;
    for (auto& n : prof_order_) {
        auto it = prof_stats_.find(n);
        if (it != prof_stats_.end()) out.push_back(it->second);
    }
    if (reset)
        for (auto& kv : prof_stats_) kv.second.count = 0, kv.second.total_ms = kv.second.max_ms = kv.second.queue_ms = 0.0;
    return out;
}

void Backend::release_event(EventHandle e) {
    if (!e) return;
    std::lock_guard<std::mutex> lk(pool_mu_);
    event_pool_.push_back(e);
}

void Backend::register_ordered_buckets(std::vector<std::shared_ptr<Bucket>> buckets) {
    {
        std::lock_guard<std::mutex> plk(prof_mu_);
        prof_stats_.clear();
        prof_order_.clear();
        for (auto& b : buckets) {
            BucketStat st;
            st.name = b->name();
            st.ops = b->describe_ops();
            st.bytes = b->bytes();
            prof_stats_[st.name] = st;
            prof_order_.push_back(st.name);
        }
    }
    // Same sanity checks as the reference (lib.rs:282-292): tensor names and storage must be unique.
    std::set<std::string> names;
    std::set<uint64_t> ptrs;
    for (auto& b : buckets) {
        for (auto& t : b->tensors()) {
            if (!names.insert(t->name()).second)
                throw std::invalid_argument("bagua: duplicated tensor name '" + t->name() + "' in registered buckets");
            if (t->numel() > 0 && !ptrs.insert(t->data_ptr()).second)
                throw std::invalid_argument("bagua: tensor '" + t->name() + "' aliases the storage of another registered tensor");
        }
    }
    std::unique_lock<std::mutex> lk(mu_);
    // Let everything already scheduled be issued before the program changes.
    cv_done_.wait(lk, [this] { return (queue_.empty() && !in_flight_) || stop_; });
    ordered_.clear();
    owner_.clear();
    first_bucket_ = buckets.empty() ? nullptr : buckets.front().get();
    for (auto& b : buckets) {
        b->pending_ = 0;
        b->producer_stream_set_ = false;
        b->user_events_.clear();
        for (auto& t : b->tensors()) {
            t->ready_.store(false);
            if (!t->always_ready_) b->pending_++;
            owner_[t.get()] = b;
        }
        ordered_.push_back(b);
    }
}

void Backend::on_tensor_ready_locked(const std::shared_ptr<Tensor>& t, std::unique_lock<std::mutex>& lk) {
    if (record_spans_.load(std::memory_order_relaxed) && spans_.size() < (1u << 16))
        spans_.push_back(ReadySpan{t->name(), now_ns(), iteration_});
    schedule_locked(lk);
}

void Backend::mark_communication_ready(const std::shared_ptr<Tensor>& t, EventHandle ready_event) {
    std::unique_lock<std::mutex> lk(mu_);
    auto it = owner_.find(t.get());
    if (it == owner_.end()) throw std::invalid_argument("bagua: tensor '" + t->name() + "' is not registered in any bucket");
    auto& b = it->second;
    if (ready_event) b->user_events_.push_back(ready_event);
    if (!t->ready_.exchange(true) && !t->always_ready_ && b->pending_ > 0) b->pending_--;
    on_tensor_ready_locked(t, lk);
}

void Backend::mark_ready_on_stream(const std::shared_ptr<Tensor>& t, StreamHandle producer) {
    std::unique_lock<std::mutex> lk(mu_);
    auto it = owner_.find(t.get());
    if (it == owner_.end()) throw std::invalid_argument("bagua: tensor '" + t->name() + "' is not registered in any bucket");
    auto& b = it->second;
    const bool newly = !t->ready_.exchange(true) && !t->always_ready_;
    if (newly && b->pending_ > 0) b->pending_--;
    if (device_ >= 0) {
        if (b->producer_stream_set_ && b->producer_stream_ != producer) {
            // Producer stream changed inside the bucket: fence the previous one now.
            EventHandle e = acquire_event();
            BAGUA_CUDA_CHECK(cudaEventRecord(E(e), S(b->producer_stream_)));
            b->user_events_.push_back(reinterpret_cast<EventHandle>(reinterpret_cast<uintptr_t>(e) | 1u));
        }
        b->producer_stream_ = producer;
        b->producer_stream_set_ = true;
        if (b->pending_ == 0 && newly) {
            EventHandle e = acquire_event();
            BAGUA_CUDA_CHECK(cudaEventRecord(E(e), S(producer)));
            // low bit tags pool-owned events (handles are pointers, ≥ 8-byte aligned)
            b->user_events_.push_back(reinterpret_cast<EventHandle>(reinterpret_cast<uintptr_t>(e) | 1u));
            b->producer_stream_set_ = false;
        }
    }
    on_tensor_ready_locked(t, lk);
}

void Backend::schedule_locked(std::unique_lock<std::mutex>& lk) {
    // Strictly in registration order: only the front of the deque is ever examined
    // (reference lib.rs:300-319) so every rank issues the same collective sequence.
    while (!ordered_.empty() && ordered_.front()->ready_for_comm()) {
        auto b = ordered_.front();
        // Bounded channel: block the marking thread while the worker is `cap_` buckets behind.
        cv_space_.wait(lk, [this] { return queue_.size() < cap_ || stop_; });
        if (stop_) return;
        if (ordered_.empty() || ordered_.front() != b || !b->ready_for_comm()) continue;
        ordered_.pop_front();
        b->reset_comm_ready();
        b->pending_ = 0;
        for (auto& t : b->tensors())
            if (!t->always_ready_) b->pending_++;
        ordered_.push_back(b);
        auto tk = std::make_shared<Ticket>();
        tk->bucket = b;
        for (auto e : b->user_events_) {
            auto raw = reinterpret_cast<uintptr_t>(e);
            if (raw & 1u) {
                auto clean = reinterpret_cast<EventHandle>(raw & ~uintptr_t(1));
                tk->wait_events.push_back(clean);
                tk->pooled_waits.push_back(clean);
            } else {
                tk->wait_events.push_back(e);
            }
        }
        b->user_events_.clear();
        b->producer_stream_set_ = false;
        tk->t_sched = std::chrono::steady_clock::now();
        tk->iteration = iteration_;
        bool issue_inline = inline_.load(std::memory_order_relaxed) && queue_.empty() && !in_flight_;
        if (issue_inline)
            for (auto& op : b->ops()) issue_inline = issue_inline && !op->host_blocking();
        if (issue_inline) {
            // the marking thread issues the bucket itself (asynchronous launches only), still under mu_: nothing can overtake it
            BAGUA_LOG(DEBUG, "bucket %s ready: issued inline (%zu wait events)", b->name().c_str(), tk->wait_events.size());
            not_waited_.push_back(tk);
            scheduled_total_++;
            inline_total_++;
            if (ordered_.front().get() == first_bucket_) iteration_++;  // the last bucket of the registered order: the step's program is complete
            in_flight_ = tk;
            int prev_dev = -1;
            if (device_ >= 0 && cudaGetDevice(&prev_dev) == cudaSuccess && prev_dev != device_) cudaSetDevice(device_);
            issue_ticket(tk);
            if (device_ >= 0 && prev_dev >= 0 && prev_dev != device_) cudaSetDevice(prev_dev);
            tk->issued = true;
            in_flight_.reset();
            cv_done_.notify_all();
            continue;
        }
        BAGUA_LOG(DEBUG, "bucket %s ready: scheduled (%zu in queue, %zu wait events)", b->name().c_str(), queue_.size() + 1, tk->wait_events.size());
        queue_.push_back(tk);
        not_waited_.push_back(tk);
        scheduled_total_++;
        if (ordered_.front().get() == first_bucket_) iteration_++;  // wrapped around the whole registered order (also the single-bucket case)
        cv_worker_.notify_one();
    }
}

void Backend::issue_ticket(const std::shared_ptr<Ticket>& tk) {
    try {
        if (device_ >= 0)
            for (auto e : tk->wait_events) BAGUA_CUDA_CHECK(cudaStreamWaitEvent(S(stream_), E(e), 0));
        const bool prof = profile_.load(std::memory_order_relaxed);
        ProfSample smp;
        std::chrono::steady_clock::time_point t_issue;
        if (prof) {
            t_issue = std::chrono::steady_clock::now();
            smp.issue_ns = std::chrono::duration_cast<std::chrono::nanoseconds>(t_issue.time_since_epoch()).count();
            smp.iteration = tk->iteration;
            smp.name = tk->bucket->name();
            smp.queue_ms = std::chrono::duration<double, std::milli>(t_issue - tk->t_sched).count();
            if (device_ >= 0) {
                smp.start = acquire_timing_event();
                smp.stop = acquire_timing_event();
                BAGUA_CUDA_CHECK(cudaEventRecord(E(smp.start), S(stream_)));
            }
        }
        nvtxRangePushA(tk->bucket->name().c_str());
        for (auto& op : tk->bucket->ops()) {
            nvtxRangePushA(op->kind());
            BAGUA_LOG(TRACE, "bucket %s: issuing op %s", tk->bucket->name().c_str(), op->kind());
            op->run(*tk->bucket, stream_, device_);
            nvtxRangePop();
        }
        nvtxRangePop();
        if (prof) {
            if (device_ >= 0)
                BAGUA_CUDA_CHECK(cudaEventRecord(E(smp.stop), S(stream_)));
            else
                smp.host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_issue).count();
            std::lock_guard<std::mutex> plk(prof_mu_);
            prof_pending_.push_back(std::move(smp));
        }
        if (device_ >= 0) {
            tk->done_event = acquire_event();
            BAGUA_CUDA_CHECK(cudaEventRecord(E(tk->done_event), S(stream_)));
        }
    } catch (const std::exception& ex) {
        tk->failed = true;
        tk->error = ex.what();
        BAGUA_LOG(ERROR, "communication of %s failed: %s", tk->bucket->describe_ops().c_str(), ex.what());
    } catch (...) {
        tk->failed = true;
        tk->error = "unknown error in comm op";
    }
    for (auto e : tk->pooled_waits) release_event(e);
}

void Backend::worker_loop() {
    bool device_set = false;
    for (;;) {
        std::shared_ptr<Ticket> tk;
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_worker_.wait(lk, [this] { return !queue_.empty() || stop_; });
            if (queue_.empty()) return;  // stop_ requested and nothing left
            tk = queue_.front();
            queue_.pop_front();
            in_flight_ = tk;
            cv_space_.notify_all();
        }
        if (device_ >= 0 && !device_set) {
            cudaError_t de = cudaSetDevice(device_);
            if (de != cudaSuccess) BAGUA_LOG(ERROR, "cudaSetDevice(%d) failed on the comm worker: %s", device_, cudaGetErrorString(de));
            device_set = true;
        }
        issue_ticket(tk);
        {
            std::lock_guard<std::mutex> lk(mu_);
            tk->issued = true;
            in_flight_.reset();
        }
        cv_done_.notify_all();
    }
}

void Backend::watchdog_loop() {
    // Reference: comm_monitor panics after 300 s without a finished op (lib.rs:255-265).
    std::unique_lock<std::mutex> lk(mu_);
    while (!stop_) {
        cv_watch_.wait_for(lk, std::chrono::milliseconds(200));
        if (stop_) break;
        std::shared_ptr<Ticket> oldest = in_flight_ ? in_flight_ : (queue_.empty() ? nullptr : queue_.front());
        if (!oldest) continue;
        double waited = std::chrono::duration<double>(std::chrono::steady_clock::now() - oldest->t_sched).count();
        if (waited > timeout_s_ && watchdog_error_.empty()) {
            std::ostringstream os;
            os << "bagua watchdog: communication of " << oldest->bucket->describe_ops() << " has not been issued for "
               << waited << " s (limit " << timeout_s_ << " s)";
            watchdog_error_ = os.str();
            std::fprintf(stderr, "%s\n", watchdog_error_.c_str());
            if (watchdog_fatal_.load()) {
                std::fflush(stderr);
                std::_Exit(1);
            }
            cv_done_.notify_all();
        }
    }
}

bool Backend::graph_capturable() {
    std::lock_guard<std::mutex> lk(mu_);
    for (auto& b : ordered_)
        for (auto& op : b->ops())
            if (op->host_blocking() || !op->step_invariant()) return false;
    return true;
}

std::string Backend::watchdog_error() {
    std::lock_guard<std::mutex> lk(mu_);
    return watchdog_error_;
}

size_t Backend::pending_count() {
    std::lock_guard<std::mutex> lk(mu_);
    return not_waited_.size();
}

size_t Backend::wait_pending_comm_ops(StreamHandle consumer, bool host_sync) {
    size_t n = 0;
    EventHandle last = nullptr;
    std::string err;
    {
        std::unique_lock<std::mutex> lk(mu_);
        while (!not_waited_.empty()) {
            auto tk = not_waited_.front();
            cv_done_.wait(lk, [&] { return tk->issued || stop_ || !watchdog_error_.empty(); });
            if (!tk->issued) {
                err = watchdog_error_.empty() ? "bagua: backend stopped while waiting" : watchdog_error_;
                break;
            }
            not_waited_.pop_front();
            if (tk->failed && err.empty()) err = "bagua comm op failed on " + tk->bucket->describe_ops() + ": " + tk->error;
            if (tk->done_event) {
                if (last) release_event(last);
                last = tk->done_event;  // comm stream is in-order: the newest event dominates
            }
            ++n;
        }
    }
    if (last) {
        if (host_sync)
            BAGUA_CUDA_CHECK(cudaEventSynchronize(E(last)));
        else
            BAGUA_CUDA_CHECK(cudaStreamWaitEvent(S(consumer), E(last), 0));
        release_event(last);
    }
    if (!err.empty()) throw std::runtime_error(err);
    return n;
}

void Backend::set_timeline(bool on) {
    if (on) {
        std::lock_guard<std::mutex> lk(prof_mu_);
        timeline_samples_.clear();
        timeline_ref_ns_ = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
        if (device_ >= 0) {
            if (!timeline_ref_) {
                cudaEvent_t e;
                BAGUA_CUDA_CHECK(cudaEventCreate(&e));
                timeline_ref_ = reinterpret_cast<EventHandle>(e);
            }
            BAGUA_CUDA_CHECK(cudaEventRecord(E(timeline_ref_), S(stream_)));
        }
    }
    if (on && !timeline_.load()) profile_before_timeline_ = profile_.load();
    timeline_ = on;
    profile_ = on ? true : profile_before_timeline_;   // switching the timeline off leaves a comm_profile() the user had on untouched
}

std::vector<BucketSample> Backend::pop_bucket_timeline() {
    (void)bucket_stats(false);   // resolves the samples whose kernels have finished
    std::lock_guard<std::mutex> lk(prof_mu_);
    std::vector<BucketSample> out;
    out.swap(timeline_samples_);
    return out;
}

double Backend::timeline_ms_of_event(uint64_t cuda_event_ptr) {
    if (device_ < 0 || !timeline_.load() || !timeline_ref_ || cuda_event_ptr == 0) return -1.0;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, E(timeline_ref_), reinterpret_cast<cudaEvent_t>(cuda_event_ptr)) != cudaSuccess) {
        (void)cudaGetLastError();
        return -1.0;
    }
    return ms;
}

std::vector<ReadySpan> Backend::pop_ready_spans() {
    std::lock_guard<std::mutex> lk(mu_);
    std::vector<ReadySpan> out;
    out.swap(spans_);
    return out;
}

}  // namespace bagua
