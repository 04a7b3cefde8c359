// PeerComm (per-group NVSwitch communicator state) and the native comm ops a bucket can run.
//
// Reference counterparts: BaguaSingleCommunicator (communicators/mod.rs:26-73, 474-489) and the six comm ops
// (comm_ops/*.rs). Here a "communicator" is nothing but symmetric signal pads + an abort/timeout flag: the
// collectives themselves are kernels (peer_kernels.cu, bytegrad_kernels.cu) that load/store peer memory.
#pragma once
#include <atomic>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "kernels.h"
#include "scheduler.h"

namespace bagua {

class PeerComm {
public:
    // flag_ptrs[p] = address (mapped in this process) of rank p's signal pad, ≥ kMaxCommBlocks*kMaxPeers*4 bytes, zeroed.
    PeerComm(int rank, int world, int device, const std::vector<uint64_t>& flag_ptrs, double timeout_s);
    ~PeerComm();
    PeerComm(const PeerComm&) = delete;
    const PeerCtx& ctx() const { return ctx_; }
    int rank() const { return ctx_.rank; }
    int world() const { return ctx_.world; }
    int device() const { return device_; }
    void abort();            // make every spinning kernel of this comm give up (ncclCommAbort analogue)
    void reset_abort();
    bool aborted() const;
    int error_code();        // synchronising read of the device error word (0 ok, 1 timeout, 2 aborted, 3 grid timeout, 4 protocol)
    // Non-synchronising read of the host-mapped mirror of the error word: what a failed kernel left behind (0 = healthy).
    int host_error() const { return __atomic_load_n(host_words_ + 1, __ATOMIC_ACQUIRE); }
    // Throws when a kernel of this communicator has failed: a collective that was skipped on some rank means diverged replicas,
    // so the failure is fatal for every later op (the reference's NCCL path blocks and then panics via its watchdog instead).
    void check_fatal(const char* what) const;
    void clear_error();
    void set_timeout(double seconds);
    uint32_t next_oneshot_parity() { return oneshot_calls_++ & 1u; }
    static size_t signal_pad_bytes() { return (static_cast<size_t>(kMaxCommBlocks) * kMaxPeers + kVoteWords) * sizeof(uint32_t); }

private:
    PeerCtx ctx_{};
    int device_;
    int* host_words_ = nullptr;  // host-mapped: [0] abort flag, [1] error mirror
    uint32_t oneshot_calls_ = 0;
};

struct SymmBuf {
    PeerBuf buf{};
    size_t bytes = 0;
    SymmBuf() = default;
    SymmBuf(const std::vector<uint64_t>& ptrs, uint64_t mc, size_t nbytes);
    bool has_multicast() const { return buf.mc != nullptr; }
};

struct LaunchCfg {
    int nblocks = 32;
    int nthreads = 512;
};

// Per-instance scratch of a quantised op (device allocations + host counters).
class QuantScratch {
public:
    QuantScratch(int device, size_t reduced_elems);
    ~QuantScratch();
    QuantScratch(const QuantScratch&) = delete;
    const ByteGradScratch& get() const { return s_; }

private:
    ByteGradScratch s_{};
    unsigned long long host_state_[2] = {0, 0};
    int device_;
};

class AllReduceOp final : public CommOp {
public:
    AllReduceOp(std::shared_ptr<PeerComm> comm, SymmBuf src, SymmBuf dst, size_t src_off, size_t dst_off, size_t bytes,
                int dtype, float scale, int variant, LaunchCfg cfg)
        : comm_(std::move(comm)), src_(src), dst_(dst), src_off_(src_off), dst_off_(dst_off), bytes_(bytes), dtype_(dtype),
          scale_(scale), variant_(variant), cfg_(cfg) {}
    const char* kind() const override { return variant_ == AR_MULTIMEM ? "allreduce_multimem" : "allreduce_twoshot"; }
    void run(Bucket&, StreamHandle stream, int device) override;
    void set_variant(int v, LaunchCfg cfg) { variant_ = v, cfg_ = cfg; }

private:
    std::shared_ptr<PeerComm> comm_;
    SymmBuf src_, dst_;
    size_t src_off_, dst_off_, bytes_;
    int dtype_;
    float scale_;
    int variant_;
    LaunchCfg cfg_;
};

class AllReduceOneShotOp final : public CommOp {
public:
    AllReduceOneShotOp(std::shared_ptr<PeerComm> comm, SymmBuf staging, size_t staging_off, size_t slot_bytes, uint64_t in, uint64_t out,
                       size_t bytes, int dtype, float scale, LaunchCfg cfg)
        : comm_(std::move(comm)), staging_(staging), staging_off_(staging_off), slot_bytes_(slot_bytes), in_(in), out_(out), bytes_(bytes), dtype_(dtype),
          scale_(scale), cfg_(cfg) {}
    const char* kind() const override { return "allreduce_oneshot"; }
    void run(Bucket&, StreamHandle stream, int device) override;
    bool step_invariant() const override { return false; }  // the staging half alternates with the communicator's call counter

private:
    std::shared_ptr<PeerComm> comm_;
    SymmBuf staging_;
    size_t staging_off_;
    size_t slot_bytes_;
    uint64_t in_, out_;
    size_t bytes_;
    int dtype_;
    float scale_;
    LaunchCfg cfg_;
};

// reduce-scatter(grads) → SGD on the owned shard → all-gather(weights), one kernel per bucket.
class AllReduceSgdOp final : public CommOp {
public:
    AllReduceSgdOp(std::shared_ptr<PeerComm> comm, SymmBuf grads, SymmBuf weights, size_t g_off, size_t w_off, size_t bytes, int dtype,
                   uint64_t master, uint64_t momentum, float scale, bool zero_grads, bool use_multimem, LaunchCfg cfg)
        : comm_(std::move(comm)), grads_(grads), weights_(weights), g_off_(g_off), w_off_(w_off), bytes_(bytes), dtype_(dtype),
          master_(master), momentum_(momentum), scale_(scale), zero_grads_(zero_grads), use_mc_(use_multimem), cfg_(cfg) {}
    const char* kind() const override { return "allreduce_sgd"; }
    void run(Bucket&, StreamHandle stream, int device) override;
    bool step_invariant() const override { return steps_ > 0; }  // only the very first launch differs (momentum initialisation)
    void set_hyper(float lr, float momentum, float dampening, float weight_decay, bool nesterov) {
        std::lock_guard<std::mutex> lk(mu_);
        hp_.lr = lr, hp_.momentum = momentum, hp_.dampening = dampening, hp_.weight_decay = weight_decay, hp_.nesterov = nesterov;
    }
    void set_grad_scale(float s) {
        std::lock_guard<std::mutex> lk(mu_);
        scale_ = s;
    }
    uint64_t steps() const { return steps_; }
    void set_steps(uint64_t s) { steps_ = s; }

private:
    std::shared_ptr<PeerComm> comm_;
    SymmBuf grads_, weights_;
    size_t g_off_, w_off_, bytes_;
    int dtype_;
    uint64_t master_, momentum_;
    float scale_;
    bool zero_grads_, use_mc_;
    LaunchCfg cfg_;
    SgdParams hp_{0.01f, 0.f, 0.f, 0.f, 0, 1};
    uint64_t steps_ = 0;
    std::mutex mu_;
};

// Intra-node halves of a hierarchical all-reduce (see reduce_scatter_kernel): ReduceScatterOp, then a callback op that
// all-reduces slice `rank` between nodes, then AllGatherOp.
class ReduceScatterOp final : public CommOp {
public:
    ReduceScatterOp(std::shared_ptr<PeerComm> comm, SymmBuf buf, size_t off, size_t bytes, int dtype, float scale, bool use_multimem, LaunchCfg cfg)
        : comm_(std::move(comm)), buf_(buf), off_(off), bytes_(bytes), dtype_(dtype), scale_(scale), use_mc_(use_multimem), cfg_(cfg) {}
    const char* kind() const override { return "reduce_scatter"; }
    void run(Bucket&, StreamHandle stream, int device) override;

private:
    std::shared_ptr<PeerComm> comm_;
    SymmBuf buf_;
    size_t off_, bytes_;
    int dtype_;
    float scale_;
    bool use_mc_;
    LaunchCfg cfg_;
};

class AllGatherOp final : public CommOp {
public:
    AllGatherOp(std::shared_ptr<PeerComm> comm, SymmBuf buf, size_t off, size_t bytes, int dtype, bool use_multimem, LaunchCfg cfg)
        : comm_(std::move(comm)), buf_(buf), off_(off), bytes_(bytes), dtype_(dtype), use_mc_(use_multimem), cfg_(cfg) {}
    const char* kind() const override { return "all_gather"; }
    void run(Bucket&, StreamHandle stream, int device) override;

private:
    std::shared_ptr<PeerComm> comm_;
    SymmBuf buf_;
    size_t off_, bytes_;
    int dtype_;
    bool use_mc_;
    LaunchCfg cfg_;
};

// The Adam / AdamW flavour of the fused bucket op; the step counter (bias correction) advances with every launch.
class AllReduceAdamOp final : public CommOp {
public:
    AllReduceAdamOp(std::shared_ptr<PeerComm> comm, SymmBuf grads, SymmBuf weights, size_t g_off, size_t w_off, size_t bytes, int dtype,
                    uint64_t master, uint64_t exp_avg, uint64_t exp_avg_sq, float scale, bool zero_grads, bool use_multimem, LaunchCfg cfg)
        : comm_(std::move(comm)), grads_(grads), weights_(weights), g_off_(g_off), w_off_(w_off), bytes_(bytes), dtype_(dtype),
          master_(master), m1_(exp_avg), m2_(exp_avg_sq), scale_(scale), zero_grads_(zero_grads), use_mc_(use_multimem), cfg_(cfg) {}
    const char* kind() const override { return "allreduce_adam"; }
    void run(Bucket&, StreamHandle stream, int device) override;
    bool step_invariant() const override { return false; }  // bias corrections are passed by value
    void set_hyper(float lr, float beta1, float beta2, float eps, float weight_decay, bool adamw) {
        std::lock_guard<std::mutex> lk(mu_);
        lr_ = lr, b1_ = beta1, b2_ = beta2, eps_ = eps, wd_ = weight_decay, adamw_ = adamw;
    }
    void set_grad_scale(float s) {
        std::lock_guard<std::mutex> lk(mu_);
        scale_ = s;
    }
    uint64_t steps() const { return steps_; }
    void set_steps(uint64_t s) { steps_ = s; }

private:
    std::shared_ptr<PeerComm> comm_;
    SymmBuf grads_, weights_;
    size_t g_off_, w_off_, bytes_;
    int dtype_;
    uint64_t master_, m1_, m2_;
    float scale_;
    bool zero_grads_, use_mc_;
    LaunchCfg cfg_;
    float lr_ = 1e-3f, b1_ = 0.9f, b2_ = 0.999f, eps_ = 1e-8f, wd_ = 0.f;
    bool adamw_ = false;
    uint64_t steps_ = 0;
    std::mutex mu_;
};

// Decentralized SGD, shift_one pairing (peer formula: comm_ops/decentralized_full_precision_synchronous.rs:81-85).
class PeerAverageOp final : public CommOp {
public:
    PeerAverageOp(std::shared_ptr<PeerComm> comm, SymmBuf weights, size_t off, uint64_t out, size_t bytes, int dtype, LaunchCfg cfg)
        : comm_(std::move(comm)), weights_(weights), off_(off), out_(out), bytes_(bytes), dtype_(dtype), cfg_(cfg) {}
    const char* kind() const override { return "peer_average_shift_one"; }
    void run(Bucket&, StreamHandle stream, int device) override;
    bool step_invariant() const override { return false; }  // the partner rotates with the step
    static int shift_one_peer(int rank, int nranks, int64_t step);
    int64_t step() const { return step_; }

private:
    std::shared_ptr<PeerComm> comm_;
    SymmBuf weights_;
    size_t off_;
    uint64_t out_;
    size_t bytes_;
    int dtype_;
    LaunchCfg cfg_;
    int64_t step_ = 0;
};

class ByteGradOp final : public CommOp {
public:
    ByteGradOp(std::shared_ptr<PeerComm> comm, uint64_t data, size_t numel, int dtype, SymmBuf inbox, size_t inbox_off, SymmBuf outbox,
               size_t outbox_off, bool average, LaunchCfg cfg);
    const char* kind() const override { return grad_ ? "qadam_momentum_bytegrad_fused" : "bytegrad_fused"; }
    void run(Bucket&, StreamHandle stream, int device) override;
    // QAdam compressed stage: the bucket holds the first moment; `grad` (same layout, same dtype) is folded in as
    // m = beta1*m + (1-beta1)*grad inside the kernel's first pass (0 switches it off).
    void set_momentum_source(uint64_t grad, float beta1) { grad_ = grad, beta1_ = beta1; }
    static size_t box_bytes(size_t numel, int nranks) { return static_cast<size_t>(nranks) * minmax_uint8_chunk_bytes(numel / nranks); }

private:
    std::shared_ptr<PeerComm> comm_;
    uint64_t data_;
    size_t numel_;
    int dtype_;
    SymmBuf inbox_, outbox_;
    size_t inbox_off_, outbox_off_;
    bool average_;
    LaunchCfg cfg_;
    std::unique_ptr<QuantScratch> scratch_;
    uint64_t grad_ = 0;
    float beta1_ = 0.f;
};

class LowPrecRingOp final : public CommOp {
public:
    LowPrecRingOp(std::shared_ptr<PeerComm> comm, uint64_t x, uint64_t w, uint64_t l, uint64_t r, size_t numel, int dtype, SymmBuf box,
                  size_t box_off, LaunchCfg cfg);
    const char* kind() const override { return "low_precision_ring_fused"; }
    void run(Bucket&, StreamHandle stream, int device) override;
    static size_t box_bytes(size_t numel) { return 6 * minmax_uint8_chunk_bytes(numel); }

private:
    std::shared_ptr<PeerComm> comm_;
    uint64_t x_, w_, l_, r_;
    size_t numel_;
    int dtype_;
    SymmBuf box_;
    size_t box_off_;
    LaunchCfg cfg_;
    std::unique_ptr<QuantScratch> scratch_;
};

// The device-side weight gate of asynchronous model averaging (see async_average_kernel): a few words in device memory that the
// trainer's stream acquires before a forward pass and releases after the optimizer step, and that the averaging kernel takes for
// its apply phase. Replaces the reference's host mutex + two host synchronisations per iteration.
class WeightGate {
public:
    explicit WeightGate(int device);
    ~WeightGate();
    WeightGate(const WeightGate&) = delete;
    uint32_t* words() const { return words_; }
    void acquire(StreamHandle stream, double timeout_s);  // enqueue on the trainer's stream
    void release(StreamHandle stream);
    uint32_t state();                                      // synchronising read (tests / diagnostics)

private:
    uint32_t* words_ = nullptr;
    int device_;
};

// One round of asynchronous model averaging = one kernel (comm op 5; reference decentralized_full_precision_asynchronous.rs:97-162).
class AsyncAverageOp final : public CommOp {
public:
    AsyncAverageOp(std::shared_ptr<PeerComm> comm, uint64_t weights, SymmBuf snap, size_t snap_off, SymmBuf avg, size_t avg_off, size_t bytes, int dtype,
                   std::shared_ptr<WeightGate> gate, double gate_timeout_s, bool use_multimem, LaunchCfg cfg);
    ~AsyncAverageOp() override;
    const char* kind() const override { return "async_model_average_fused"; }
    void run(Bucket&, StreamHandle stream, int device) override;
    bool step_invariant() const override { return false; }
    void abort() { go_.store(false); }   // this rank votes "stop" from the next round on
    void reset() {
        go_.store(true);
        *status_host_ = 1;
    }
    // Outcome of the last completed round (host-mapped word written by the kernel): 1 averaged, 0 a rank voted stop (all ranks see
    // the same), -1 failed. Valid once the round's ticket has completed.
    int status() const { return __atomic_load_n(status_host_, __ATOMIC_ACQUIRE); }
    uint64_t rounds() const { return seq_; }

private:
    std::shared_ptr<PeerComm> comm_;
    uint64_t w_;
    SymmBuf snap_, avg_;
    size_t snap_off_, avg_off_, bytes_;
    int dtype_;
    std::shared_ptr<WeightGate> gate_;
    double gate_timeout_s_;
    bool use_mc_;
    LaunchCfg cfg_;
    std::atomic<bool> go_{true};
    uint32_t seq_ = 0;
    int* status_host_ = nullptr;
    int* status_dev_ = nullptr;
};

// Plain device-to-device copy on the comm stream (snapshots / copy-back).
class CopyOp final : public CommOp {
public:
    CopyOp(uint64_t dst, uint64_t src, size_t bytes) : dst_(dst), src_(src), bytes_(bytes) {}
    const char* kind() const override { return "copy"; }
    void run(Bucket&, StreamHandle stream, int device) override;

private:
    uint64_t dst_, src_;
    size_t bytes_;
};

}  // namespace bagua
