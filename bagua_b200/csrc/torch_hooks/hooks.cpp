// Native post-accumulate-grad hooks (optional torch extension, host code only).
//
// The engine marks a parameter's gradient "ready for communication" from a hook that autograd runs right after the gradient
// has been accumulated. As a Python callable that costs ≈ 10–15 µs per parameter on the thread that also launches the backward
// kernels (VGG16: 32 parameters, BERT-large: ≈ 400 per step). Here the hook is a C++ object: no GIL, no Python frame — it calls
// straight into the scheduler of bagua_b200/_C.so through a function pointer handed over at start-up (the core stays torch-free,
// this file is the only one compiled against libtorch). Python is entered once per backward pass, to queue the engine's
// post-backward callback.
#include <torch/csrc/autograd/engine.h>
#include <torch/csrc/autograd/function_hook.h>
#include <torch/csrc/autograd/variable.h>
#include <torch/extension.h>

#include <atomic>
#include <memory>
#include <string>

namespace {

using MarkFn = void (*)(void* backend, void* tensor_handle, uint64_t stream);

struct EngineState {
    uint64_t backend = 0;
    MarkFn mark = nullptr;
    py::object post_backward;          // engine._real_post_backward_hook
    std::atomic<uint64_t> stream{0};   // comm-ready events are recorded on this stream (looked up once per forward)
    std::atomic<int> enabled{1};       // require_backward_grad_sync (no_sync() clears it)
    std::atomic<int> queued{0};        // post-backward callback already queued in this backward pass
    std::atomic<uint64_t> step{0};

    ~EngineState() {
        py::gil_scoped_acquire gil;
        post_backward = py::object();
    }
};

struct ReadyHook final : torch::autograd::PostAccumulateGradHook {
    std::shared_ptr<EngineState> st;
    uint64_t tensor_handle;
    uint64_t expected_grad_ptr;
    std::string name;

    ReadyHook(std::shared_ptr<EngineState> s, uint64_t handle, uint64_t grad_ptr, std::string n)
        : st(std::move(s)), tensor_handle(handle), expected_grad_ptr(grad_ptr), name(std::move(n)) {}

    void operator()(const at::Tensor& param) override {
        EngineState& s = *st;
        if (!s.enabled.load(std::memory_order_relaxed)) return;
        const uint64_t step = s.step.load(std::memory_order_relaxed);
        if (step < 4 || (step & 63) == 0) {  // the gradient must still be the bucket view registered with the scheduler
            const at::Tensor& g = param.grad();
            if (!g.defined() || reinterpret_cast<uint64_t>(g.data_ptr()) != expected_grad_ptr)
                throw std::runtime_error("bagua backend tensor data_ptr should match parameter grad (" + name + ")");
        }
        s.mark(reinterpret_cast<void*>(s.backend), reinterpret_cast<void*>(tensor_handle), s.stream.load(std::memory_order_relaxed));
        if (!s.queued.exchange(1)) {
            std::shared_ptr<EngineState> keep = st;
            torch::autograd::Engine::get_default_engine().queue_callback([keep] {
                py::gil_scoped_acquire gil;
                keep->post_backward();
            });
        }
    }
};

class HookState {
public:
    HookState(uint64_t backend, uint64_t mark_fn, py::object post_backward) : st_(std::make_shared<EngineState>()) {
        st_->backend = backend;
        st_->mark = reinterpret_cast<MarkFn>(mark_fn);
        st_->post_backward = std::move(post_backward);
    }
    void new_pass(uint64_t step, uint64_t stream, bool enabled) {
        st_->step.store(step, std::memory_order_relaxed);
        st_->stream.store(stream, std::memory_order_relaxed);
        st_->enabled.store(enabled ? 1 : 0, std::memory_order_relaxed);
        st_->queued.store(0, std::memory_order_relaxed);
    }
    void set_enabled(bool on) { st_->enabled.store(on ? 1 : 0, std::memory_order_relaxed); }
    // true exactly once per backward pass (shared with the Python-level hooks of parameters that keep them)
    bool try_queue() { return st_->queued.exchange(1) == 0; }
    // Installs the native hook unless the parameter already carries post-accumulate-grad hooks (those stay in charge).
    bool install(const at::Tensor& param, uint64_t tensor_handle, uint64_t expected_grad_ptr, const std::string& name) {
        if (torch::autograd::impl::post_acc_grad_hooks(param) != nullptr) return false;
        torch::autograd::impl::set_post_acc_grad_hooks(param, std::make_unique<ReadyHook>(st_, tensor_handle, expected_grad_ptr, name));
        return true;
    }
    static void remove(const at::Tensor& param) {
        auto& slot = torch::autograd::impl::post_acc_grad_hooks(param);
        if (slot != nullptr && dynamic_cast<ReadyHook*>(slot.get()) != nullptr) torch::autograd::impl::set_post_acc_grad_hooks(param, nullptr);
    }

private:
    std::shared_ptr<EngineState> st_;
};

}  // namespace

void init_nhwc_bindings(py::module_& m);  // nhwc_functions.cpp

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "bagua_b200 native autograd hooks and C++ autograd Functions";
    init_nhwc_bindings(m);
    py::class_<HookState>(m, "HookState")
        .def(py::init<uint64_t, uint64_t, py::object>(), py::arg("backend_ptr"), py::arg("mark_fn_ptr"), py::arg("post_backward"))
        .def("new_pass", &HookState::new_pass)
        .def("set_enabled", &HookState::set_enabled)
        .def("try_queue", &HookState::try_queue)
        .def("install", &HookState::install)
        .def_static("remove", &HookState::remove);
}
