// C++ autograd Functions for the fused NHWC conv-block epilogues (kernels: csrc/nhwc_fused.cu, reached through the C table
// `BaguaNhwcApi` of bagua_b200/_C.so). Same maths as the Python Functions in ops/nhwc.py; what changes is the host cost:
// a Python autograd.Function costs 26–68 µs per call on the launching thread (≈ 1.3 ms per VGG16 step for 26 calls), the C++
// one a few µs. Opt-in (BAGUA_NATIVE_NHWC=1) until the A/B on hardware has been run.
#include <ATen/ATen.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/extension.h>

#include <map>
#include <mutex>

namespace {

struct BaguaNhwcApi {  // must match csrc/kernels.h
    int (*bias_relu_fwd)(void* y, const void* bias, size_t rows, int C, int dtype, void* stream);
    int (*bias_relu_bwd)(const void* g, const void* y, void* gout, float* bias_grad, size_t rows, int C, int dtype, void* stream, void* bias_grad_out,
                         unsigned int* ticket);
    int (*pool_fwd)(const void* x, const void* bias, void* out, uint8_t* idx, int N, int H, int W, int C, int dtype, void* stream);
    int (*pool_bwd)(const void* g, const void* out, const uint8_t* idx, void* gin, float* bias_grad, int N, int H, int W, int C, int dtype, void* stream,
                    void* bias_grad_out, unsigned int* ticket);
    const char* (*last_error)();
};

const BaguaNhwcApi* g_api = nullptr;
bool g_finish_in_kernel = false;
constexpr int kMaxC = 2048;

void check(int rc, const char* what) { TORCH_CHECK(rc == 0, "bagua ", what, ": ", g_api->last_error()); }

int dtype_code(const at::Tensor& t) {  // bagua::DType
    if (t.scalar_type() == at::kBFloat16) return 4;
    if (t.scalar_type() == at::kHalf) return 1;
    TORCH_CHECK(false, "bagua fused NHWC epilogues need f16/bf16 tensors");
}

void* current_stream(const at::Tensor& t) {
    if (!t.is_cuda()) return nullptr;  // only reachable with a host test double of the kernel table
    return static_cast<void*>(c10::cuda::getCurrentCUDAStream(t.get_device()).stream());
}

// zeroed fp32[kMaxC] + ticket per (device, stream): the FIN kernels hand it back zeroed
at::Tensor workspace(const at::Tensor& like, void* stream) {
    static std::mutex mu;
    static std::map<std::pair<int, void*>, at::Tensor> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_pair(static_cast<int>(like.get_device()), stream);
    auto it = cache.find(key);
    if (it == cache.end()) it = cache.emplace(key, at::zeros({kMaxC + 4}, like.options().dtype(at::kFloat))).first;
    return it->second;
}

using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

struct BiasReLU : torch::autograd::Function<BiasReLU> {
    static at::Tensor forward(AutogradContext* ctx, at::Tensor y, at::Tensor bias) {
        const int64_t N = y.size(0), C = y.size(1), H = y.size(2), W = y.size(3);
        check(g_api->bias_relu_fwd(y.data_ptr(), bias.data_ptr(), static_cast<size_t>(N * H * W), static_cast<int>(C), dtype_code(y), current_stream(y)),
              "bias_relu_fwd");
        ctx->mark_dirty({y});
        ctx->save_for_backward({y});
        ctx->saved_data["bias_dtype"] = static_cast<int64_t>(bias.scalar_type());
        return y;
    }
    static variable_list backward(AutogradContext* ctx, variable_list grads) {
        const at::Tensor y = ctx->get_saved_variables()[0];
        const int64_t N = y.size(0), C = y.size(1), H = y.size(2), W = y.size(3);
        const at::Tensor g = grads[0].contiguous(at::MemoryFormat::ChannelsLast);
        at::Tensor gout = at::empty_like(g, g.options(), at::MemoryFormat::ChannelsLast);
        const auto bias_dtype = static_cast<at::ScalarType>(ctx->saved_data["bias_dtype"].toInt());
        void* stream = current_stream(y);
        if (g_finish_in_kernel) {
            at::Tensor ws = workspace(y, stream);
            at::Tensor bias_grad = at::empty({C}, y.options().dtype(bias_dtype));
            check(g_api->bias_relu_bwd(g.data_ptr(), y.data_ptr(), gout.data_ptr(), ws.data_ptr<float>(), static_cast<size_t>(N * H * W), static_cast<int>(C),
                                       dtype_code(y), stream, bias_grad.data_ptr(), reinterpret_cast<unsigned int*>(ws.data_ptr<float>() + kMaxC)),
                  "bias_relu_bwd");
            return {gout, bias_grad};
        }
        at::Tensor bg = at::zeros({C}, y.options().dtype(at::kFloat));
        check(g_api->bias_relu_bwd(g.data_ptr(), y.data_ptr(), gout.data_ptr(), bg.data_ptr<float>(), static_cast<size_t>(N * H * W), static_cast<int>(C),
                                   dtype_code(y), stream, nullptr, nullptr),
              "bias_relu_bwd");
        return {gout, bg.to(bias_dtype)};
    }
};

struct BiasReLUMaxPool2 : torch::autograd::Function<BiasReLUMaxPool2> {
    static at::Tensor forward(AutogradContext* ctx, at::Tensor x, at::Tensor bias) {
        const int64_t N = x.size(0), C = x.size(1), H = x.size(2), W = x.size(3);
        at::Tensor out = at::empty({N, C, H / 2, W / 2}, x.options(), at::MemoryFormat::ChannelsLast);
        at::Tensor idx = at::empty({N, C, H / 2, W / 2}, x.options().dtype(at::kByte), at::MemoryFormat::ChannelsLast);
        check(g_api->pool_fwd(x.data_ptr(), bias.data_ptr(), out.data_ptr(), idx.data_ptr<uint8_t>(), static_cast<int>(N), static_cast<int>(H),
                              static_cast<int>(W), static_cast<int>(C), dtype_code(x), current_stream(x)),
              "bias_relu_pool_fwd");
        ctx->save_for_backward({out, idx});
        ctx->saved_data["H"] = H;
        ctx->saved_data["W"] = W;
        ctx->saved_data["bias_dtype"] = static_cast<int64_t>(bias.scalar_type());
        return out;
    }
    static variable_list backward(AutogradContext* ctx, variable_list grads) {
        const auto saved = ctx->get_saved_variables();
        const at::Tensor out = saved[0], idx = saved[1];
        const int64_t N = out.size(0), C = out.size(1), H = ctx->saved_data["H"].toInt(), W = ctx->saved_data["W"].toInt();
        const at::Tensor g = grads[0].contiguous(at::MemoryFormat::ChannelsLast);
        at::Tensor gin = at::empty({N, C, H, W}, out.options(), at::MemoryFormat::ChannelsLast);
        const auto bias_dtype = static_cast<at::ScalarType>(ctx->saved_data["bias_dtype"].toInt());
        void* stream = current_stream(out);
        if (g_finish_in_kernel) {
            at::Tensor ws = workspace(out, stream);
            at::Tensor bias_grad = at::empty({C}, out.options().dtype(bias_dtype));
            check(g_api->pool_bwd(g.data_ptr(), out.data_ptr(), idx.data_ptr<uint8_t>(), gin.data_ptr(), ws.data_ptr<float>(), static_cast<int>(N),
                                  static_cast<int>(H), static_cast<int>(W), static_cast<int>(C), dtype_code(out), stream, bias_grad.data_ptr(),
                                  reinterpret_cast<unsigned int*>(ws.data_ptr<float>() + kMaxC)),
                  "bias_relu_pool_bwd");
            return {gin, bias_grad};
        }
        at::Tensor bg = at::zeros({C}, out.options().dtype(at::kFloat));
        check(g_api->pool_bwd(g.data_ptr(), out.data_ptr(), idx.data_ptr<uint8_t>(), gin.data_ptr(), bg.data_ptr<float>(), static_cast<int>(N),
                              static_cast<int>(H), static_cast<int>(W), static_cast<int>(C), dtype_code(out), stream, nullptr, nullptr),
              "bias_relu_pool_bwd");
        return {gin, bg.to(bias_dtype)};
    }
};

}  // namespace

void init_nhwc_bindings(py::module_& m) {
    m.def("nhwc_init", [](uint64_t api_ptr, bool finish_in_kernel) {
        g_api = reinterpret_cast<const BaguaNhwcApi*>(api_ptr);
        g_finish_in_kernel = finish_in_kernel;
    });
    m.def("nhwc_ready", [] { return g_api != nullptr; });
    m.def("bias_relu", [](at::Tensor y, at::Tensor bias) {
        TORCH_CHECK(g_api != nullptr, "call nhwc_init first");
        return BiasReLU::apply(std::move(y), std::move(bias));
    });
    m.def("bias_relu_maxpool2", [](at::Tensor x, at::Tensor bias) {
        TORCH_CHECK(g_api != nullptr, "call nhwc_init first");
        return BiasReLUMaxPool2::apply(std::move(x), std::move(bias));
    });
    // Whole conv block in one call from Python: bias-free convolution (cuDNN) + fused epilogue.
    m.def("conv_bias_relu", [](const at::Tensor& x, const at::Tensor& weight, const at::Tensor& bias, std::vector<int64_t> stride,
                               std::vector<int64_t> padding, std::vector<int64_t> dilation, int64_t groups, bool pool) {
        TORCH_CHECK(g_api != nullptr, "call nhwc_init first");
        at::Tensor y = at::conv2d(x, weight, c10::nullopt, stride, padding, dilation, groups);
        if (pool) return BiasReLUMaxPool2::apply(std::move(y), bias);
        return BiasReLU::apply(std::move(y), bias);
    });
}
