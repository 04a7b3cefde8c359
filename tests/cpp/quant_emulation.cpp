// Host build of the SCALAR MinMaxUInt8 math of csrc/quant.cuh (make_quant / quantize / dequantize, the order-preserving float
// encoding, the bf16 / fp16 conversions around it).  Every CUDA intrinsic the header uses is a single IEEE-754 round-to-nearest
// operation, so defining them as the corresponding C++ operation (and compiling without FMA contraction) reproduces the device
// results bit for bit — which lets the CPU suite check the numerics of kernel code it cannot run: the quantised bytes against the
// python oracle (exact), the reciprocal-multiply decode against the division (at most one ulp) and the -DBAGUA_DEQUANT_IEEE_DIV
// build against the oracle (exact).
#include <cmath>
#include <cstdint>
#include <cstring>

#define BAGUA_QUANT_HOST_EMULATION
#define __device__
#define __forceinline__ inline

struct __half { _Float16 v; };
struct __nv_bfloat16 { uint16_t bits; };
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline float __half2float(__half h) { return static_cast<float>(h.v); }
static inline __half __float2half_rn(float f) { return __half{static_cast<_Float16>(f)}; }
static inline float __bfloat162float(__nv_bfloat16 b) { return __uint_as_float(static_cast<uint32_t>(b.bits) << 16); }
static inline __nv_bfloat16 __float2bfloat16_rn(float f) {   // round to nearest even, NaN kept quiet
    uint32_t u = __float_as_uint(f);
    if ((u & 0x7fffffffu) > 0x7f800000u) return __nv_bfloat16{static_cast<uint16_t>((u >> 16) | 0x40u)};
    u += 0x7fffu + ((u >> 16) & 1u);
    return __nv_bfloat16{static_cast<uint16_t>(u >> 16)};
}

#include "quant.cuh"

using namespace bagua::dev;

extern "C" {
// params: [scale, lower, upper, inv_scale]; q: n levels; deq: n decoded values (fp32)
void emu_minmax_uint8(const float* x, int n, float mn, float mx, uint8_t* q, float* deq, float* params) {
    const QuantParams p = make_quant(mn, mx);
    params[0] = p.scale, params[1] = p.lower, params[2] = p.upper, params[3] = p.inv_scale;
    for (int i = 0; i < n; ++i) {
        q[i] = quantize(x[i], p);
        deq[i] = dequantize(q[i], p);
    }
}
// decode through a 16-bit type as the kernels store it: kind 1 = bf16, 2 = fp16
void emu_decode_16bit(const uint8_t* q, int n, float mn, float mx, int kind, uint16_t* out) {
    const QuantParams p = make_quant(mn, mx);
    for (int i = 0; i < n; ++i) {
        const float v = dequantize(q[i], p);
        if (kind == 1) {
            out[i] = from_f32<__nv_bfloat16>(v).bits;
        } else {
            const __half h = from_f32<__half>(v);
            std::memcpy(&out[i], &h.v, 2);
        }
    }
}
uint32_t emu_f32_to_ordered(float f) { return f32_to_ordered(f); }
float emu_ordered_to_f32(uint32_t u) { return ordered_to_f32(u); }
}
