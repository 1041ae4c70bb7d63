// Multi-tensor Adam step: the optimizer update of the reference's training loop
// (/root/reference/Main.py:192 `torch.optim.Adam(params, lr, weight_decay)`,
// /root/reference/Helpers/TrainTestHelper.py:142 `optimizer.step()`) over up to kAdamMaxTensors
// parameter tensors per launch.  The arithmetic is torch's fused CUDA Adam, operation for
// operation (torch/include/ATen/native/cuda/fused_adam_utils.cuh `adam_math`, ADAM_MODE::ORIGINAL,
// no amsgrad), with every rounding made explicit so that the results are bit-identical:
//     g   += wd * p                                (wd != 0)
//     m    = fma(b1, m, fma(-b1, g, g))
//     v    = fma(b2, v, fma(-b2, g*g, g*g))
//     p   -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// t = step + 1 is read from the tensor's device-resident fp32 step counter (graph-capturable: no
// host value changes between replays); a second tiny kernel increments the counters.
//
// Roofline: HBM, 28 bytes per element (read p, g, m, v; write p, m, v).
#include "common.cuh"

namespace ihg {

constexpr int kAdamMaxTensors = 24;
constexpr int kAdamThreads = 512;
constexpr int kAdamChunk = 65536;          // elements per block

struct AdamArgs {
    float* p[kAdamMaxTensors];
    const float* g[kAdamMaxTensors];
    float* m[kAdamMaxTensors];
    float* v[kAdamMaxTensors];
    float* step[kAdamMaxTensors];
    int64_t numel[kAdamMaxTensors];
    int32_t chunk0[kAdamMaxTensors + 1];   // first block of every tensor
    int32_t n;
};

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float wd, float b1, float b2,
                                          float step_size, float bc2_sqrt, float eps) {
    if (wd != 0.f) g = __fmaf_rn(p, wd, g);      // `grad += param * weight_decay`, contracted by nvcc in torch's build
    m = __fmaf_rn(b1, m, __fmaf_rn(-b1, g, g));
    const float gg = __fmul_rn(g, g);
    v = __fmaf_rn(b2, v, __fmaf_rn(-b2, gg, gg));
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2_sqrt), eps);
    p = __fsub_rn(p, __fdiv_rn(__fmul_rn(step_size, m), denom));
}

__global__ void __launch_bounds__(kAdamThreads)
adam_step_kernel(const AdamArgs a, const float* __restrict__ lr_ptr, float b1, float b2, float eps, float wd) {
    int t = 0;
    while (t + 1 < a.n && (int)blockIdx.x >= a.chunk0[t + 1]) ++t;
    const int64_t off = (int64_t)(blockIdx.x - a.chunk0[t]) * kAdamChunk;
    const int64_t n = a.numel[t] - off < kAdamChunk ? a.numel[t] - off : kAdamChunk;
    const float step = a.step[t][0] + 1.0f;
    const float bc1 = 1.0f - powf(b1, step);
    const float bc2_sqrt = sqrtf(1.0f - powf(b2, step));
    const float step_size = __fdiv_rn(lr_ptr[0], bc1);
    float* __restrict__ p = a.p[t] + off;
    const float* __restrict__ g = a.g[t] + off;
    float* __restrict__ m = a.m[t] + off;
    float* __restrict__ v = a.v[t] + off;
    const bool vec = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                                       reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) % 16 == 0);
    if (vec) {
        for (int64_t i = (int64_t)threadIdx.x * 4; i < n; i += (int64_t)kAdamThreads * 4) {
            float4 pp = *reinterpret_cast<const float4*>(p + i), gg = ldg4(g + i);
            float4 mm = *reinterpret_cast<const float4*>(m + i), vv = *reinterpret_cast<const float4*>(v + i);
            adam_elem(pp.x, gg.x, mm.x, vv.x, wd, b1, b2, step_size, bc2_sqrt, eps);
            adam_elem(pp.y, gg.y, mm.y, vv.y, wd, b1, b2, step_size, bc2_sqrt, eps);
            adam_elem(pp.z, gg.z, mm.z, vv.z, wd, b1, b2, step_size, bc2_sqrt, eps);
            adam_elem(pp.w, gg.w, mm.w, vv.w, wd, b1, b2, step_size, bc2_sqrt, eps);
            stg4(p + i, pp); stg4(m + i, mm); stg4(v + i, vv);
        }
    } else {
        for (int64_t i = threadIdx.x; i < n; i += kAdamThreads) {
            float pp = p[i], mm = m[i], vv = v[i];
            adam_elem(pp, g[i], mm, vv, wd, b1, b2, step_size, bc2_sqrt, eps);
            p[i] = pp; m[i] = mm; v[i] = vv;
        }
    }
}

__global__ void adam_advance_kernel(const AdamArgs a) {
    const int t = threadIdx.x;
    if (t < a.n) a.step[t][0] += 1.0f;
}

}  // namespace ihg

using namespace ihg;

extern "C" int ihg_adam_step(const ihg_adam_tensor* tensors_host, int32_t n_tensors, const float* lr,
                             float beta1, float beta2, float eps, float weight_decay, void* stream) {
    IHG_REQUIRE(n_tensors >= 0 && (n_tensors == 0 || tensors_host), "adam_step: null tensor table");
    IHG_REQUIRE(lr, "adam_step: null lr pointer");
    cudaStream_t st = as_stream(stream);
    for (int base = 0; base < n_tensors; base += kAdamMaxTensors) {
        AdamArgs a;
        a.n = n_tensors - base < kAdamMaxTensors ? n_tensors - base : kAdamMaxTensors;
        int64_t blocks = 0;
        for (int t = 0; t < a.n; ++t) {
            const ihg_adam_tensor& s = tensors_host[base + t];
            IHG_REQUIRE(s.param && s.grad && s.exp_avg && s.exp_avg_sq && s.step && s.numel >= 0,
                        "adam_step: tensor %d has a null pointer", base + t);
            a.p[t] = s.param, a.g[t] = s.grad, a.m[t] = s.exp_avg, a.v[t] = s.exp_avg_sq, a.step[t] = s.step;
            a.numel[t] = s.numel;
            a.chunk0[t] = (int32_t)blocks;
            blocks += ceil_div(s.numel, kAdamChunk);
            IHG_REQUIRE(blocks < (1ll << 31), "adam_step: too many elements");
        }
        a.chunk0[a.n] = (int32_t)blocks;
        if (blocks > 0) {
            adam_step_kernel<<<(unsigned)blocks, kAdamThreads, 0, st>>>(a, lr, beta1, beta2, eps, weight_decay);
            IHG_LAUNCH_CHECK();
        }
        adam_advance_kernel<<<1, 32, 0, st>>>(a);
        IHG_LAUNCH_CHECK();
    }
    return IHG_OK;
}
