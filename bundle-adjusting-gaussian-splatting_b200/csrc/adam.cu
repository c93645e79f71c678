// Fused multi-tensor Adam step for sm_100a (the optimizer step that follows the
// rasterizer's backward in every training iteration).
//
// The reference creates `torch.optim.Adam(l, lr=0.0, eps=1e-15)` over six parameter groups (scene/gaussian_model.py:
// 197-206: xyz, f_dc, f_rest, opacity, scaling, rotation; no weight decay, no amsgrad) and calls `optimizer.step()`
// once per iteration (train.py:420).  PyTorch runs that as ~9 `_foreach_*` passes over every tensor
// (lerp, mul, addcmul, sqrt, div, add, addcdiv ...): 28 algorithmic bytes per element become ~100 bytes of traffic
// and, for the small tensors, tens of launches.  Here the whole step of up to 8 tensors is ONE launch that reads
// param / grad / exp_avg / exp_avg_sq once and writes param / exp_avg / exp_avg_sq once (28 B per element).
//
// Arithmetic mirrors torch's foreach implementation operation by operation, with a rounding to fp32 wherever torch
// has a kernel boundary and the fused multiply-adds nvcc forms inside each of torch's element-wise functors:
//   exp_avg     <- lerp(exp_avg, grad, 1 - beta1)          = fma(w, grad - exp_avg, exp_avg)         (|w| < 0.5)
//                                                          = grad - (grad - exp_avg) * (1 - w)       (otherwise)
//   exp_avg_sq  <- exp_avg_sq * beta2;  exp_avg_sq <- exp_avg_sq + (1 - beta2) * grad * grad
//   denom       <- sqrt(exp_avg_sq) / sqrt(1 - beta2^t) + eps
//   param       <- param + (-lr / (1 - beta1^t)) * (exp_avg / denom)
// The per-tensor scalars (step size, sqrt of the second bias correction) are computed by the caller in double, like
// torch's Python code does, and rounded to fp32 here, like torch's scalar arguments are.
#include "../../include/bags_b200.h"
#include "common.cuh"

namespace bags {

constexpr int ADAM_MAX_TENSORS = 8;
constexpr int ADAM_THREADS = 256;
constexpr int ADAM_ITEMS = 4;                                   // elements per thread (one float4 when aligned)
constexpr int ADAM_BLOCK_ELEMS = ADAM_THREADS * ADAM_ITEMS;

struct AdamBatch {
    float* param[ADAM_MAX_TENSORS];
    const float* grad[ADAM_MAX_TENSORS];
    float* exp_avg[ADAM_MAX_TENSORS];
    float* exp_avg_sq[ADAM_MAX_TENSORS];
    long long n[ADAM_MAX_TENSORS];
    float step_size[ADAM_MAX_TENSORS];        // -lr / bias_correction1
    float bc2_sqrt[ADAM_MAX_TENSORS];         // sqrt(bias_correction2)
    unsigned int block_start[ADAM_MAX_TENSORS + 1];
    int vec4[ADAM_MAX_TENSORS];               // all four pointers 16-byte aligned
    int count;
};

__device__ __forceinline__ void adam_update(float& p, const float g, float& m, float& v, const float w1,
                                            const float beta2, const float w2, const float eps, const float step_size,
                                            const float bc2_sqrt) {
    const float diff = __fsub_rn(g, m);
    m = (fabsf(w1) < 0.5f) ? __fmaf_rn(w1, diff, m) : __fsub_rn(g, __fmul_rn(diff, __fsub_rn(1.0f, w1)));
    v = __fmul_rn(v, beta2);
    v = __fmaf_rn(__fmul_rn(w2, g), g, v);
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2_sqrt), eps);
    p = __fmaf_rn(step_size, __fdiv_rn(m, denom), p);
}

__global__ void __launch_bounds__(ADAM_THREADS)
adam_step_kernel(const AdamBatch b, const float w1, const float beta2, const float w2, const float eps) {
    int t = 0;
#pragma unroll
    for (int k = 1; k < ADAM_MAX_TENSORS; k++)
        if (k < b.count && blockIdx.x >= b.block_start[k]) t = k;
    const long long base = (long long)(blockIdx.x - b.block_start[t]) * ADAM_BLOCK_ELEMS + (long long)threadIdx.x * ADAM_ITEMS;
    const long long n = b.n[t];
    if (base >= n) return;
    float* __restrict__ P = b.param[t];
    const float* __restrict__ G = b.grad[t];
    float* __restrict__ M = b.exp_avg[t];
    float* __restrict__ V = b.exp_avg_sq[t];
    const float step_size = b.step_size[t], bc2_sqrt = b.bc2_sqrt[t];
    if (b.vec4[t] && base + ADAM_ITEMS <= n) {
        float4 p = *reinterpret_cast<const float4*>(P + base);
        const float4 g = *reinterpret_cast<const float4*>(G + base);
        float4 m = *reinterpret_cast<const float4*>(M + base);
        float4 v = *reinterpret_cast<const float4*>(V + base);
        adam_update(p.x, g.x, m.x, v.x, w1, beta2, w2, eps, step_size, bc2_sqrt);
        adam_update(p.y, g.y, m.y, v.y, w1, beta2, w2, eps, step_size, bc2_sqrt);
        adam_update(p.z, g.z, m.z, v.z, w1, beta2, w2, eps, step_size, bc2_sqrt);
        adam_update(p.w, g.w, m.w, v.w, w1, beta2, w2, eps, step_size, bc2_sqrt);
        *reinterpret_cast<float4*>(P + base) = p;
        *reinterpret_cast<float4*>(M + base) = m;
        *reinterpret_cast<float4*>(V + base) = v;
    } else {
#pragma unroll
        for (int k = 0; k < ADAM_ITEMS; k++) {
            const long long i = base + k;
            if (i >= n) break;
            float p = P[i], m = M[i], v = V[i];
            adam_update(p, G[i], m, v, w1, beta2, w2, eps, step_size, bc2_sqrt);
            P[i] = p; M[i] = m; V[i] = v;
        }
    }
}

}  // namespace bags

using namespace bags;

extern "C" int bags_adam_step(int ntensors, const bags_adam_tensor* tensors, double beta1, double beta2, double eps,
                              void* stream) {
    if (ntensors < 0 || (ntensors > 0 && !tensors)) {
        set_error("bags_adam_step: invalid argument");
        return BAGS_E_INVALID;
    }
    for (int i = 0; i < ntensors; i++) {
        const bags_adam_tensor& t = tensors[i];
        if (t.n < 0 || (t.n > 0 && (!t.param || !t.grad || !t.exp_avg || !t.exp_avg_sq))) {
            set_error("bags_adam_step: tensor %d has a null pointer or a negative size", i);
            return BAGS_E_INVALID;
        }
        if (t.n >= (1ll << 40)) {
            set_error("bags_adam_step: tensor %d is too large", i);
            return BAGS_E_OVERFLOW;
        }
    }
    cudaStream_t st = (cudaStream_t)stream;
    // torch passes these as Python floats (double) and its kernels round them to the tensors' opmath type (float)
    const float w1 = (float)(1.0 - beta1), b2 = (float)beta2, w2 = (float)(1.0 - beta2), e = (float)eps;
    int i = 0;
    while (i < ntensors) {
        AdamBatch b;
        b.count = 0;
        unsigned long long blocks = 0;
        b.block_start[0] = 0;
        while (i < ntensors && b.count < ADAM_MAX_TENSORS) {
            const bags_adam_tensor& t = tensors[i++];
            if (t.n == 0) continue;
            const unsigned long long nb = (unsigned long long)((t.n + ADAM_BLOCK_ELEMS - 1) / ADAM_BLOCK_ELEMS);
            if (blocks + nb > 0x7fffffffull) {
                if (b.count == 0) {
                    set_error("bags_adam_step: tensor needs more than 2^31 blocks");
                    return BAGS_E_OVERFLOW;
                }
                i--;
                break;
            }
            const int k = b.count++;
            b.param[k] = t.param; b.grad[k] = t.grad; b.exp_avg[k] = t.exp_avg; b.exp_avg_sq[k] = t.exp_avg_sq;
            b.n[k] = t.n;
            b.step_size[k] = (float)t.step_size;
            b.bc2_sqrt[k] = (float)t.bias_correction2_sqrt;
            b.vec4[k] = ((((size_t)t.param | (size_t)t.grad | (size_t)t.exp_avg | (size_t)t.exp_avg_sq) & 15) == 0) ? 1 : 0;
            blocks += nb;
            b.block_start[k + 1] = (unsigned int)blocks;
        }
        if (b.count == 0) continue;
        adam_step_kernel<<<(unsigned int)blocks, ADAM_THREADS, 0, st>>>(b, w1, b2, w2, e);
        BAGS_LAUNCH_OK("adam_step_kernel");
    }
    return 0;
}
