// Photometric loss of the training step for sm_100a: L1 and SSIM of the rendered image
// against the ground truth, forward and backward, fused.
//
// The reference (utils/loss_utils.py:18-19 l1_loss, :49-76 ssim/_ssim; combined at train.py:325 as
// (1 - lambda) * L1 + lambda * (1 - SSIM)) runs five grouped 11x11 convolutions (mu1, mu2, E[x^2], E[y^2], E[xy], zero
// padding 5) and about fifteen element-wise kernels forward, and autograd's mirror of all of them backward -- some
// twenty full-image round trips on the tensor the blend kernel has just written.  Here:
//
//   loss_fwd_kernel : one pass over a 32x32 tile (+5 halo) per CTA.  The 11x11 window of the reference is the outer
//                     product of a 1-D Gaussian (sigma 1.5), so the five windowed sums are evaluated separably in
//                     shared memory (11 + 11 taps instead of 121).  Per pixel it forms the SSIM map value, adds it and
//                     |x - y| to per-CTA partial sums (summed in a fixed order by loss_finalize_kernel: the result is
//                     deterministic), and stores the three partial derivatives of the map that the backward needs
//                     (d/dmu1, d/dE[x^2], d/dE[xy]);
//   loss_bwd_kernel : dL/dx(q) = c_ssim * sum_p w(p - q) [dmu1(p) + 2 x(q) dE11(p) + y(q) dE12(p)] + c_l1 * sign(x - y):
//                     the same separable stencil over the three stored maps.
//
// Algorithmic traffic per pixel and channel: forward 8 B read + 12 B written, backward 20 B read + 4 B written.
#include "../../include/bags_b200.h"
#include "common.cuh"

#include <math.h>

namespace bags {

constexpr int LW = 11, LR = 5;                 // window size and radius (loss_utils.py:49: window_size = 11)
constexpr int LT = 32;                         // output tile edge
constexpr int LH = LT + 2 * LR;                // staged tile edge (42)
constexpr int LTHREADS = 256;

struct LossWindow { float g[LW]; };

// loss_utils.py:36-38: gauss = Tensor([exp(-(x - 5)^2 / (2 * 1.5^2))]) / sum, evaluated in double by Python and
// rounded to fp32 by torch.Tensor, then divided in fp32 by gauss.sum().  torch's CPU sum of the 11 values is a
// multi-accumulator reduction whose result equals the correctly rounded sum; a left-to-right fp32 loop lands one ulp
// lower, which scales every weight by 1 + 6e-8 and -- because sigma12 = E[xy] - mu1 mu2 is a difference of nearly
// equal numbers -- biased the mean SSIM of smooth 1080p images by 1.2e-5.  The sum is therefore formed in double and
// rounded once.
static LossWindow make_window() {
    LossWindow w;
    double s = 0.0;
    for (int x = 0; x < LW; x++) {
        w.g[x] = (float)exp(-(double)((x - LW / 2) * (x - LW / 2)) / (2.0 * 1.5 * 1.5));
        s += (double)w.g[x];
    }
    const float sf = (float)s;
    for (int x = 0; x < LW; x++) w.g[x] = w.g[x] / sf;
    return w;
}

__device__ __forceinline__ float block_sum_256(float v, float* s_red /*[8]*/) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < 8) t = s_red[threadIdx.x];
    if (threadIdx.x < 32) {
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    __syncthreads();
    return t;   // valid in thread 0
}

// grid: (tiles_x, tiles_y, planes)
__global__ void __launch_bounds__(LTHREADS)
loss_fwd_kernel(const float* __restrict__ img1, const float* __restrict__ img2, int H, int W, const LossWindow win,
                float* __restrict__ d_mu1, float* __restrict__ d_e11, float* __restrict__ d_e12,
                float2* __restrict__ partial, const int want_ssim) {
    __shared__ float s_x[LH][LH + 1], s_y[LH][LH + 1];
    __shared__ float s_h[5][LH][LT + 1];      // horizontal sums of x, y, x^2, y^2, xy for the 42 staged rows
    __shared__ float s_red[8];
    const int tid = threadIdx.x;
    const size_t plane = (size_t)blockIdx.z * H * W;
    const int x0 = blockIdx.x * LT - LR, y0 = blockIdx.y * LT - LR;

    for (int i = tid; i < LH * LH; i += LTHREADS) {
        const int r = i / LH, c = i - r * LH;
        const int yy = y0 + r, xx = x0 + c;
        float a = 0.f, b = 0.f;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
            a = __ldg(img1 + plane + (size_t)yy * W + xx);
            b = __ldg(img2 + plane + (size_t)yy * W + xx);
        }
        s_x[r][c] = a;
        s_y[r][c] = b;
    }
    __syncthreads();

    float l1 = 0.f, ss = 0.f;
    if (want_ssim) {
        for (int i = tid; i < LH * LT; i += LTHREADS) {
            const int r = i / LT, c = i - r * LT;
            float sx = 0.f, sy = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
            for (int k = 0; k < LW; k++) {
                const float a = s_x[r][c + k], b = s_y[r][c + k], g = win.g[k];
                sx += g * a;
                sy += g * b;
                sxx += g * (a * a);
                syy += g * (b * b);
                sxy += g * (a * b);
            }
            s_h[0][r][c] = sx; s_h[1][r][c] = sy; s_h[2][r][c] = sxx; s_h[3][r][c] = syy; s_h[4][r][c] = sxy;
        }
        __syncthreads();
    }

    const int c = tid & 31;
#pragma unroll
    for (int rr = 0; rr < LT / 8; rr++) {
        const int r = (tid >> 5) + rr * 8;
        const int yy = y0 + LR + r, xx = x0 + LR + c;
        if (yy >= H || xx >= W) continue;
        const float a = s_x[r + LR][c + LR], b = s_y[r + LR][c + LR];
        l1 += fabsf(a - b);
        if (!want_ssim) continue;
        float mu1 = 0.f, mu2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
        for (int k = 0; k < LW; k++) {
            const float g = win.g[k];
            mu1 += g * s_h[0][r + k][c];
            mu2 += g * s_h[1][r + k][c];
            e11 += g * s_h[2][r + k][c];
            e22 += g * s_h[3][r + k][c];
            e12 += g * s_h[4][r + k][c];
        }
        // loss_utils.py:60-71
        const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu1_mu2 = mu1 * mu2;
        const float sigma1_sq = e11 - mu1_sq, sigma2_sq = e22 - mu2_sq, sigma12 = e12 - mu1_mu2;
        const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
        const float A = 2.f * mu1_mu2 + C1, B = 2.f * sigma12 + C2;
        const float Cc = mu1_sq + mu2_sq + C1, D = sigma1_sq + sigma2_sq + C2;
        const float CD = Cc * D;
        const float S = (A * B) / CD;
        ss += S;
        if (d_mu1) {
            // partial derivatives of S with E[x^2], E[xy] held fixed for d/dmu1 (sigma1_sq, sigma12 depend on mu1)
            const float inv = 1.0f / CD;
            const size_t o = plane + (size_t)yy * W + xx;
            d_mu1[o] = 2.f * mu2 * (B - A) * inv - 2.f * mu1 * S * (D - Cc) * inv;
            d_e11[o] = -S / D;
            d_e12[o] = 2.f * A * inv;
        }
    }
    const float tl1 = block_sum_256(l1, s_red);
    const float tss = block_sum_256(ss, s_red);
    if (tid == 0)
        partial[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = make_float2(tl1, tss);
}

// One CTA per plane: fixed-order double sum of the per-CTA partials -> sums[plane] = {sum |x-y|, sum ssim_map}.
__global__ void __launch_bounds__(256)
loss_finalize_kernel(const float2* __restrict__ partial, int nblocks, double* __restrict__ sums) {
    __shared__ double s_a[256], s_b[256];
    const float2* p = partial + (size_t)blockIdx.x * nblocks;
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 256) {
        a += (double)p[i].x;
        b += (double)p[i].y;
    }
    s_a[threadIdx.x] = a;
    s_b[threadIdx.x] = b;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            s_a[threadIdx.x] += s_a[threadIdx.x + o];
            s_b[threadIdx.x] += s_b[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        sums[2 * blockIdx.x] = s_a[0];
        sums[2 * blockIdx.x + 1] = s_b[0];
    }
}

// coef[plane] = {dL/d(sum |x-y|), dL/d(sum ssim_map)} (device memory: the upstream gradient never visits the host).
__global__ void __launch_bounds__(LTHREADS)
loss_bwd_kernel(const float* __restrict__ img1, const float* __restrict__ img2, int H, int W, const LossWindow win,
                const float* __restrict__ d_mu1, const float* __restrict__ d_e11, const float* __restrict__ d_e12,
                const float2* __restrict__ coef, float* __restrict__ dL_dimg1) {
    __shared__ float s_m[3][LH][LH + 1];
    __shared__ float s_h[3][LH][LT + 1];
    const int tid = threadIdx.x;
    const size_t plane = (size_t)blockIdx.z * H * W;
    const int x0 = blockIdx.x * LT - LR, y0 = blockIdx.y * LT - LR;
    const float2 cf = coef[blockIdx.z];
    const bool with_ssim = d_mu1 != nullptr;

    if (with_ssim) {
        for (int i = tid; i < LH * LH; i += LTHREADS) {
            const int r = i / LH, c = i - r * LH;
            const int yy = y0 + r, xx = x0 + c;
            float a = 0.f, b = 0.f, d = 0.f;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
                const size_t o = plane + (size_t)yy * W + xx;
                a = __ldg(d_mu1 + o);
                b = __ldg(d_e11 + o);
                d = __ldg(d_e12 + o);
            }
            s_m[0][r][c] = a; s_m[1][r][c] = b; s_m[2][r][c] = d;
        }
        __syncthreads();
        for (int i = tid; i < LH * LT; i += LTHREADS) {
            const int r = i / LT, c = i - r * LT;
            float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
            for (int k = 0; k < LW; k++) {
                const float g = win.g[k];
                a += g * s_m[0][r][c + k];
                b += g * s_m[1][r][c + k];
                d += g * s_m[2][r][c + k];
            }
            s_h[0][r][c] = a; s_h[1][r][c] = b; s_h[2][r][c] = d;
        }
        __syncthreads();
    }
    const int c = tid & 31;
#pragma unroll
    for (int rr = 0; rr < LT / 8; rr++) {
        const int r = (tid >> 5) + rr * 8;
        const int yy = y0 + LR + r, xx = x0 + LR + c;
        if (yy >= H || xx >= W) continue;
        const size_t o = plane + (size_t)yy * W + xx;
        const float x = __ldg(img1 + o), y = __ldg(img2 + o);
        float g = 0.f;
        if (with_ssim) {
            float a = 0.f, b = 0.f, d = 0.f;
#pragma unroll
            for (int k = 0; k < LW; k++) {
                const float w = win.g[k];
                a += w * s_h[0][r + k][c];
                b += w * s_h[1][r + k][c];
                d += w * s_h[2][r + k][c];
            }
            g = cf.y * (a + 2.f * x * b + y * d);
        }
        const float df = x - y;
        g += cf.x * (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f));
        dL_dimg1[o] = g;
    }
}

static inline int loss_blocks(int H, int W) { return ((W + LT - 1) / LT) * ((H + LT - 1) / LT); }

}  // namespace bags

using namespace bags;

extern "C" int bags_loss_window(float* taps11) {
    if (!taps11) {
        set_error("bags_loss_window: null output");
        return BAGS_E_INVALID;
    }
    const LossWindow w = make_window();
    for (int x = 0; x < LW; x++) taps11[x] = w.g[x];
    return 0;
}

extern "C" size_t bags_loss_scratch_bytes(int planes, int H, int W) {
    if (planes < 0 || H < 0 || W < 0) return 0;
    return align_up((size_t)planes * loss_blocks(H, W) * sizeof(float2)) + 256;
}

extern "C" int bags_loss_forward(const float* img1, const float* img2, int planes, int H, int W, int want_ssim,
                                 float* d_mu1, float* d_e11, float* d_e12, char* scratch, double* sums, void* stream) {
    if (!img1 || !img2 || !scratch || !sums || planes < 1 || H < 1 || W < 1 || planes > 65535) {
        set_error("bags_loss_forward: invalid argument");
        return BAGS_E_INVALID;
    }
    if ((d_mu1 != nullptr) != (d_e11 != nullptr) || (d_mu1 != nullptr) != (d_e12 != nullptr) ||
        (d_mu1 && !want_ssim)) {
        set_error("bags_loss_forward: pass all three derivative maps (with want_ssim) or none");
        return BAGS_E_INVALID;
    }
    static const LossWindow win = make_window();
    cudaStream_t st = (cudaStream_t)stream;
    float2* partial = reinterpret_cast<float2*>(align_up((size_t)scratch));
    dim3 grid((W + LT - 1) / LT, (H + LT - 1) / LT, planes);
    loss_fwd_kernel<<<grid, LTHREADS, 0, st>>>(img1, img2, H, W, win, d_mu1, d_e11, d_e12, partial, want_ssim);
    BAGS_LAUNCH_OK("loss_fwd_kernel");
    loss_finalize_kernel<<<planes, 256, 0, st>>>(partial, loss_blocks(H, W), sums);
    BAGS_LAUNCH_OK("loss_finalize_kernel");
    return 0;
}

extern "C" int bags_loss_backward(const float* img1, const float* img2, int planes, int H, int W, const float* d_mu1,
                                  const float* d_e11, const float* d_e12, const float* coef, float* dL_dimg1,
                                  void* stream) {
    if (!img1 || !img2 || !coef || !dL_dimg1 || planes < 1 || H < 1 || W < 1 || planes > 65535) {
        set_error("bags_loss_backward: invalid argument");
        return BAGS_E_INVALID;
    }
    if ((d_mu1 != nullptr) != (d_e11 != nullptr) || (d_mu1 != nullptr) != (d_e12 != nullptr)) {
        set_error("bags_loss_backward: pass all three derivative maps or none");
        return BAGS_E_INVALID;
    }
    static const LossWindow win = make_window();
    dim3 grid((W + LT - 1) / LT, (H + LT - 1) / LT, planes);
    loss_bwd_kernel<<<grid, LTHREADS, 0, (cudaStream_t)stream>>>(img1, img2, H, W, win, d_mu1, d_e11, d_e12,
                                                                reinterpret_cast<const float2*>(coef), dL_dimg1);
    BAGS_LAUNCH_OK("loss_bwd_kernel");
    return 0;
}

// =================================================================================================================
// L1 against an 8-bit ground truth, forward and gradient in one pass.
//
// The data set stores 8-bit images; the reference converts them to fp32 once (utils/general_utils.py PILtoTorch:
// `/ 255.0`) and evaluates l1_loss = mean |render - gt| (utils/loss_utils.py:18-19) with autograd's mirror
// (sign(render - gt) / numel) behind it: about ten element-wise passes over the image per iteration.  Here one kernel
// reads the render (fp32) and the ground truth as stored (uint8, scaled by 1/255 in registers like PILtoTorch's
// division), writes d(mean |.|)/d(render) = sign / numel and a per-CTA partial sum; a second tiny kernel adds the
// partials in a fixed order in double and writes the fp32 mean.  Deterministic.
// =================================================================================================================
namespace bags {

constexpr int L1U8_THREADS = 256;
constexpr int L1U8_ITEMS = 16;     // elements per thread: four 16-byte fp32 loads + one 16-byte u8 load

__global__ void __launch_bounds__(L1U8_THREADS)
l1_u8_fwd_kernel(const float* __restrict__ img, const uint8_t* __restrict__ gt, long long n, float inv_n,
                 float* __restrict__ dL_dimg, float* __restrict__ partial, int vec) {
    __shared__ float s_red[8];
    const long long base = ((long long)blockIdx.x * L1U8_THREADS + threadIdx.x) * L1U8_ITEMS;
    float acc = 0.f;
    if (vec && base + L1U8_ITEMS <= n) {
        const uint4 g16 = *reinterpret_cast<const uint4*>(gt + base);
        const unsigned int gw[4] = {g16.x, g16.y, g16.z, g16.w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const float4 x = *reinterpret_cast<const float4*>(img + base + 4 * q);
            const float xs[4] = {x.x, x.y, x.z, x.w};
            float gs[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float y = (float)((gw[q] >> (8 * k)) & 0xffu) / 255.0f;
                const float d = xs[k] - y;
                acc += fabsf(d);
                gs[k] = d > 0.f ? inv_n : (d < 0.f ? -inv_n : 0.f);
            }
            *reinterpret_cast<float4*>(dL_dimg + base + 4 * q) = make_float4(gs[0], gs[1], gs[2], gs[3]);
        }
    } else {
        for (int k = 0; k < L1U8_ITEMS; k++) {
            const long long i = base + k;
            if (i >= n) break;
            const float d = img[i] - (float)gt[i] / 255.0f;
            acc += fabsf(d);
            dL_dimg[i] = d > 0.f ? inv_n : (d < 0.f ? -inv_n : 0.f);
        }
    }
    const float t = block_sum_256(acc, s_red);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

__global__ void __launch_bounds__(256)
l1_u8_finalize_kernel(const float* __restrict__ partial, int nblocks, double inv_n, float* __restrict__ loss) {
    __shared__ double s_a[256];
    double a = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 256) a += (double)partial[i];
    s_a[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s_a[threadIdx.x] += s_a[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss[0] = (float)(s_a[0] * inv_n);
}

static inline long long l1_u8_blocks(long long n) {
    const long long per = (long long)L1U8_THREADS * L1U8_ITEMS;
    return (n + per - 1) / per;
}

}  // namespace bags

extern "C" size_t bags_l1_u8_scratch_bytes(long long n) {
    if (n < 0) return 0;
    return align_up((size_t)l1_u8_blocks(n) * sizeof(float)) + 256;
}

extern "C" int bags_l1_u8_forward(const float* img, const uint8_t* gt, long long n, float* dL_dimg, float* loss,
                                  char* scratch, void* stream) {
    if (!img || !gt || !dL_dimg || !loss || !scratch || n < 1 || l1_u8_blocks(n) > 0x7fffffffll) {
        set_error("bags_l1_u8_forward: invalid argument");
        return BAGS_E_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    float* partial = reinterpret_cast<float*>(align_up((size_t)scratch));
    const int nb = (int)l1_u8_blocks(n);
    const int vec = ((((size_t)img | (size_t)gt | (size_t)dL_dimg) & 15) == 0) ? 1 : 0;
    l1_u8_fwd_kernel<<<nb, L1U8_THREADS, 0, st>>>(img, gt, n, (float)(1.0 / (double)n), dL_dimg, partial, vec);
    BAGS_LAUNCH_OK("l1_u8_fwd_kernel");
    l1_u8_finalize_kernel<<<1, 256, 0, st>>>(partial, nb, 1.0 / (double)n, loss);
    BAGS_LAUNCH_OK("l1_u8_finalize_kernel");
    return 0;
}
