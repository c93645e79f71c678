// Lens-distortion post-step of the rasterizer output (BASELINE.json config 4) for sm_100a.
//
// The reference applies, right after the rasterizer, three full-canvas PyTorch passes
// (utils/util_distortion.py:271-311 apply_distortion, :58-77 center_crop):
//     flow  = F.interpolate(control-point flow, size=(Hf, Wf), mode='bilinear', align_corners=False)
//     img2  = F.grid_sample(image, flow, mode='bilinear', padding_mode='zeros', align_corners=True)      # (C, Hf, Wf)
//     out   = F.grid_sample(img2, centre-crop grid (Hc, Wc), align_corners=True)                         # center_crop
//     mask  = ~((out[0] == 0) & (out[1] == 0))
// plus the three matching backward passes.  Here the whole chain is ONE gather kernel forward and ONE scatter kernel
// backward: the upsampled flow and img2 are never materialised (img2 alone is C*Hf*Wf*4 bytes = 100 MB on a 4K canvas,
// written once and read twice by the reference).  Only the Hc x Wc crop window is ever evaluated.
//
// Arithmetic restates ATen's kernels operation by operation so that the result is the reference's:
//   * upsample_bilinear2d (UpSampleBilinear2d.cu): src = scale * (dst + 0.5) - 0.5 clamped at 0, scale = in / out,
//     i1 = i0 + (i0 < in - 1), lambda1 = src - i0, lambda0 = 1 - lambda1,
//     val = h0 * (w0 * a + w1 * b) + h1 * (w0 * c + w1 * d);
//   * grid_sampler_2d (GridSampler.cu): ix = ((x + 1) / 2) * (W - 1); weights nw = (x1 - ix)(y1 - iy), ...; taps outside
//     the input contribute nothing; accumulation order nw, ne, sw, se;
//   * center_crop's grid: g = 2 * (start + j) / (size - 1) - 1 as three separately rounded fp32 operations (they are
//     three PyTorch kernels), then the grid_sampler un-normalisation above.  The result is an integer position up to
//     an fp32 rounding error, so the crop is "almost" a copy: the second tap carries a weight of order 1e-5 which the
//     reference does apply -- it is applied here too.
// The backward follows grid_sampler_2d_backward / upsample_bilinear2d_backward (atomic scatter, like ATen's).
#include "../../include/bags_b200.h"
#include "common.cuh"

namespace bags {

constexpr int DIST_MAX_C = 4;

struct UpTaps {
    int i0, i1;
    float l0, l1;
};

// area_pixel_compute_source_index(scale, dst, align_corners=false, cubic=false) + the tap/lambda selection of
// upsample_bilinear2d_out_frame.
__device__ __forceinline__ UpTaps up_taps(const float scale, const int dst, const int in_size) {
    float src = scale * ((float)dst + 0.5f) - 0.5f;
    if (src < 0.f) src = 0.f;
    UpTaps t;
    t.i0 = min((int)src, in_size - 1);
    t.i1 = t.i0 + ((t.i0 < in_size - 1) ? 1 : 0);
    t.l1 = src - (float)t.i0;
    t.l0 = 1.0f - t.l1;
    return t;
}

struct BilTaps {
    int x0, y0;           // north-west tap
    float nw, ne, sw, se;
    float ix, iy;
    bool any;             // false: every tap is outside (also NaN / huge coordinates)
};

__device__ __forceinline__ BilTaps bil_taps(const float gx, const float gy, const int W, const int H) {
    BilTaps t;
    t.ix = ((gx + 1.f) / 2) * (float)(W - 1);
    t.iy = ((gy + 1.f) / 2) * (float)(H - 1);
    t.any = (t.ix > -1.0f) && (t.ix < (float)W) && (t.iy > -1.0f) && (t.iy < (float)H);   // false for NaN
    if (!t.any) {
        t.x0 = t.y0 = -2;
        t.nw = t.ne = t.sw = t.se = 0.f;
        return t;
    }
    const float fx = floorf(t.ix), fy = floorf(t.iy);
    t.x0 = (int)fx;
    t.y0 = (int)fy;
    const float x1 = (float)(t.x0 + 1), y1 = (float)(t.y0 + 1);
    t.nw = (x1 - t.ix) * (y1 - t.iy);
    t.ne = (t.ix - fx) * (y1 - t.iy);
    t.sw = (x1 - t.ix) * (t.iy - fy);
    t.se = (t.ix - fx) * (t.iy - fy);
    return t;
}

// center_crop's normalised grid coordinate of crop index j (utils/util_distortion.py:66-72): linspace gives the
// integer start + j exactly; "2.0 * grid / (size - 1) - 1.0" is three element-wise kernels, i.e. three roundings.
__device__ __forceinline__ float crop_coord(const int start, const int j, const int size) {
    return __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, (float)(start + j)), (float)(size - 1)), 1.0f);
}

struct DistortArgs {
    const float* image;   // [C][H][W]
    const float* flow;    // [Hs][Ws][2] control-point flow (or the full-resolution flow when Hs == Hf, Ws == Wf)
    int C, H, W, Hs, Ws, Hf, Wf, Hc, Wc;
    int crop;             // 1: centre crop Hc x Wc out of the Hf x Wf canvas; 0: Hc == Hf, Wc == Wf, no second sampling
    int mask_mode;        // 0: out[0] == 0 && out[1] == 0 (image branch); 1: out[0] < 1e-5 && out[1] < 1e-5 (apply2gt)
    float scale_h, scale_w;
    int start_y, start_x;
};

// Upsampled flow at canvas pixel (xi, yi).
__device__ __forceinline__ float2 flow_at(const DistortArgs& a, const int xi, const int yi, UpTaps* tx_out = nullptr,
                                          UpTaps* ty_out = nullptr) {
    const UpTaps ty = up_taps(a.scale_h, yi, a.Hs);
    const UpTaps tx = up_taps(a.scale_w, xi, a.Ws);
    const float2* f = reinterpret_cast<const float2*>(a.flow);
    const float2 v00 = __ldg(f + (size_t)ty.i0 * a.Ws + tx.i0), v01 = __ldg(f + (size_t)ty.i0 * a.Ws + tx.i1);
    const float2 v10 = __ldg(f + (size_t)ty.i1 * a.Ws + tx.i0), v11 = __ldg(f + (size_t)ty.i1 * a.Ws + tx.i1);
    float2 r;
    r.x = ty.l0 * (tx.l0 * v00.x + tx.l1 * v01.x) + ty.l1 * (tx.l0 * v10.x + tx.l1 * v11.x);
    r.y = ty.l0 * (tx.l0 * v00.y + tx.l1 * v01.y) + ty.l1 * (tx.l0 * v10.y + tx.l1 * v11.y);
    if (tx_out) *tx_out = tx;
    if (ty_out) *ty_out = ty;
    return r;
}

// img2[c](xi, yi) = grid_sample(image, flow)(xi, yi) for every channel.
__device__ __forceinline__ void canvas_pixel(const DistortArgs& a, const int xi, const int yi, float (&v)[DIST_MAX_C]) {
    const float2 g = flow_at(a, xi, yi);
    const BilTaps t = bil_taps(g.x, g.y, a.W, a.H);
#pragma unroll
    for (int c = 0; c < DIST_MAX_C; c++) v[c] = 0.f;
    if (!t.any) return;
    const bool xl = t.x0 >= 0, xr = t.x0 + 1 < a.W, yt = t.y0 >= 0, yb = t.y0 + 1 < a.H;
    const size_t plane = (size_t)a.H * a.W;
    const float* p = a.image + (ptrdiff_t)t.y0 * a.W + t.x0;
#pragma unroll
    for (int c = 0; c < DIST_MAX_C; c++) {
        if (c < a.C) {
            const float* pc = p + c * plane;
            float acc = 0.f;
            if (yt && xl) acc += __ldg(pc) * t.nw;
            if (yt && xr) acc += __ldg(pc + 1) * t.ne;
            if (yb && xl) acc += __ldg(pc + a.W) * t.sw;
            if (yb && xr) acc += __ldg(pc + a.W + 1) * t.se;
            v[c] = acc;
        }
    }
}

__global__ void __launch_bounds__(256)
distort_fwd_kernel(const DistortArgs a, float* __restrict__ out, float* __restrict__ mask) {
    const int xo = blockIdx.x * 32 + (threadIdx.x & 31);
    const int yo = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (xo >= a.Wc || yo >= a.Hc) return;
    float acc[DIST_MAX_C];
    if (!a.crop) {
        canvas_pixel(a, xo, yo, acc);
    } else {
        const BilTaps t = bil_taps(crop_coord(a.start_x, xo, a.Wf), crop_coord(a.start_y, yo, a.Hf), a.Wf, a.Hf);
#pragma unroll
        for (int c = 0; c < DIST_MAX_C; c++) acc[c] = 0.f;
        if (t.any) {
            const float w[4] = {t.nw, t.ne, t.sw, t.se};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int xi = t.x0 + (k & 1), yi = t.y0 + (k >> 1);
                if (xi < 0 || xi >= a.Wf || yi < 0 || yi >= a.Hf) continue;
                if (w[k] == 0.f) continue;   // 0 * finite adds nothing
                float v[DIST_MAX_C];
                canvas_pixel(a, xi, yi, v);
#pragma unroll
                for (int c = 0; c < DIST_MAX_C; c++) acc[c] += v[c] * w[k];
            }
        }
    }
    const size_t oplane = (size_t)a.Hc * a.Wc;
    const size_t o = (size_t)yo * a.Wc + xo;
#pragma unroll
    for (int c = 0; c < DIST_MAX_C; c++)
        if (c < a.C) out[c * oplane + o] = acc[c];
    if (mask) {
        const float c0 = acc[0], c1 = a.C > 1 ? acc[1] : acc[0];
        const bool hole = a.mask_mode ? (c0 < 0.00001f && c1 < 0.00001f) : (c0 == 0.0f && c1 == 0.0f);
        mask[o] = hole ? 0.f : 1.f;
    }
}

// Full-resolution flow (apply_distortion returns it so that later iterations can reuse it).
__global__ void __launch_bounds__(256)
flow_upsample_kernel(const DistortArgs a, float2* __restrict__ out_flow) {
    const int xi = blockIdx.x * 32 + (threadIdx.x & 31);
    const int yi = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (xi >= a.Wf || yi >= a.Hf) return;
    out_flow[(size_t)yi * a.Wf + xi] = flow_at(a, xi, yi);
}

// Backward.  One thread per CANVAS pixel (xi, yi) of the crop window grown by one pixel: it first gathers
// dL/dimg2(xi, yi) from the <= 3x3 output pixels whose crop taps can touch it (grid_sampler_2d_backward of the crop,
// as a gather: no atomics, no intermediate tensor), then runs grid_sampler_2d_backward of the flow sampling for that
// one pixel: 4 x C atomic adds into dL/dimage, and the flow gradient, which upsample_bilinear2d_backward spreads over
// 4 control points.  A warp is 32 consecutive xi of one row, so control-point columns repeat along the warp: the
// contributions are summed per run of equal columns with a segmented shuffle reduction before they reach memory.
__device__ __forceinline__ float seg_reduce(float v, const int key, const int lane) {
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int k2 = __shfl_down_sync(0xffffffffu, key, off);
        const float t = __shfl_down_sync(0xffffffffu, v, off);
        if (lane + off < 32 && k2 == key) v += t;
    }
    return v;
}

__global__ void __launch_bounds__(256)
distort_bwd_kernel(const DistortArgs a, const float* __restrict__ dL_dout, float* __restrict__ dL_dimage,
                   float* __restrict__ dL_dflow, const int x_lo, const int y_lo, const int x_hi, const int y_hi) {
    const int lane = threadIdx.x & 31;
    const int xi = x_lo + blockIdx.x * 32 + lane;
    const int yi = y_lo + blockIdx.y * 8 + (threadIdx.x >> 5);
    const bool active = xi < x_hi && yi < y_hi;   // the whole warp shares yi
    float G[DIST_MAX_C];
#pragma unroll
    for (int c = 0; c < DIST_MAX_C; c++) G[c] = 0.f;
    const size_t oplane = (size_t)a.Hc * a.Wc;
    bool any = false;
    if (active) {
        if (!a.crop) {
#pragma unroll
            for (int c = 0; c < DIST_MAX_C; c++)
                if (c < a.C) G[c] = __ldg(dL_dout + c * oplane + (size_t)yi * a.Wc + xi);
            any = true;
        } else {
            // output pixels whose north-west crop tap can be xi-1 or xi: (xi - start_x) + {-1, 0, 1}
            for (int dy = -1; dy <= 1; dy++) {
                const int yo = yi - a.start_y + dy;
                if (yo < 0 || yo >= a.Hc) continue;
                const float gy = crop_coord(a.start_y, yo, a.Hf);
                for (int dx = -1; dx <= 1; dx++) {
                    const int xo = xi - a.start_x + dx;
                    if (xo < 0 || xo >= a.Wc) continue;
                    const BilTaps t = bil_taps(crop_coord(a.start_x, xo, a.Wf), gy, a.Wf, a.Hf);
                    if (!t.any) continue;
                    const int kx = xi - t.x0, ky = yi - t.y0;
                    if (kx < 0 || kx > 1 || ky < 0 || ky > 1) continue;
                    const float w = ky ? (kx ? t.se : t.sw) : (kx ? t.ne : t.nw);
                    if (w == 0.f) continue;
                    any = true;
#pragma unroll
                    for (int c = 0; c < DIST_MAX_C; c++)
                        if (c < a.C) G[c] += w * __ldg(dL_dout + c * oplane + (size_t)yo * a.Wc + xo);
                }
            }
        }
    }
    // grid_sampler_2d_backward at (xi, yi).  The control-point taps are computed for every lane (also idle ones) so
    // that the keys of the segmented reduction below are non-decreasing along the whole warp.
    float gfx = 0.f, gfy = 0.f;
    const UpTaps tx = up_taps(a.scale_w, xi, a.Ws), ty = up_taps(a.scale_h, yi, a.Hs);
    if (any) {
        const float2 g = flow_at(a, xi, yi);
        const BilTaps t = bil_taps(g.x, g.y, a.W, a.H);
        if (t.any) {
            const bool xl = t.x0 >= 0, xr = t.x0 + 1 < a.W, yt = t.y0 >= 0, yb = t.y0 + 1 < a.H;
            const size_t plane = (size_t)a.H * a.W;
            const ptrdiff_t o = (ptrdiff_t)t.y0 * a.W + t.x0;
            const float x1 = (float)(t.x0 + 1), y1 = (float)(t.y0 + 1), x0f = (float)t.x0, y0f = (float)t.y0;
            float gix = 0.f, giy = 0.f;
#pragma unroll
            for (int c = 0; c < DIST_MAX_C; c++) {
                if (c >= a.C) continue;
                const float gO = G[c];
                const float* pc = a.image + c * plane + o;
                float* gc = dL_dimage ? dL_dimage + c * plane + o : nullptr;
                if (yt && xl) {
                    const float v = __ldg(pc);
                    if (gc) atomicAdd(gc, t.nw * gO);
                    gix -= v * (y1 - t.iy) * gO;
                    giy -= v * (x1 - t.ix) * gO;
                }
                if (yt && xr) {
                    const float v = __ldg(pc + 1);
                    if (gc) atomicAdd(gc + 1, t.ne * gO);
                    gix += v * (y1 - t.iy) * gO;
                    giy -= v * (t.ix - x0f) * gO;
                }
                if (yb && xl) {
                    const float v = __ldg(pc + a.W);
                    if (gc) atomicAdd(gc + a.W, t.sw * gO);
                    gix -= v * (t.iy - y0f) * gO;
                    giy += v * (x1 - t.ix) * gO;
                }
                if (yb && xr) {
                    const float v = __ldg(pc + a.W + 1);
                    if (gc) atomicAdd(gc + a.W + 1, t.se * gO);
                    gix += v * (t.iy - y0f) * gO;
                    giy += v * (t.ix - x0f) * gO;
                }
            }
            gfx = ((float)(a.W - 1) / 2) * gix;   // grid_sampler_unnormalize_set_grad, align_corners = true
            gfy = ((float)(a.H - 1) / 2) * giy;
        }
    }
    if (!dL_dflow) return;
    // upsample_bilinear2d_backward: control point (ty.i*, tx.i*) += lambda_y * lambda_x * g.  Idle lanes carry zeros.
    const int key = tx.i0;
    const float s0x = seg_reduce(tx.l0 * gfx, key, lane), s0y = seg_reduce(tx.l0 * gfy, key, lane);
    const float s1x = seg_reduce(tx.l1 * gfx, key, lane), s1y = seg_reduce(tx.l1 * gfy, key, lane);
    const int prev = __shfl_up_sync(0xffffffffu, key, 1);
    if (lane == 0 || prev != key) {
        float* r0 = dL_dflow + ((size_t)ty.i0 * a.Ws) * 2;
        float* r1 = dL_dflow + ((size_t)ty.i1 * a.Ws) * 2;
        if (s0x != 0.f || s0y != 0.f) {
            atomicAdd(r0 + 2 * tx.i0, ty.l0 * s0x); atomicAdd(r0 + 2 * tx.i0 + 1, ty.l0 * s0y);
            if (ty.l1 != 0.f) { atomicAdd(r1 + 2 * tx.i0, ty.l1 * s0x); atomicAdd(r1 + 2 * tx.i0 + 1, ty.l1 * s0y); }
        }
        if (s1x != 0.f || s1y != 0.f) {
            atomicAdd(r0 + 2 * tx.i1, ty.l0 * s1x); atomicAdd(r0 + 2 * tx.i1 + 1, ty.l0 * s1y);
            if (ty.l1 != 0.f) { atomicAdd(r1 + 2 * tx.i1, ty.l1 * s1x); atomicAdd(r1 + 2 * tx.i1 + 1, ty.l1 * s1y); }
        }
    }
}

// center_crop alone (utils/util_distortion.py:58-77; used on its own by the evaluation loop): [C,H,W] -> [C,Hc,Wc].
__global__ void __launch_bounds__(256)
crop_fwd_kernel(const float* __restrict__ image, int C, int H, int W, int Hc, int Wc, int start_y, int start_x,
                float* __restrict__ out) {
    const int xo = blockIdx.x * 32 + (threadIdx.x & 31);
    const int yo = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (xo >= Wc || yo >= Hc) return;
    const BilTaps t = bil_taps(crop_coord(start_x, xo, W), crop_coord(start_y, yo, H), W, H);
    const bool xl = t.x0 >= 0, xr = t.x0 + 1 < W, yt = t.y0 >= 0, yb = t.y0 + 1 < H;
    const size_t plane = (size_t)H * W, oplane = (size_t)Hc * Wc;
    for (int c = 0; c < C; c++) {
        float acc = 0.f;
        if (t.any) {
            const float* pc = image + c * plane + (ptrdiff_t)t.y0 * W + t.x0;
            if (yt && xl) acc += __ldg(pc) * t.nw;
            if (yt && xr) acc += __ldg(pc + 1) * t.ne;
            if (yb && xl) acc += __ldg(pc + W) * t.sw;
            if (yb && xr) acc += __ldg(pc + W + 1) * t.se;
        }
        out[c * oplane + (size_t)yo * Wc + xo] = acc;
    }
}

// Its backward as a gather: one thread per INPUT pixel; every input pixel is written (zeros outside the window).
__global__ void __launch_bounds__(256)
crop_bwd_kernel(const float* __restrict__ dL_dout, int C, int H, int W, int Hc, int Wc, int start_y, int start_x,
                float* __restrict__ dL_dimage) {
    const int xi = blockIdx.x * 32 + (threadIdx.x & 31);
    const int yi = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (xi >= W || yi >= H) return;
    float G[DIST_MAX_C];
#pragma unroll
    for (int c = 0; c < DIST_MAX_C; c++) G[c] = 0.f;
    const size_t oplane = (size_t)Hc * Wc;
    for (int dy = -1; dy <= 1; dy++) {
        const int yo = yi - start_y + dy;
        if (yo < 0 || yo >= Hc) continue;
        const float gy = crop_coord(start_y, yo, H);
        for (int dx = -1; dx <= 1; dx++) {
            const int xo = xi - start_x + dx;
            if (xo < 0 || xo >= Wc) continue;
            const BilTaps t = bil_taps(crop_coord(start_x, xo, W), gy, W, H);
            if (!t.any) continue;
            const int kx = xi - t.x0, ky = yi - t.y0;
            if (kx < 0 || kx > 1 || ky < 0 || ky > 1) continue;
            const float w = ky ? (kx ? t.se : t.sw) : (kx ? t.ne : t.nw);
            if (w == 0.f) continue;
#pragma unroll
            for (int c = 0; c < DIST_MAX_C; c++)
                if (c < C) G[c] += w * __ldg(dL_dout + c * oplane + (size_t)yo * Wc + xo);
        }
    }
    const size_t plane = (size_t)H * W;
#pragma unroll
    for (int c = 0; c < DIST_MAX_C; c++)
        if (c < C) dL_dimage[c * plane + (size_t)yi * W + xi] = G[c];
}

static int fill_args(DistortArgs& a, const float* image, int C, int H, int W, const float* flow, int Hs, int Ws, int Hf,
                     int Wf, int Hc, int Wc, int crop, int mask_mode) {
    if (!image || !flow || C < 1 || C > DIST_MAX_C || H < 1 || W < 1 || Hs < 1 || Ws < 1 || Hf < 1 || Wf < 1 ||
        Hc < 1 || Wc < 1) {
        set_error("distort: invalid shape or null input (C must be 1..%d)", DIST_MAX_C);
        return BAGS_E_INVALID;
    }
    if (crop ? (Hc > Hf || Wc > Wf || Hf < 2 || Wf < 2) : (Hc != Hf || Wc != Wf)) {
        set_error("distort: crop %dx%d does not fit the %dx%d canvas", Hc, Wc, Hf, Wf);
        return BAGS_E_INVALID;
    }
    a.image = image; a.flow = flow;
    a.C = C; a.H = H; a.W = W; a.Hs = Hs; a.Ws = Ws; a.Hf = Hf; a.Wf = Wf; a.Hc = Hc; a.Wc = Wc;
    a.crop = crop; a.mask_mode = mask_mode;
    a.scale_h = (float)Hs / (float)Hf;   // area_pixel_compute_scale, align_corners = false, no explicit scale factor
    a.scale_w = (float)Ws / (float)Wf;
    a.start_y = (Hf - Hc) / 2;           // center_crop: (height - target_height) // 2
    a.start_x = (Wf - Wc) / 2;
    return 0;
}

}  // namespace bags

using namespace bags;

extern "C" int bags_distort_forward(const float* image, int C, int H, int W, const float* flow, int Hs, int Ws, int Hf,
                                    int Wf, int Hc, int Wc, int crop, int mask_mode, float* out_image, float* out_mask,
                                    float* out_flow, void* stream) {
    DistortArgs a;
    int rc = fill_args(a, image, C, H, W, flow, Hs, Ws, Hf, Wf, Hc, Wc, crop, mask_mode);
    if (rc) return rc;
    if (!out_image) {
        set_error("bags_distort_forward: null output");
        return BAGS_E_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((Wc + 31) / 32, (Hc + 7) / 8);
    distort_fwd_kernel<<<grid, 256, 0, st>>>(a, out_image, out_mask);
    BAGS_LAUNCH_OK("distort_fwd_kernel");
    if (out_flow) {
        dim3 g2((Wf + 31) / 32, (Hf + 7) / 8);
        flow_upsample_kernel<<<g2, 256, 0, st>>>(a, reinterpret_cast<float2*>(out_flow));
        BAGS_LAUNCH_OK("flow_upsample_kernel");
    }
    return 0;
}

extern "C" int bags_distort_backward(const float* image, int C, int H, int W, const float* flow, int Hs, int Ws, int Hf,
                                     int Wf, int Hc, int Wc, int crop, const float* dL_dout, float* dL_dimage,
                                     float* dL_dflow, void* stream) {
    DistortArgs a;
    int rc = fill_args(a, image, C, H, W, flow, Hs, Ws, Hf, Wf, Hc, Wc, crop, 0);
    if (rc) return rc;
    if (!dL_dout || (!dL_dimage && !dL_dflow)) {
        set_error("bags_distort_backward: null gradient pointer");
        return BAGS_E_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (dL_dimage) BAGS_CUDA_OK(cudaMemsetAsync(dL_dimage, 0, (size_t)C * H * W * sizeof(float), st));
    if (dL_dflow) BAGS_CUDA_OK(cudaMemsetAsync(dL_dflow, 0, (size_t)Hs * Ws * 2 * sizeof(float), st));
    // canvas window that the crop can touch: the crop rectangle grown by one pixel, clipped to the canvas
    int x_lo = 0, y_lo = 0, x_hi = Wf, y_hi = Hf;
    if (crop) {
        x_lo = a.start_x - 1 < 0 ? 0 : a.start_x - 1;
        y_lo = a.start_y - 1 < 0 ? 0 : a.start_y - 1;
        x_hi = a.start_x + Wc + 1 > Wf ? Wf : a.start_x + Wc + 1;
        y_hi = a.start_y + Hc + 1 > Hf ? Hf : a.start_y + Hc + 1;
    }
    dim3 grid((x_hi - x_lo + 31) / 32, (y_hi - y_lo + 7) / 8);
    distort_bwd_kernel<<<grid, 256, 0, st>>>(a, dL_dout, dL_dimage, dL_dflow, x_lo, y_lo, x_hi, y_hi);
    BAGS_LAUNCH_OK("distort_bwd_kernel");
    return 0;
}

extern "C" int bags_center_crop_forward(const float* image, int C, int H, int W, int Hc, int Wc, float* out,
                                        void* stream) {
    if (!image || !out || C < 1 || C > DIST_MAX_C || H < 2 || W < 2 || Hc < 1 || Wc < 1 || Hc > H || Wc > W) {
        set_error("bags_center_crop_forward: invalid argument (C must be 1..%d, crop must fit)", DIST_MAX_C);
        return BAGS_E_INVALID;
    }
    dim3 grid((Wc + 31) / 32, (Hc + 7) / 8);
    crop_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(image, C, H, W, Hc, Wc, (H - Hc) / 2, (W - Wc) / 2, out);
    BAGS_LAUNCH_OK("crop_fwd_kernel");
    return 0;
}

extern "C" int bags_center_crop_backward(const float* dL_dout, int C, int H, int W, int Hc, int Wc, float* dL_dimage,
                                         void* stream) {
    if (!dL_dout || !dL_dimage || C < 1 || C > DIST_MAX_C || H < 2 || W < 2 || Hc < 1 || Wc < 1 || Hc > H || Wc > W) {
        set_error("bags_center_crop_backward: invalid argument");
        return BAGS_E_INVALID;
    }
    dim3 grid((W + 31) / 32, (H + 7) / 8);
    crop_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dL_dout, C, H, W, Hc, Wc, (H - Hc) / 2, (W - Wc) / 2,
                                                           dL_dimage);
    BAGS_LAUNCH_OK("crop_bwd_kernel");
    return 0;
}
