// Front-to-back alpha blending (forward) and its back-to-front gradient (backward) for sm_100a.
//
// Forward restates renderCUDA of 3dgs-pose/cuda_rasterizer/forward.cu:662-798: per 16x16 tile, walk the tile's
// depth-sorted splat list; alpha = min(0.99, o * exp(power)); skip power > 0 and alpha < 1/255; stop a pixel when
// T * (1 - alpha) < 1e-4; accumulate colour, depth (sum d * alpha * T) and alpha (sum alpha * T); write final_T and
// n_contrib for the backward.  Backward restates renderCUDA of backward.cu:925-1129.
//
// Differences from the reference are in staging and reduction only:
//   * one packed 48-byte record per splat (see preprocess.cu) is gathered into shared memory -- colour and depth
//     included, where the reference re-reads colour from global memory per pixel (forward.cu:764);
//   * a warp owns an 8x4 pixel patch and leaves the splat loop as soon as all of its pixels are done
//     (the reference only stops when the whole 256-pixel CTA is done);
//   * the backward only walks list positions below the CTA's largest n_contrib, reduces the 12 per-splat gradient
//     terms across the warp with a transposed butterfly (16 + 3 shuffles instead of 60) and commits them once per
//     (warp, splat) with three 128-bit vector reductions (red.global.add.v4.f32) into a packed 16-float record --
//     instead of 13 scalar global atomics per (pixel, splat) pair (backward.cu:1067-1116).  (Shared-memory fp32
//     atomics are CAS loops on this architecture, so a per-CTA shared accumulation stage was measured and dropped.)
#include "../../include/bags_b200.h"
#include "common.cuh"

namespace bags {

constexpr int BLEND_THREADS = 256;
constexpr int BATCH = 256;

// thread -> pixel inside the 16x16 tile: warp w covers the 8x4 patch at ((w&1)*8, (w>>1)*4)
__device__ __forceinline__ void thread_pixel(int tid, int& lx, int& ly) {
    const int w = tid >> 5, l = tid & 31;
    lx = ((w & 1) << 3) + (l & 7);
    ly = ((w >> 1) << 2) + (l >> 3);
}

// Shared-memory accesses through 32-bit shared-window addresses: one IMAD + LDS.128 per record field instead of
// the generic-pointer conversion sequence the compiler otherwise re-emits inside the hot loop.
__device__ __forceinline__ u32 smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds128(u32 a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ u32 lds32(u32 a) {
    u32 v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

// Conservative screen-space half extents of the region where a splat can reach alpha >= 1/255:
// o * exp(-0.5 d^T Q d) >= 1/255  <=>  d^T Q d <= 2 ln(255 o) =: 2t, whose bounding box has half widths
// sqrt(2t * Qinv_xx) = sqrt(2t * c / det), sqrt(2t * a / det).  Inflated by 0.1% + half a pixel so that fp32 rounding
// of `power` can never make a culled pair pass the reference's alpha test.  o < 1/255 can never contribute
// (alpha <= o), encoded as a negative extent.  Non-finite / indefinite conics yield NaN extents, which the
// NaN-safe test below treats as "hit" (never culled).
__device__ __forceinline__ float2 splat_extent(const float4 con_o) {
    const float t = logf(255.0f * con_o.w);
    if (t < 0.0f) return make_float2(-1.0f, -1.0f);
    const float det = con_o.x * con_o.z - con_o.y * con_o.y;
    const float k = 2.0f * t / det;
    return make_float2(sqrtf(k * con_o.z) * 1.001f + 0.5f, sqrtf(k * con_o.x) * 1.001f + 0.5f);
}

// Does the splat's box {x, y, ex, ey} touch the pixel rectangle [x0,x1] x [y0,y1]?  (NaN-safe towards "yes")
__device__ __forceinline__ bool splat_hits_rect(const float4 box, const float x0, const float x1, const float y0,
                                                const float y1) {
    const float dx = box.x - fminf(fmaxf(box.x, x0), x1);
    const float dy = box.y - fminf(fmaxf(box.y, y0), y1);
    return !(fabsf(dx) > box.z) && !(fabsf(dy) > box.w);
}

// Per-batch shared staging: packed 48-byte records + one {x, y, ex, ey} box per splat for the patch test.
struct BatchSmem {
    float4 rec[BATCH * 3];
    float4 box[BATCH];
};

__global__ void __launch_bounds__(BLEND_THREADS)
blend_fwd_kernel(const uint2* __restrict__ ranges, const u32* __restrict__ point_list, int W, int H,
                 const float4* __restrict__ rec, float* __restrict__ final_T, u32* __restrict__ n_contrib,
                 const float* __restrict__ bg_color, float* __restrict__ out_color, float* __restrict__ out_depth,
                 float* __restrict__ out_alpha) {
    __shared__ __align__(16) BatchSmem sm;
    const u32 s_rec = smem_addr(sm.rec);
    const u32 s_box = smem_addr(sm.box);

    const int tid = threadIdx.x, lane = tid & 31;
    int lx, ly;
    thread_pixel(tid, lx, ly);
    const u32 px = blockIdx.x * TILE_X + lx;
    const u32 py = blockIdx.y * TILE_Y + ly;
    const u32 pix_id = (u32)W * py + px;
    const float2 pixf = {(float)px, (float)py};
    const bool inside = px < (u32)W && py < (u32)H;
    bool done = !inside;

    // pixel rectangle of this warp's 8x4 patch
    const int w = tid >> 5;
    const float wx0 = (float)(blockIdx.x * TILE_X + ((w & 1) << 3)), wx1 = wx0 + 7.0f;
    const float wy0 = (float)(blockIdx.y * TILE_Y + ((w >> 1) << 2)), wy1 = wy0 + 3.0f;

    const uint2 range = ranges[blockIdx.y * gridDim.x + blockIdx.x];
    const int rounds = ((range.y - range.x + BATCH - 1) / BATCH);
    int toDo = range.y - range.x;

    float T = 1.0f;
    u32 last_contributor = 0;
    float C[3] = {0.f, 0.f, 0.f};
    float D = 0.0f;
    float weight = 0.0f;

    for (int i = 0; i < rounds; i++, toDo -= BATCH) {
        const int num_done = __syncthreads_count(done);
        if (num_done == BLEND_THREADS) break;

        const int progress = i * BATCH + tid;
        if (range.x + progress < range.y) {
            const u32 id = point_list[range.x + progress];
            const float4* r = rec + 3 * (size_t)id;
            const float4 q0 = __ldg(r + 0);
            const float4 q1 = __ldg(r + 1);
            sm.rec[3 * tid + 0] = q0;
            sm.rec[3 * tid + 1] = q1;
            sm.rec[3 * tid + 2] = __ldg(r + 2);
            const float2 e = splat_extent(q1);
            sm.box[tid] = make_float4(q0.x, q0.y, e.x, e.y);
        }
        __syncthreads();

        if (__all_sync(0xffffffffu, done)) continue;  // whole warp finished: keep helping with the fetch only

        const int cnt = min(BATCH, toDo);
        // Per-warp culling: 32 splats are box-tested against the warp's patch at once; only the survivors are
        // evaluated per pixel, in list order (bit order of the ballot).
        for (int base = 0; base < cnt; base += 32) {
            const int jt = base + lane;
            bool hit = false;
            if (jt < cnt) hit = splat_hits_rect(lds128(s_box + jt * 16), wx0, wx1, wy0, wy1);
            u32 mask = __ballot_sync(0xffffffffu, hit);
            while (mask) {
                const int j = base + __ffs(mask) - 1;
                mask &= mask - 1;
                if (done) continue;
                const u32 a = s_rec + j * 48;
                const float4 q0 = lds128(a);
                const float4 con_o = lds128(a + 16);
                const float2 d = {q0.x - pixf.x, q0.y - pixf.y};
                const float power = -0.5f * (con_o.x * d.x * d.x + con_o.z * d.y * d.y) - con_o.y * d.x * d.y;
                if (power > 0.0f) continue;
                const float alpha = min(0.99f, con_o.w * exp(power));
                if (alpha < 1.0f / 255.0f) continue;
                const float test_T = T * (1 - alpha);
                if (test_T < 0.0001f) {
                    done = true;
                    continue;
                }
                const float4 col = lds128(a + 32);
                C[0] += col.x * alpha * T;
                C[1] += col.y * alpha * T;
                C[2] += col.z * alpha * T;
                weight += alpha * T;
                D += q0.z * alpha * T;
                T = test_T;
                last_contributor = (u32)(i * BATCH + j + 1);
            }
            if (__all_sync(0xffffffffu, done)) break;
        }
    }

    if (inside) {
        final_T[pix_id] = T;
        n_contrib[pix_id] = last_contributor;
        const size_t plane = (size_t)H * W;
        out_color[0 * plane + pix_id] = C[0] + T * bg_color[0];
        out_color[1 * plane + pix_id] = C[1] + T * bg_color[1];
        out_color[2 * plane + pix_id] = C[2] + T * bg_color[2];
        out_depth[pix_id] = D;
        out_alpha[pix_id] = weight;
    }
}

int launch_blend_fwd(int W, int H, const GeomState& g, const BinningState& b, ImageState& img, const float* bg,
                     float* out_color, float* out_depth, float* out_alpha, cudaStream_t st) {
    dim3 grid(tiles_x(W), tiles_y(H), 1);
    blend_fwd_kernel<<<grid, BLEND_THREADS, 0, st>>>(img.ranges, b.point_list, W, H, g.rec, img.final_T,
                                                    img.n_contrib, bg, out_color, out_depth, out_alpha);
    BAGS_LAUNCH_OK("blend_fwd_kernel");
    return 0;
}

// =================================================================================================================
// Backward
// =================================================================================================================

// Transposed butterfly: every lane holds 16 partial values; on return lane L holds the warp-wide sum of value
// ((L>>4)&1)*8 + ((L>>3)&1)*4 + ((L>>2)&1)*2 + ((L>>1)&1)  (so lane 2k holds value k).  16 shuffles in total.
__device__ __forceinline__ float warp_reduce16(float (&v)[16], int lane) {
    float a8[8];
    {
        const bool up = lane & 16;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const float send = up ? v[k] : v[k + 8];
            const float keep = up ? v[k + 8] : v[k];
            a8[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
    float a4[4];
    {
        const bool up = lane & 8;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float send = up ? a8[k] : a8[k + 4];
            const float keep = up ? a8[k + 4] : a8[k];
            a4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    }
    float a2[2];
    {
        const bool up = lane & 4;
#pragma unroll
        for (int k = 0; k < 2; k++) {
            const float send = up ? a4[k] : a4[k + 2];
            const float keep = up ? a4[k + 2] : a4[k];
            a2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    float a1;
    {
        const bool up = lane & 2;
        const float send = up ? a2[0] : a2[1];
        const float keep = up ? a2[1] : a2[0];
        a1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    a1 += __shfl_xor_sync(0xffffffffu, a1, 1);
    return a1;
}

struct BatchSmemBwd {
    float4 rec[BATCH * 3];
    float4 box[BATCH];
    u32 id[BATCH];
};

__global__ void __launch_bounds__(BLEND_THREADS)
blend_bwd_kernel(const uint2* __restrict__ ranges, const u32* __restrict__ point_list, int W, int H,
                 const float* __restrict__ bg_color, const float4* __restrict__ rec,
                 const float* __restrict__ final_Ts, const u32* __restrict__ n_contrib,
                 const float* __restrict__ dL_dpixels, const float* __restrict__ dL_dpixel_depths,
                 const float* __restrict__ dL_dalphas, float* __restrict__ gacc) {
    __shared__ __align__(16) BatchSmemBwd sm;
    __shared__ u32 s_wmax[BLEND_THREADS / 32];
    const u32 s_rec = smem_addr(sm.rec);
    const u32 s_box = smem_addr(sm.box);
    const u32 s_id = smem_addr(sm.id);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int lx, ly;
    thread_pixel(tid, lx, ly);
    const u32 px = blockIdx.x * TILE_X + lx;
    const u32 py = blockIdx.y * TILE_Y + ly;
    const u32 pix_id = (u32)W * py + px;
    const float2 pixf = {(float)px, (float)py};
    const bool inside = px < (u32)W && py < (u32)H;

    const uint2 range = ranges[blockIdx.y * gridDim.x + blockIdx.x];

    const float T_final = inside ? final_Ts[pix_id] : 0;
    float T = T_final;
    const u32 last_contributor = inside ? n_contrib[pix_id] : 0;

    float accum_rec[3] = {0.f, 0.f, 0.f};
    float dL_dpixel[3] = {0.f, 0.f, 0.f};
    float accum_depth_rec = 0;
    float dL_dpixel_depth = 0;
    float accum_alpha_rec = 0;
    float dL_dweight_alpha = 0;
    if (inside) {
        const size_t plane = (size_t)H * W;
#pragma unroll
        for (int i = 0; i < 3; i++) dL_dpixel[i] = dL_dpixels[i * plane + pix_id];
        dL_dpixel_depth = dL_dpixel_depths[pix_id];
        dL_dweight_alpha = dL_dalphas[pix_id];
    }
    float last_alpha = 0;
    float last_color[3] = {0.f, 0.f, 0.f};
    float last_depth = 0;
    float bg_dot_dpixel = 0;
#pragma unroll
    for (int i = 0; i < 3; i++) bg_dot_dpixel += bg_color[i] * dL_dpixel[i];
    const float neg_Tfinal_bg = -T_final * bg_dot_dpixel;

    const float ddelx_dx = 0.5 * W;
    const float ddely_dy = 0.5 * H;

    // pixel rectangle of this warp's 8x4 patch (per-warp culling, see splat_extent)
    const float wx0 = (float)(blockIdx.x * TILE_X + ((warp & 1) << 3)), wx1 = wx0 + 7.0f;
    const float wy0 = (float)(blockIdx.y * TILE_Y + ((warp >> 1) << 2)), wy1 = wy0 + 3.0f;

    // Only list positions below the largest n_contrib of the CTA (resp. warp) can contribute.
    const u32 warp_max = __reduce_max_sync(0xffffffffu, last_contributor);
    if (lane == 0) s_wmax[warp] = warp_max;
    __syncthreads();
    u32 block_max = 0;
#pragma unroll
    for (int w = 0; w < BLEND_THREADS / 32; w++) block_max = max(block_max, s_wmax[w]);

    // walk positions [0, block_max) back to front in chunks of BATCH; slot j of a chunk is position hi-1-j
    for (int hi = (int)block_max; hi > 0; hi -= BATCH) {
        const int cnt = min(BATCH, hi);
        __syncthreads();  // everyone is done with the previous chunk
        if (tid < cnt) {
            const u32 id = point_list[range.x + (u32)(hi - 1 - tid)];
            const float4* r = rec + 3 * (size_t)id;
            const float4 q0 = __ldg(r + 0);
            const float4 q1 = __ldg(r + 1);
            sm.id[tid] = id;
            sm.rec[3 * tid + 0] = q0;
            sm.rec[3 * tid + 1] = q1;
            sm.rec[3 * tid + 2] = __ldg(r + 2);
            const float2 e = splat_extent(q1);
            sm.box[tid] = make_float4(q0.x, q0.y, e.x, e.y);
        }
        __syncthreads();

        // first slot this warp can contribute to: position hi-1-j < warp_max  <=>  j >= hi-warp_max
        int j0 = hi - (int)warp_max;
        if (j0 < 0) j0 = 0;
        for (int base = j0 & ~31; base < cnt; base += 32) {
            const int jt = base + lane;
            bool hit = false;
            if (jt >= j0 && jt < cnt) hit = splat_hits_rect(lds128(s_box + jt * 16), wx0, wx1, wy0, wy1);
            u32 hmask = __ballot_sync(0xffffffffu, hit);
            while (hmask) {
                const int j = base + __ffs(hmask) - 1;
                hmask &= hmask - 1;
                const u32 pos = (u32)(hi - 1 - j);
                const u32 a = s_rec + j * 48;
                const float4 q0 = lds128(a);
                const float4 con_o = lds128(a + 16);
                const float2 d = {q0.x - pixf.x, q0.y - pixf.y};
                const float power = -0.5f * (con_o.x * d.x * d.x + con_o.z * d.y * d.y) - con_o.y * d.x * d.y;
                const float G = exp(power);
                const float alpha = min(0.99f, con_o.w * G);
                const bool valid = (pos < last_contributor) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
                const u32 vmask = __ballot_sync(0xffffffffu, valid);
                if (vmask == 0) continue;

                float v[16];
#pragma unroll
                for (int k = 0; k < 16; k++) v[k] = 0.f;
                if (valid) {
                    // T_j = T_{j+1} / (1 - alpha); one reciprocal serves both divisions of backward.cu:1051,1090
                    const float rcp = __fdividef(1.0f, 1.f - alpha);
                    T = T * rcp;
                    const float dchannel_dcolor = alpha * T;
                    const float4 col = lds128(a + 32);
                    const float c3[3] = {col.x, col.y, col.z};
                    float dL_dalpha = 0.0f;
#pragma unroll
                    for (int ch = 0; ch < 3; ch++) {
                        const float c = c3[ch];
                        accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                        last_color[ch] = c;
                        const float dL_dchannel = dL_dpixel[ch];
                        dL_dalpha += (c - accum_rec[ch]) * dL_dchannel;
                        v[GA_R + ch] = dchannel_dcolor * dL_dchannel;
                    }
                    const float c_d = q0.z;
                    accum_depth_rec = last_alpha * last_depth + (1.f - last_alpha) * accum_depth_rec;
                    last_depth = c_d;
                    dL_dalpha += (c_d - accum_depth_rec) * dL_dpixel_depth;
                    v[GA_DEPTH] = dchannel_dcolor * dL_dpixel_depth;

                    accum_alpha_rec = last_alpha + (1.f - last_alpha) * accum_alpha_rec;
                    dL_dalpha += (1 - accum_alpha_rec) * dL_dweight_alpha;

                    dL_dalpha *= T;
                    last_alpha = alpha;
                    dL_dalpha += neg_Tfinal_bg * rcp;

                    const float dL_dG = con_o.w * dL_dalpha;
                    const float gdx = G * d.x;
                    const float gdy = G * d.y;
                    const float dG_ddelx = -gdx * con_o.x - gdy * con_o.y;
                    const float dG_ddely = -gdy * con_o.z - gdx * con_o.y;
                    const float mx = dL_dG * dG_ddelx * ddelx_dx;
                    const float my = dL_dG * dG_ddely * ddely_dy;
                    v[GA_MX] = mx;
                    v[GA_MY] = my;
                    v[GA_AMX] = fabsf(mx);
                    v[GA_AMY] = fabsf(my);
                    v[GA_CA] = -0.5f * gdx * d.x * dL_dG;
                    v[GA_CB] = -0.5f * gdx * d.y * dL_dG;
                    v[GA_CC] = -0.5f * gdy * d.y * dL_dG;
                    v[GA_OP] = G * dL_dalpha;
                }
                // warp sum of the 12 terms; lane 2k holds term k.  Lanes 0, 8, 16 gather their four neighbours and
                // commit one 128-bit vector reduction each; lane 24 adds the covariance diagnostic.
                const float r0 = warp_reduce16(v, lane);
                const float r1 = __shfl_down_sync(0xffffffffu, r0, 2);
                const float r2 = __shfl_down_sync(0xffffffffu, r0, 4);
                const float r3 = __shfl_down_sync(0xffffffffu, r0, 6);
                if ((lane & 7) == 0) {
                    float* dst = gacc + (size_t)lds32(s_id + j * 4) * GACC_STRIDE;
                    if (lane < 24) atomicAdd(reinterpret_cast<float4*>(dst + (lane >> 1)), make_float4(r0, r1, r2, r3));
                    else atomicAdd(dst + GA_COV, (float)__popc(vmask) * (con_o.x + con_o.z));
                }
            }
        }
    }
}

int launch_blend_bwd(int W, int H, const GeomState& g, const BinningState& b, const ImageState& img, const float* bg,
                     const float* dL_dpix, const float* dL_dpix_depth, const float* dL_dalpha, float* gacc,
                     cudaStream_t st) {
    dim3 grid(tiles_x(W), tiles_y(H), 1);
    blend_bwd_kernel<<<grid, BLEND_THREADS, 0, st>>>(img.ranges, b.point_list, W, H, bg, g.rec, img.final_T,
                                                    img.n_contrib, dL_dpix, dL_dpix_depth, dL_dalpha, gacc);
    BAGS_LAUNCH_OK("blend_bwd_kernel");
    return 0;
}

}  // namespace bags
