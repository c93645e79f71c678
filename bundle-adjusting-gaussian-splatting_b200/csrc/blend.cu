// Front-to-back alpha blending (forward) and its back-to-front gradient (backward) for sm_100a.
//
// Forward restates renderCUDA of 3dgs-pose/cuda_rasterizer/forward.cu:662-798: per 16x16 tile, walk the tile's
// depth-sorted splat list; alpha = min(0.99, o * exp(power)); skip power > 0 and alpha < 1/255; stop a pixel when
// T * (1 - alpha) < 1e-4; accumulate colour, depth (sum d * alpha * T) and alpha (sum alpha * T); write final_T and
// n_contrib for the backward.  Backward restates renderCUDA of backward.cu:925-1129.
//
// Differences from the reference are in staging, scheduling and reduction only:
//   * one packed 48-byte record per splat (see preprocess.cu) is gathered into shared memory -- colour and depth
//     included, where the reference re-reads colour from global memory per pixel (forward.cu:764);
//   * culling below tile granularity.  The tile is cut into 4x2-pixel cells; while a batch is staged, the loading
//     thread turns the splat's alpha >= 1/255 bounding box into a 32-bit cell mask.  The tile is split into eight 8x4
//     patches; a warp owns PPT of them (one pixel per lane and patch, PPT = 1, 2 or 4 pixels per thread), tests 32
//     candidates at a time against its cells with one ballot and evaluates, per patch, only the splats whose box
//     reaches that patch.  List order per pixel is untouched, so n_contrib, final_T and every threshold decision are
//     those of the reference;
//   * the backward only walks list positions below the largest n_contrib of the CTA / warp / patch.  The 13
//     per-splat gradient terms of a thread's PPT pixels are added in registers, then reduced across the warp with
//     ONE transposed butterfly per (warp, splat) (13 shuffles for 12 slots of 32 * PPT pixels) and committed with one
//     predicated red.global.add.f32 (13 lanes, 13 consecutive words of a packed 16-float record) -- instead of 13
//     scalar global atomics per (pixel, splat) pair (backward.cu:1067-1116).  (Shared-memory fp32 atomics are CAS
//     loops on this architecture, so a per-CTA shared accumulation stage was measured and dropped.)
//   * instruction count per contributing (pixel, splat) pair (round 2; the kernel is FP32-issue bound):
//       - the five back-to-front "colour behind" recurrences (backward.cu:1060,1069,1077) are replaced by ONE scalar
//         recurrence over sum_ch c_ch dL_dchannel (see the kernel);
//       - the per-splat constants (opacity, 0.5 W, 0.5 H, -0.5, conic.x + conic.z) are factored out of the pixel
//         loop: the record holds raw moments (sum q u, sum |q u|, sum q dx^2, ..., the contributing-pixel count) and
//         preprocess_bwd.cu applies the constants once per Gaussian;
//     same mathematical result, different fp32 rounding: covered by the 1e-4 relative-L2 gradient bar.
#include "../../include/bags_b200.h"
#include "common.cuh"

namespace bags {

constexpr int BATCH = 256;
constexpr u32 FULL = 0xffffffffu;

// pixels per thread of the two blend kernels (1, 2 or 4); see bags_set_tuning
static int g_fwd_ppt = 2;
static int g_bwd_ppt = 2;
static int g_cull = 1;   // 0: every splat of the tile list is evaluated for every patch (debugging / tests)
static int g_staging = 0;   // 1: cp.async double-buffered batch staging (see stage_issue); 0: ldg -> sts, one buffer
static int g_bwd_reduce = 1;   // 1 (default): per-splat warp reduction of the backward through shared memory; 0: shuffle butterfly

// Tile decomposition.  Cell grid: 4 columns of 4 pixels x 8 rows of 2 pixels, bit = row * 4 + column.  Patch (x, y),
// x in 0..1, y in 0..3, is the 8x4 pixel block at (8x, 4y) = cells {2x, 2x+1} x {2y, 2y+1} = bits 0x33 << (8y + 2x).
// A CTA has 8 / PPT warps; warp w owns the PPT vertically stacked patches (w & 1, (w >> 1) * PPT + k), k < PPT;
// lane l is pixel (l & 7, l >> 3) of each of them.
template <int PPT>
struct WarpPixels {
    int x, ybase;
    __device__ __forceinline__ WarpPixels(int w) : x(w & 1), ybase((w >> 1) * PPT) {}
    __device__ __forceinline__ int cell_shift(int k) const { return ((ybase + k) << 3) + (x << 1); }
    __device__ __forceinline__ u32 cells() const {
        u32 m = 0;
#pragma unroll
        for (int k = 0; k < PPT; k++) m |= 0x33u << cell_shift(k);
        return m;
    }
    __device__ __forceinline__ int lx(int lane) const { return (x << 3) + (lane & 7); }
    __device__ __forceinline__ int ly(int lane, int k) const { return ((ybase + k) << 2) + (lane >> 3); }
};

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

// power = -0.5 (a dx^2 + c dy^2) - b dx dy with exactly the operation order nvcc gives the reference's expression
// (forward.cu:733, backward.cu:1041; checked against the reference cubin: c*dy*dy and b*dx*dy are rounded products,
// a*dx*dx is fused into the sum, the final -0.5 scaling is fused with the cross term).  Threshold decisions
// (alpha < 1/255, T < 1e-4) flip on the last bit, so the order is pinned with intrinsics instead of left to the
// contraction choices of whatever surrounds the expression.  adx = a*dx and bdx = b*dx are shared by the pixels of a
// thread (same column).
__device__ __forceinline__ float splat_power(const float dx, const float adx, const float bdx, const float dy,
                                             const float c) {
    const float t = __fmul_rn(dy, __fmul_rn(dy, c));
    const float s = __fmaf_rn(dx, adx, t);
    const float u = __fmul_rn(dy, bdx);
    return __fmaf_rn(s, -0.5f, -u);
}

// Cells of the tile at (tx0, ty0) that a splat can reach with alpha >= 1/255.
// o * exp(-0.5 d^T Q d) >= 1/255  <=>  d^T Q d <= 2 ln(255 o) =: 2t, whose bounding box has half widths
// sqrt(2t * Qinv_xx) = sqrt(2t * c / det), sqrt(2t * a / det).  Inflated by 0.1% + half a pixel so that fp32 rounding
// of `power` can never make a culled pair pass the reference's alpha test.  o < 1/255 can never contribute
// (alpha <= o): empty mask.  Non-finite / indefinite conics yield NaN extents, which the clamps below turn into
// "every cell" (never culled).
__device__ __forceinline__ u32 splat_cellmask(const float4 q0, const float4 con_o, const float tx0, const float ty0) {
    const float t = logf(255.0f * con_o.w);
    if (t < 0.0f) return 0u;
    const float det = con_o.x * con_o.z - con_o.y * con_o.y;
    const float k = 2.0f * t / det;
    const float ex = sqrtf(k * con_o.z) * 1.001f + 0.5f;
    const float ey = sqrtf(k * con_o.x) * 1.001f + 0.5f;
    // pixel centres are integers: cell column i holds x in [4i, 4i+3], cell row j holds y in [2j, 2j+1] (tile-local)
    float lo_x = q0.x - ex - tx0, hi_x = q0.x + ex - tx0;
    float lo_y = q0.y - ey - ty0, hi_y = q0.y + ey - ty0;
    lo_x = fminf(fmaxf(lo_x, -8.0f), 32.0f);   // NaN -> -8
    hi_x = fmaxf(fminf(hi_x, 32.0f), -8.0f);   // NaN -> 32
    lo_y = fminf(fmaxf(lo_y, -8.0f), 32.0f);
    hi_y = fmaxf(fminf(hi_y, 32.0f), -8.0f);
    const int i0 = max(0, (int)ceilf((lo_x - 3.0f) * 0.25f)), i1 = min(3, (int)floorf(hi_x * 0.25f));
    const int j0 = max(0, (int)ceilf((lo_y - 1.0f) * 0.5f)), j1 = min(7, (int)floorf(hi_y * 0.5f));
    if (i0 > i1 || j0 > j1) return 0u;
    const u32 cm = (2u << i1) - (1u << i0);
    u32 x = (2u << j1) - (1u << j0);           // row bits r -> bit 4r
    x = (x | (x << 12)) & 0x000F000Fu;
    x = (x | (x << 6)) & 0x03030303u;
    x = (x | (x << 3)) & 0x11111111u;
    return x * cm;
}


// Stage one batch of the tile's list (positions first .. first+cnt-1 in steps of `dir`) into shared memory.
template <int NT, bool WITH_ID>
__device__ __forceinline__ void stage_batch(const u32* __restrict__ point_list, const float4* __restrict__ rec,
                                            const u32 first, const int dir, const int cnt, const float tx0,
                                            const float ty0, float4* sm_rec, u32* sm_cell, u32* sm_id,
                                            const bool cull) {
    for (int s = threadIdx.x; s < cnt; s += NT) {
        const u32 id = point_list[first + (u32)(dir * s)];
        const float4* r = rec + 3 * (size_t)id;
        const float4 q0 = __ldg(r + 0);
        const float4 q1 = __ldg(r + 1);
        if (WITH_ID) sm_id[s] = id;
        sm_rec[3 * s + 0] = q0;
        sm_rec[3 * s + 1] = q1;
        sm_rec[3 * s + 2] = __ldg(r + 2);
        sm_cell[s] = cull ? splat_cellmask(q0, q1, tx0, ty0) : 0xffffffffu;
    }
}

// ---- asynchronous staging (BAGS_TUNE_BLEND_STAGING = 1) --------------------------------------------------------------
// The batch's records go from global to shared memory with cp.async (LDGSTS: no register round trip) into one of two
// buffers, one batch ahead of the one being blended; the cell masks are computed from shared memory once the copies
// of the batch have landed.
__device__ __forceinline__ void cp_async16(const u32 smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NT, bool WITH_ID>
__device__ __forceinline__ void stage_issue(const u32* __restrict__ point_list, const float4* __restrict__ rec,
                                            const u32 first, const int dir, const int cnt, const u32 s_rec,
                                            u32* sm_id) {
    for (int s = threadIdx.x; s < cnt; s += NT) {
        const u32 id = point_list[first + (u32)(dir * s)];
        const float4* r = rec + 3 * (size_t)id;
        if (WITH_ID) sm_id[s] = id;
        cp_async16(s_rec + s * 48, r);
        cp_async16(s_rec + s * 48 + 16, r + 1);
        cp_async16(s_rec + s * 48 + 32, r + 2);
    }
}
template <int NT>
__device__ __forceinline__ void stage_finish(const int cnt, const float tx0, const float ty0, const u32 s_rec,
                                             u32* sm_cell, const bool cull) {
    for (int s = threadIdx.x; s < cnt; s += NT)
        sm_cell[s] = cull ? splat_cellmask(lds128(s_rec + s * 48), lds128(s_rec + s * 48 + 16), tx0, ty0)
                          : 0xffffffffu;
}

template <int PPT, bool ASYNC>
__global__ void __launch_bounds__(256 / PPT)
blend_fwd_kernel(const uint2* __restrict__ ranges, const u32* __restrict__ point_list, int W, int H,
                 const float4* __restrict__ rec, float* __restrict__ final_T, u32* __restrict__ n_contrib,
                 const float* __restrict__ bg_color, float* __restrict__ out_color, float* __restrict__ out_depth,
                 float* __restrict__ out_alpha, const bool cull, const ViewStride vs,
                 const u32* __restrict__ order, u32* __restrict__ bwd_cost, const u32 grid_x) {
    constexpr int NT = 256 / PPT;
    constexpr int NBUF = ASYNC ? 2 : 1;
    if (blockIdx.z != 0) {   // batched launch: blockIdx.z = view, the images of the views are stacked
        const unsigned v = blockIdx.z;
        const size_t pl = (size_t)H * W * v;
        order = at_view(order, vs.img, v);
        bwd_cost = at_view(bwd_cost, vs.img, v);
        ranges = at_view(ranges, vs.img, v);
        final_T = at_view(final_T, vs.img, v);
        n_contrib = at_view(n_contrib, vs.img, v);
        point_list = at_view(point_list, vs.bin, v);
        rec = at_view(rec, vs.geom, v);
        out_color += 3 * pl;
        out_depth += pl;
        out_alpha += pl;
    }
    __shared__ __align__(16) float4 sm_rec[NBUF][BATCH * 3];
    __shared__ u32 sm_cell[NBUF][BATCH];
    const u32 s_rec0 = smem_addr(sm_rec);
    const u32 s_cell0 = smem_addr(sm_cell);

    const int tid = threadIdx.x, lane = tid & 31;
    const WarpPixels<PPT> wp(tid >> 5);
    // CTA i works on the i-th longest tile list (see tile_order_kernel)
    const u32 tile_id = order[blockIdx.x];
    const u32 tile_y = tile_id / grid_x, tile_x = tile_id - tile_y * grid_x;
    const u32 px = tile_x * TILE_X + wp.lx(lane);
    const float pxf = (float)px;
    const float tx0 = (float)(tile_x * TILE_X), ty0 = (float)(tile_y * TILE_Y);

    u32 py[PPT];
    float pyf[PPT];
    bool inside[PPT], done[PPT];
    float T[PPT], C0[PPT], C1[PPT], C2[PPT], D[PPT], Wt[PPT];
    u32 last_contributor[PPT];
#pragma unroll
    for (int k = 0; k < PPT; k++) {
        py[k] = tile_y * TILE_Y + wp.ly(lane, k);
        pyf[k] = (float)py[k];
        inside[k] = px < (u32)W && py[k] < (u32)H;
        done[k] = !inside[k];
        T[k] = 1.0f;
        C0[k] = C1[k] = C2[k] = D[k] = Wt[k] = 0.0f;
        last_contributor[k] = 0;
    }

    const uint2 range = ranges[tile_id];
    const int rounds = ((range.y - range.x + BATCH - 1) / BATCH);
    int toDo = range.y - range.x;

    if (ASYNC && rounds > 0) {
        stage_issue<NT, false>(point_list, rec, range.x, 1, min(BATCH, toDo), s_rec0, nullptr);
        cp_async_commit();
    }
    for (int i = 0; i < rounds; i++, toDo -= BATCH) {
        bool all_done = true;
#pragma unroll
        for (int k = 0; k < PPT; k++) all_done = all_done && done[k];
        const int num_done = __syncthreads_count(all_done);   // also: everyone has left batch i-1's buffer
        if (num_done == NT) break;

        const int cnt = min(BATCH, toDo);
        const int buf = ASYNC ? (i & 1) : 0;
        const u32 s_rec = s_rec0 + buf * (BATCH * 48);
        const u32 s_cell = s_cell0 + buf * (BATCH * 4);
        if (ASYNC) {
            if (i + 1 < rounds)
                stage_issue<NT, false>(point_list, rec, range.x + (u32)((i + 1) * BATCH), 1, min(BATCH, toDo - BATCH),
                                       s_rec0 + (buf ^ 1) * (BATCH * 48), nullptr);
            cp_async_commit();                                // (possibly empty) group of batch i+1
            cp_async_wait<1>();                               // this thread's copies of batch i have landed
            stage_finish<NT>(cnt, tx0, ty0, s_rec, sm_cell[buf], cull);
        } else {
            stage_batch<NT, false>(point_list, rec, range.x + (u32)(i * BATCH), 1, cnt, tx0, ty0, sm_rec[0], sm_cell[0],
                                   nullptr, cull);
        }
        __syncthreads();

        // cells of the patches that still have unfinished pixels
        u32 live = 0;
#pragma unroll
        for (int k = 0; k < PPT; k++)
            if (__ballot_sync(FULL, done[k]) != FULL) live |= 0x33u << wp.cell_shift(k);
        if (live == 0) continue;  // whole warp finished: keep helping with the fetch only

        // 32 candidates are matched against the warp's live cells at once; only the survivors are evaluated, in
        // list order (bit order of the ballot), and only for the patches their box reaches.
        for (int base = 0; base < cnt; base += 32) {
            const int jt = base + lane;
            u32 h = 0;
            if (jt < cnt) h = lds32(s_cell + jt * 4) & live;
            u32 hmask = __ballot_sync(FULL, h != 0u);
            while (hmask) {
                const int j = base + __ffs(hmask) - 1;
                hmask &= hmask - 1;
                const u32 cm = lds32(s_cell + j * 4) & live;
                const u32 a = s_rec + j * 48;
                const float4 q0 = lds128(a);
                const float4 con_o = lds128(a + 16);
                const float dx = q0.x - pxf;
                const float adx = __fmul_rn(dx, con_o.x), bdx = __fmul_rn(dx, con_o.y);
#pragma unroll
                for (int k = 0; k < PPT; k++) {
                    if (((cm >> wp.cell_shift(k)) & 0x33u) == 0u) continue;  // warp-uniform
                    if (done[k]) continue;
                    const float power = splat_power(dx, adx, bdx, q0.y - pyf[k], con_o.z);
                    if (power > 0.0f) continue;
                    const float alpha = min(0.99f, con_o.w * exp(power));
                    if (alpha < 1.0f / 255.0f) continue;
                    const float test_T = T[k] * (1 - alpha);
                    if (test_T < 0.0001f) {
                        done[k] = true;
                        continue;
                    }
                    const float4 col = lds128(a + 32);
                    C0[k] += col.x * alpha * T[k];
                    C1[k] += col.y * alpha * T[k];
                    C2[k] += col.z * alpha * T[k];
                    Wt[k] += alpha * T[k];
                    D[k] += q0.z * alpha * T[k];
                    T[k] = test_T;
                    last_contributor[k] = (u32)(i * BATCH + j + 1);
                }
            }
            live = 0;
#pragma unroll
            for (int k = 0; k < PPT; k++)
                if (__ballot_sync(FULL, done[k]) != FULL) live |= 0x33u << wp.cell_shift(k);
            if (live == 0) break;
        }
    }

    if (ASYNC) cp_async_wait<0>();   // nothing may be in flight into shared memory when the CTA retires
    // the tile's cost in the backward: the list positions it will walk = the largest n_contrib of its pixels
    {
        __shared__ u32 s_cmax[NT / 32];
        u32 m = 0;
#pragma unroll
        for (int k = 0; k < PPT; k++) m = max(m, inside[k] ? last_contributor[k] : 0u);
        m = __reduce_max_sync(FULL, m);
        __syncthreads();             // (also: everyone has left the staging buffers)
        if (lane == 0) s_cmax[tid >> 5] = m;
        __syncthreads();
        if (tid == 0) {
#pragma unroll
            for (int w = 1; w < NT / 32; w++) m = max(m, s_cmax[w]);
            bwd_cost[tile_id] = m;
        }
    }
    const size_t plane = (size_t)H * W;
    const float bg0 = bg_color[0], bg1 = bg_color[1], bg2 = bg_color[2];
#pragma unroll
    for (int k = 0; k < PPT; k++) {
        if (!inside[k]) continue;
        const u32 pix_id = (u32)W * py[k] + px;
        final_T[pix_id] = T[k];
        n_contrib[pix_id] = last_contributor[k];
        out_color[0 * plane + pix_id] = C0[k] + T[k] * bg0;
        out_color[1 * plane + pix_id] = C1[k] + T[k] * bg1;
        out_color[2 * plane + pix_id] = C2[k] + T[k] * bg2;
        out_depth[pix_id] = D[k];
        out_alpha[pix_id] = Wt[k];
    }
}

int launch_blend_fwd(int W, int H, const GeomState& g, const BinningState& b, ImageState& img, const float* bg,
                     float* out_color, float* out_depth, float* out_alpha, cudaStream_t st, int nviews,
                     const ViewStride& vs) {
    const int ntiles = tiles_x(W) * tiles_y(H);
    int rc = launch_tile_order(ntiles, img.ranges, nullptr, img.fwd_order, st, nviews, vs);
    if (rc) return rc;
    dim3 grid(ntiles, 1, nviews > 1 ? nviews : 1);
#define BAGS_FWD(PPT, ASYNC)                                                                                     \
    blend_fwd_kernel<PPT, ASYNC><<<grid, 256 / PPT, 0, st>>>(img.ranges, b.point_list, W, H, g.rec, img.final_T, \
                                                            img.n_contrib, bg, out_color, out_depth, out_alpha, \
                                                            g_cull != 0, vs, img.fwd_order, img.bwd_cost,       \
                                                            (u32)tiles_x(W))
    if (g_staging) {
        if (g_fwd_ppt == 4) BAGS_FWD(4, true);
        else if (g_fwd_ppt == 2) BAGS_FWD(2, true);
        else BAGS_FWD(1, true);
    } else {
        if (g_fwd_ppt == 4) BAGS_FWD(4, false);
        else if (g_fwd_ppt == 2) BAGS_FWD(2, false);
        else BAGS_FWD(1, false);
    }
#undef BAGS_FWD
    BAGS_LAUNCH_OK("blend_fwd_kernel");
    return 0;
}

// =================================================================================================================
// Backward
// =================================================================================================================

// One level of the transposed butterfly: N partial values per lane -> N/2, exchanging with lane ^ OFF.  Lanes with
// bit OFF set keep the upper half of the slots.
template <int N, int OFF>
__device__ __forceinline__ void bfly(const float (&in)[N], float (&out)[N / 2], const int lane) {
    const bool up = lane & OFF;
#pragma unroll
    for (int k = 0; k < N / 2; k++) {
        const float send = up ? in[k] : in[k + N / 2];
        const float keep = up ? in[k + N / 2] : in[k];
        out[k] = keep + __shfl_xor_sync(FULL, send, OFF);
    }
}

// Warp sum of the 12 gradient slots of `v`, added into the splat's 16-float record `dst`, plus the number of
// contributing pixels `cnt` (warp-uniform, from the validity ballots) into slot GA_COV.
// Transposed butterfly 12 -> 6 -> 3 -> {2,1} -> 1 -> 1: 13 shuffles (a plain butterfly per slot would be 60).  Lane
// (b16 b8 b4 b2 b1) ends up with slot 6*b16 + 3*b8 + (b4 ? 2 : b2); the 12 lanes with b1 == 0 that hold distinct slots
// and one spare lane (the count) commit with ONE predicated red.global.add.f32: 13 consecutive words of one 64-byte
// record.
__device__ __forceinline__ void warp_reduce_commit12(const float (&v)[12], const float cnt, const int lane,
                                                     float* dst) {
    float a6[6], a3[3];
    bfly<12, 16>(v, a6, lane);
    bfly<6, 8>(a6, a3, lane);
    const bool up4 = lane & 4, up2 = lane & 2;
    // 3 -> lower lanes keep {0, 1}, upper lanes keep {2}
    const float r0 = __shfl_xor_sync(FULL, up4 ? a3[0] : a3[2], 4);
    const float r1 = __shfl_xor_sync(FULL, up4 ? a3[1] : 0.0f, 4);
    const float b0 = (up4 ? a3[2] : a3[0]) + r0;
    const float b1 = a3[1] + r1;                      // meaningful in the lower lanes only
    // lower lanes: {0, 1} -> one each; upper lanes: plain add of their single slot
    const float send = up4 ? b0 : (up2 ? b0 : b1);
    const float keep = up4 ? b0 : (up2 ? b1 : b0);
    float c = keep + __shfl_xor_sync(FULL, send, 2);
    c += __shfl_xor_sync(FULL, c, 1);
    int slot = ((lane >> 4) & 1) * 6 + ((lane >> 3) & 1) * 3 + (up4 ? 2 : (up2 ? 1 : 0));
    bool commit = !(lane & 1) && !(up4 && up2);
    if (lane == 6) {                                  // (b4, b2) = (1, 1): a duplicate of lane 4's slot -> carries the count
        slot = GA_COV;
        c = cnt;
        commit = true;
    }
    // `dst` is global memory: a plain RED instead of the generic-address atomic (which carries a shared / local window
    // check and two CAS loops along)
    if (commit) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + slot), "f"(c) : "memory");
}

// exp() of the backward.  The forward needs the reference's exact expf (its alpha decides thresholds and colours bit for
// bit); the backward only needs (1) the same contribute / skip decision as the forward and (2) alpha to gradient
// accuracy.  ex2.approx(x log2 e) is within 4e-7 relative on the range that matters (power >= -5.6): the decision is
// re-taken with the exact expression only when alpha lands within 1e-6 of the 1/255 threshold.
__device__ __forceinline__ float exp_bwd(const float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
}
__device__ __forceinline__ float rcp_fast(const float x) {   // x in [0.01, 1]: no denormal handling needed
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// The same reduction through shared memory (BAGS_TUNE_BLEND_BWD_REDUCE = 1): every lane stores its 12 partial sums as
// one 48-byte row (3 STS.128; rows of lanes 16..31 shifted by one row so that the two halves fall on different banks),
// lane L < 24 then adds slot L % 12 over the 16 rows of half L / 12 (16 conflict-free LDS + a 4-level add tree), one
// shuffle joins the halves, and lanes 0..12 commit as above.  About 45 instructions per (warp, splat) instead of 68.
constexpr int RED_ROWS = 33;                      // 32 lanes + the padding row between the halves
constexpr int RED_WARP_BYTES = RED_ROWS * 48;
__device__ __forceinline__ void warp_reduce_commit12_smem(const float (&v)[12], const float cnt, const int lane,
                                                          float* dst, const u32 s_red /* this warp's buffer */) {
    __syncwarp();                                 // the previous splat's rows have been read by everyone
    const u32 row = s_red + (u32)(lane + (lane >> 4)) * 48u;
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(row), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(row + 16u), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(row + 32u), "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]) : "memory");
    __syncwarp();
    const int half = lane >= 12 ? 1 : 0;          // lanes 24..31 read along harmlessly (slots 0..7 of half 1 again)
    const int slot = lane - 12 * half - (lane >= 24 ? 12 : 0);
    const u32 col = s_red + (u32)(half * 17) * 48u + (u32)slot * 4u;
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; i++)      // ("memory": the loads must stay behind the warp barrier above)
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x[i]) : "r"(col + (u32)i * 48u) : "memory");
#pragma unroll
    for (int w = 8; w > 0; w >>= 1)
#pragma unroll
        for (int i = 0; i < w; i++) x[i] += x[i + w];
    float c = x[0] + __shfl_down_sync(FULL, x[0], 12);
    if (lane == 12) c = cnt;                      // GA_COV = 12: the contributing-pixel count rides in the same RED
    if (lane <= 12) asm volatile("red.global.add.f32 [%0], %1;" ::"l"(dst + lane), "f"(c) : "memory");
}

#ifdef BAGS_STATS
__device__ unsigned long long g_stats[64];
#endif

// (held to 64 registers -- possible without spills, 8 instead of 7 resident CTAs per SM -- the kernel was measured SLOWER:
// 0.774 vs 0.741 ms per view; the register allocation is left to ptxas: 70)
template <int PPT, bool ASYNC, bool REDSM>
__global__ void __launch_bounds__(256 / PPT)
blend_bwd_kernel(const uint2* __restrict__ ranges, const u32* __restrict__ point_list, int W, int H,
                 const float* __restrict__ bg_color, const float4* __restrict__ rec,
                 const float* __restrict__ final_Ts, const u32* __restrict__ n_contrib,
                 const float* __restrict__ dL_dpixels, const float* __restrict__ dL_dpixel_depths,
                 const float* __restrict__ dL_dalphas, float* __restrict__ gacc, const bool cull, const ViewStride vs,
                 const size_t gacc_stride, const u32* __restrict__ order, const u32 grid_x) {
    constexpr int NT = 256 / PPT;
    constexpr int NW = NT / 32;
    if (blockIdx.z != 0) {   // batched launch: blockIdx.z = view
        const unsigned v = blockIdx.z;
        const size_t pl = (size_t)H * W * v;
        order = at_view(order, vs.img, v);
        ranges = at_view(ranges, vs.img, v);
        final_Ts = at_view(final_Ts, vs.img, v);
        n_contrib = at_view(n_contrib, vs.img, v);
        point_list = at_view(point_list, vs.bin, v);
        rec = at_view(rec, vs.geom, v);
        dL_dpixels += 3 * pl;
        if (dL_dpixel_depths) dL_dpixel_depths += pl;
        if (dL_dalphas) dL_dalphas += pl;
        gacc = at_view(gacc, gacc_stride, v);
    }
    constexpr int NBUF = ASYNC ? 2 : 1;
    __shared__ __align__(16) float4 sm_rec[NBUF][BATCH * 3];
    __shared__ u32 sm_cell[NBUF][BATCH];
    __shared__ u32 sm_id[NBUF][BATCH];
    __shared__ u32 s_wmax[NW];
    __shared__ __align__(16) unsigned char sm_red[REDSM ? NW * RED_WARP_BYTES : 16];
    const u32 s_rec0 = smem_addr(sm_rec);
    const u32 s_cell0 = smem_addr(sm_cell);
    const u32 s_id0 = smem_addr(sm_id);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const WarpPixels<PPT> wp(warp);
    const u32 tile_id = order[blockIdx.x];             // CTA i works on the tile with the i-th largest n_contrib
    const u32 tile_y = tile_id / grid_x, tile_x = tile_id - tile_y * grid_x;
    const u32 px = tile_x * TILE_X + wp.lx(lane);
    const float pxf = (float)px;
    const float tx0 = (float)(tile_x * TILE_X), ty0 = (float)(tile_y * TILE_Y);
    const uint2 range = ranges[tile_id];
    const size_t plane = (size_t)H * W;
    const float bg0 = bg_color[0], bg1 = bg_color[1], bg2 = bg_color[2];

    // Per-pixel state (registers; every loop over k is fully unrolled).  The pixel's upstream gradients are
    // loop-invariant: with 4 pixels per thread they live in shared memory (two conflict-free loads per evaluated
    // pixel) to keep the kernel under 96 registers; with 1 or 2 they stay in registers.
    constexpr bool G_IN_SMEM = PPT >= 4;
    __shared__ __align__(16) float4 sm_g[G_IN_SMEM ? 256 : 1];   // {dL_dR, dL_dG, dL_dB, dL_ddepth}
    __shared__ __align__(8) float2 sm_w[G_IN_SMEM ? 256 : 1];    // {dL_dweight, -T_final * (bg . dL_dcolour)}
    const u32 s_g = smem_addr(sm_g), s_w = smem_addr(sm_w);
    const float pyf0 = (float)(tile_y * TILE_Y + wp.ly(lane, 0));   // row of pixel k is pyf0 + 4k (exact)
    float T[PPT], Dacc[PPT];
    float g0[G_IN_SMEM ? 1 : PPT], g1[G_IN_SMEM ? 1 : PPT], g2[G_IN_SMEM ? 1 : PPT], gd[G_IN_SMEM ? 1 : PPT],
        gw[G_IN_SMEM ? 1 : PPT], nTbg[G_IN_SMEM ? 1 : PPT];
    u32 last_contributor[PPT], pmax[PPT];
    u32 warp_max = 0;
#pragma unroll
    for (int k = 0; k < PPT; k++) {
        const u32 py = tile_y * TILE_Y + wp.ly(lane, k);
        const bool inside = px < (u32)W && py < (u32)H;
        const u32 pix_id = (u32)W * py + px;
        const float T_final = inside ? final_Ts[pix_id] : 0;
        T[k] = T_final;
        last_contributor[k] = inside ? n_contrib[pix_id] : 0;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        float w = 0.f;
        if (inside) {
            g.x = dL_dpixels[0 * plane + pix_id];
            g.y = dL_dpixels[1 * plane + pix_id];
            g.z = dL_dpixels[2 * plane + pix_id];
            g.w = dL_dpixel_depths ? dL_dpixel_depths[pix_id] : 0.f;   // NULL: no gradient w.r.t. the depth image
            w = dL_dalphas ? dL_dalphas[pix_id] : 0.f;                  // NULL: none w.r.t. the weight image
        }
        float bg_dot_dpixel = 0;
        bg_dot_dpixel += bg0 * g.x;
        bg_dot_dpixel += bg1 * g.y;
        bg_dot_dpixel += bg2 * g.z;
        const float ntb = -T_final * bg_dot_dpixel;
        if (G_IN_SMEM) {
            sm_g[k * NT + tid] = g;
            sm_w[k * NT + tid] = make_float2(w, ntb);
        } else {
            g0[k] = g.x; g1[k] = g.y; g2[k] = g.z; gd[k] = g.w; gw[k] = w; nTbg[k] = ntb;
        }
        Dacc[k] = 0.f;
        // Only list positions below the largest n_contrib of the patch / warp / CTA can contribute.
        pmax[k] = __reduce_max_sync(FULL, last_contributor[k]);
        warp_max = max(warp_max, pmax[k]);
    }
    const u32 warp_cells = wp.cells();

    if (lane == 0) s_wmax[warp] = warp_max;
    __syncthreads();
    u32 block_max = 0;
#pragma unroll
    for (int w = 0; w < NW; w++) block_max = max(block_max, s_wmax[w]);

    // walk positions [0, block_max) back to front in chunks of BATCH; slot j of a chunk is position hi-1-j
    if (ASYNC && block_max > 0) {
        stage_issue<NT, true>(point_list, rec, range.x + (u32)(block_max - 1), -1, min(BATCH, (int)block_max), s_rec0,
                              sm_id[0]);
        cp_async_commit();
    }
    int it = 0;
    for (int hi = (int)block_max; hi > 0; hi -= BATCH, it++) {
        const int cnt = min(BATCH, hi);
        const int buf = ASYNC ? (it & 1) : 0;
        const u32 s_rec = s_rec0 + buf * (BATCH * 48);
        const u32 s_cell = s_cell0 + buf * (BATCH * 4);
        const u32 s_id = s_id0 + buf * (BATCH * 4);
        __syncthreads();  // everyone is done with the previous chunk
        if (ASYNC) {
            if (hi > BATCH)
                stage_issue<NT, true>(point_list, rec, range.x + (u32)(hi - BATCH - 1), -1, min(BATCH, hi - BATCH),
                                      s_rec0 + (buf ^ 1) * (BATCH * 48), sm_id[buf ^ 1]);
            cp_async_commit();
            cp_async_wait<1>();
            stage_finish<NT>(cnt, tx0, ty0, s_rec, sm_cell[buf], cull);
        } else {
            stage_batch<NT, true>(point_list, rec, range.x + (u32)(hi - 1), -1, cnt, tx0, ty0, sm_rec[0], sm_cell[0],
                                  sm_id[0], cull);
        }
        __syncthreads();

        // first slot this warp can contribute to: position hi-1-j < warp_max  <=>  j >= hi-warp_max
        int j0 = hi - (int)warp_max;
        if (j0 < 0) j0 = 0;
        for (int base = j0 & ~31; base < cnt; base += 32) {
            const int jt = base + lane;
            u32 h = 0;
            if (jt >= j0 && jt < cnt) h = lds32(s_cell + jt * 4) & warp_cells;
            u32 hmask = __ballot_sync(FULL, h != 0u);
            while (hmask) {
                const int j = base + __ffs(hmask) - 1;
                hmask &= hmask - 1;
                const u32 pos = (u32)(hi - 1 - j);
                const u32 cm = lds32(s_cell + j * 4);
                const u32 a = s_rec + j * 48;
                const float4 q0 = lds128(a);
                const float4 con_o = lds128(a + 16);
                const float dx = q0.x - pxf;
                const float adx = __fmul_rn(dx, con_o.x), bdx = __fmul_rn(dx, con_o.y);

                float v[12];
#pragma unroll
                for (int s = 0; s < 12; s++) v[s] = 0.f;
                u32 nvalid = 0;                      // contributing pixels of this warp (warp-uniform)
#pragma unroll
                for (int k = 0; k < PPT; k++) {
                    // warp-uniform: does the splat's box reach this patch, and is the position still live there?
                    if (((cm >> wp.cell_shift(k)) & 0x33u) == 0u || pos >= pmax[k]) continue;
                    const float dy = q0.y - (pyf0 + (float)(4 * k));
                    const float power = splat_power(dx, adx, bdx, dy, con_o.z);
                    float Gv = exp_bwd(power);
                    float alpha = min(0.99f, con_o.w * Gv);
                    if (fabsf(alpha - 1.0f / 255.0f) < 1e-6f) {      // (rare) borderline: the forward's own expression
                        Gv = exp(power);
                        alpha = min(0.99f, con_o.w * Gv);
                    }
                    const bool valid = (pos < last_contributor[k]) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
                    nvalid += __popc(__ballot_sync(FULL, valid));
                    if (!valid) continue;
                    float G0, G1, G2, GD, GW, NTBG;
                    if (G_IN_SMEM) {
                        const float4 g = lds128(s_g + (k * NT + tid) * 16);
                        float2 w;
                        asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(w.x), "=f"(w.y) : "r"(s_w + (k * NT + tid) * 8));
                        G0 = g.x; G1 = g.y; G2 = g.z; GD = g.w; GW = w.x; NTBG = w.y;
                    } else {
                        G0 = g0[k]; G1 = g1[k]; G2 = g2[k]; GD = gd[k]; GW = gw[k]; NTBG = nTbg[k];
                    }
                    // T_j = T_{j+1} / (1 - alpha); one (approximate, 1-ulp) reciprocal serves both divisions of
                    // backward.cu:1051,1090 -- see DESIGN.md, tolerance notes
                    const float rcp = rcp_fast(1.f - alpha);
                    T[k] = T[k] * rcp;
                    const float dchannel_dcolor = alpha * T[k];
                    const float4 col = lds128(a + 32);
                    v[GA_R] = fmaf(dchannel_dcolor, G0, v[GA_R]);
                    v[GA_G] = fmaf(dchannel_dcolor, G1, v[GA_G]);
                    v[GA_B] = fmaf(dchannel_dcolor, G2, v[GA_B]);
                    v[GA_DEPTH] = fmaf(dchannel_dcolor, GD, v[GA_DEPTH]);
                    // dL/dalpha.  The reference keeps five "colour behind this splat" recurrences (accum_rec per
                    // channel, depth and weight; backward.cu:1060-1079) and sums (c - accum_rec) * dL_dchannel.  With
                    // cg = sum_ch c_ch * dL_dchannel (depth and the weight's constant 1 included) the same quantity is
                    //   T_j * cg_j - D_j / (1 - alpha_j),   D_j = sum_{i behind j} alpha_i T_i cg_i
                    // -- one scalar recurrence instead of five.
                    const float cg = fmaf(col.x, G0, fmaf(col.y, G1, fmaf(col.z, G2, fmaf(q0.z, GD, GW))));
                    const float dL_dalpha = fmaf(T[k], cg, rcp * (NTBG - Dacc[k]));
                    Dacc[k] = fmaf(dchannel_dcolor, cg, Dacc[k]);
                    // Raw moments; the per-splat constants (opacity, 0.5 W, 0.5 H, -0.5) are applied once per
                    // Gaussian by the backward preprocess instead of once per pixel here:
                    //   dL_dmean2D.x = -o (W/2) sum q u      u = a dx + b dy        (backward.cu:1098-1103)
                    //   dL_dconic.x  = -o/2 sum q dx^2 ...                          (backward.cu:1107-1110)
                    //   dL_dopacity  = sum q,   q = G dL_dalpha                     (backward.cu:1113)
                    const float q = Gv * dL_dalpha;
                    const float mxq = q * fmaf(dy, con_o.y, adx);
                    const float myq = q * fmaf(dy, con_o.z, bdx);
                    v[GA_MX] += mxq;
                    v[GA_MY] += myq;
                    v[GA_AMX] += fabsf(mxq);
                    v[GA_AMY] += fabsf(myq);
                    const float qdx = q * dx, qdy = q * dy;
                    v[GA_CA] = fmaf(qdx, dx, v[GA_CA]);
                    v[GA_CB] = fmaf(qdx, dy, v[GA_CB]);
                    v[GA_CC] = fmaf(qdy, dy, v[GA_CC]);
                    v[GA_OP] += q;
                }
#ifdef BAGS_STATS
                if (lane == 0) atomicAdd(&g_stats[min(nvalid, 63u)], 1ull);
#endif
                if (nvalid == 0u) continue;
                float* rec_out = gacc + (size_t)lds32(s_id + j * 4) * GACC_STRIDE;
                if (REDSM) warp_reduce_commit12_smem(v, (float)nvalid, lane, rec_out,
                                                     smem_addr(sm_red) + (u32)warp * (u32)RED_WARP_BYTES);
                else warp_reduce_commit12(v, (float)nvalid, lane, rec_out);
            }
        }
    }
    if (ASYNC) cp_async_wait<0>();
}

int launch_blend_bwd(int W, int H, const GeomState& g, const BinningState& b, const ImageState& img, const float* bg,
                     const float* dL_dpix, const float* dL_dpix_depth, const float* dL_dalpha, float* gacc,
                     cudaStream_t st, int nviews, const ViewStride& vs, size_t gacc_stride) {
    const int ntiles = tiles_x(W) * tiles_y(H);
    int rc = launch_tile_order(ntiles, nullptr, img.bwd_cost, img.bwd_order, st, nviews, vs);
    if (rc) return rc;
    dim3 grid(ntiles, 1, nviews > 1 ? nviews : 1);
#define BAGS_BWD(PPT, ASYNC)                                                                                     \
    if (g_bwd_reduce) BAGS_BWD_R(PPT, ASYNC, true); else BAGS_BWD_R(PPT, ASYNC, false)
#define BAGS_BWD_R(PPT, ASYNC, REDSM)                                                                            \
    blend_bwd_kernel<PPT, ASYNC, REDSM><<<grid, 256 / PPT, 0, st>>>(img.ranges, b.point_list, W, H, bg, g.rec,   \
                                                            img.final_T, img.n_contrib, dL_dpix, dL_dpix_depth, \
                                                            dL_dalpha, gacc, g_cull != 0, vs, gacc_stride,      \
                                                            img.bwd_order, (u32)tiles_x(W))
    if (g_staging) {
        if (g_bwd_ppt == 4) BAGS_BWD(4, true);
        else if (g_bwd_ppt == 2) BAGS_BWD(2, true);
        else BAGS_BWD(1, true);
    } else {
        if (g_bwd_ppt == 4) BAGS_BWD(4, false);
        else if (g_bwd_ppt == 2) BAGS_BWD(2, false);
        else BAGS_BWD(1, false);
    }
#undef BAGS_BWD
#undef BAGS_BWD_R
    BAGS_LAUNCH_OK("blend_bwd_kernel");
    return 0;
}

}  // namespace bags

#ifdef BAGS_STATS
extern "C" int bags_debug_stats(unsigned long long* out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, bags::g_stats, sizeof(bags::g_stats));
    if (reset) {
        unsigned long long z[64] = {0};
        cudaMemcpyToSymbol(bags::g_stats, z, sizeof(z));
    }
    return 0;
}
#endif

extern "C" int bags_set_tuning(int knob, int value) {
    if (knob == BAGS_TUNE_BLEND_CULL) {
        const int prev = bags::g_cull;
        if (value == 0 || value == 1) bags::g_cull = value;
        return prev;
    }
    if (knob == BAGS_TUNE_BLEND_BWD_REDUCE) {
        const int prev = bags::g_bwd_reduce;
        if (value == 0 || value == 1) bags::g_bwd_reduce = value;
        return prev;
    }
    if (knob == BAGS_TUNE_BLEND_STAGING) {
        const int prev = bags::g_staging;
        if (value == 0 || value == 1) bags::g_staging = value;
        return prev;
    }
    if (knob == BAGS_TUNE_SORT_LOOKBACK) return bags::set_sort_lookback(value);
    if (knob == BAGS_TUNE_SORT_SMALL_TILES) return bags::set_sort_small_tiles(value);
    if (knob == BAGS_TUNE_SORT_HWMATCH) return bags::set_sort_hwmatch(value);
    if (knob == BAGS_TUNE_PBWD_MIN_BLOCKS) return bags::set_pbwd_min_blocks(value);
    int* target = knob == BAGS_TUNE_BLEND_FWD_PPT ? &bags::g_fwd_ppt : knob == BAGS_TUNE_BLEND_BWD_PPT ? &bags::g_bwd_ppt : nullptr;
    if (!target) return BAGS_E_INVALID;
    const int prev = *target;
    if (value == 1 || value == 2 || value == 4) *target = value;
    return prev;
}
