// Forward per-Gaussian preprocess (K1), markVisible (K10) and MCMC relocation (K11) for sm_100a.
//
// Restates, from scratch, what preprocessCUDA / in_frustum / computeCov3D / computeCov2D / computeColorFromSH do in
// the reference (3dgs-pose/cuda_rasterizer/forward.cu:20-96,99-164,169-203,339-657; auxiliary.h:41-56,139-164),
// including its quirks: near-plane-only culling, the entrance-pupil shift added to p_hom.w and t.z, the
// un-normalised quaternion, max(0.1, mid^2-det), ndc2Pix in FP64, and means2Dx/y = p_hom.z / shifted p_hom.w.
//
// What is different from the reference (layout / fusion, not results):
//   * the blend kernels' inputs are packed into one 48-byte record per Gaussian (xy, depth, clamp mask | conic,
//     opacity | rgb) instead of five separate arrays, so a tile-list gather touches two 32-byte sectors;
//   * cov3D is not stored (the backward recomputes it), the tile rectangle is stored (8 B) instead of only its area,
//   * the depth key for the two-level sort and the total tile-instance count are produced here, so no separate
//     InclusiveSum / 4-byte blocking copy sits between preprocess and binning,
//   * every output element (radii, mean2Dx/y) is written, so the binding allocates with empty() not zeros().
#include "../../include/bags_b200.h"
#include "common.cuh"
#include "math.cuh"

namespace bags {

constexpr int PRE_THREADS = 128;
constexpr int PRE_WARPS = PRE_THREADS / 32;

// Row stride (floats) of the per-warp SH staging tile: odd => lane-per-row reads are bank-conflict free.
__host__ __device__ inline int sh_stride(int M) { return (3 * M) | 1; }

// Coalesced load of the SH rows selected by `mask` (one row per Gaussian of the warp) into the warp's tile.
template <int VEC>
__device__ __forceinline__ void sh_tile_load(float* tile, int stride, const float* gptr, int row_len, int rows,
                                             u32 mask, int lane) {
    const int total = rows * row_len;
    for (int f = lane * VEC; f < total; f += 32 * VEC) {
        const int row = f / row_len;
        if (!((mask >> row) & 1u)) continue;
        float* t = tile + row * stride + (f - row * row_len);
        if (VEC == 4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(gptr + f));
            t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
        } else {
            t[0] = __ldg(gptr + f);
        }
    }
}

struct CamConst {
    float view[16];
    float proj[16];
    float campos[3];
    float shift[3];
};

// One launch serves B views of the same Gaussians (B = 1: the single-view forward).  The view-independent part of a
// Gaussian -- mean, scale, rotation -> 3D covariance, and its SH row in the warp's shared-memory tile -- is fetched and
// computed once and used by every view; the per-view outputs go to B stacked copies of the single-view arrays
// (`geom_stride` bytes apart inside the private state, P elements apart for radii / mean2Dx / mean2Dy).  The per-view
// arithmetic is the same code whatever B is, so a batched view is bit-identical to the same view rendered alone.
template <bool HAS_SH, bool HAS_COV>
__global__ void __launch_bounds__(PRE_THREADS, 8)
preprocess_fwd_kernel(int P, int D, int M, const float* __restrict__ means3D, const float* __restrict__ scales,
                      float scale_modifier, const float* __restrict__ rotations, const float* __restrict__ opacities,
                      const float* __restrict__ shs, const float* __restrict__ cov3D_precomp,
                      const float* __restrict__ colors_precomp, const float* __restrict__ viewmatrix,
                      const float* __restrict__ projmatrix, const float* __restrict__ cam_pos,
                      const float* __restrict__ shift_factors, int W, int H, const __grid_constant__ ViewFov fov, const int B,
                      int grid_x, int grid_y, int prefiltered,
                      int* __restrict__ radii, float* __restrict__ mean2Dx, float* __restrict__ mean2Dy,
                      float4* __restrict__ rec0, u32* __restrict__ depth_key0, ushort4* __restrict__ rect_out0,
                      unsigned long long* __restrict__ total_tiles, const size_t geom_stride, int sh_vec4) {
    extern __shared__ float sh_tiles[];      // [PRE_WARPS][32][sh_stride(M)], only used when HAS_SH
    __shared__ CamConst cams[MAX_BATCH];
    __shared__ float c2ws[MAX_BATCH][4][4];
    __shared__ u32 warp_sums[MAX_BATCH][PRE_THREADS / 32];
    __shared__ float4 s_fov[MAX_BATCH];

    fov_to_shared(fov, s_fov);
    for (int t = threadIdx.x; t < 16 * B; t += PRE_THREADS) {
        cams[t >> 4].view[t & 15] = viewmatrix[t];
        cams[t >> 4].proj[t & 15] = projmatrix[t];
    }
    if (threadIdx.x < 3 * B) {
        cams[threadIdx.x / 3].campos[threadIdx.x % 3] = cam_pos[threadIdx.x];
        cams[threadIdx.x / 3].shift[threadIdx.x % 3] = shift_factors[threadIdx.x];
    }
    if (HAS_SH && threadIdx.x >= 32 && threadIdx.x < 32 + B)
        inverse_w2c(viewmatrix + 16 * (threadIdx.x - 32), c2ws[threadIdx.x - 32]);
    __syncthreads();

    const int idx = blockIdx.x * PRE_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // ---- view-independent inputs: fetched together (one memory round trip on the critical path of every warp) -----
    // (the 3D covariance does not depend on the view: evaluated once, also for the ~5% of rows every view culls)
    float3 p_orig = {0.f, 0.f, 0.f};
    float cov6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (idx < P) {
        p_orig = {means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]};
        if (HAS_COV) {
#pragma unroll
            for (int k = 0; k < 6; k++) cov6[k] = cov3D_precomp[6 * idx + k];
        } else {
            const float sc0 = scales[3 * idx], sc1 = scales[3 * idx + 1], sc2 = scales[3 * idx + 2];
            const float q0 = rotations[4 * idx], q1 = rotations[4 * idx + 1], q2 = rotations[4 * idx + 2],
                        q3 = rotations[4 * idx + 3];
            cov3d_from_scale_rot(sc0, sc1, sc2, scale_modifier, q0, q1, q2, q3, cov6);
        }
    }
    u32 sh_loaded = 0;           // rows of the warp's SH tile that are in shared memory

    for (int v = 0; v < B; v++) {
    const CamConst& cam = cams[v];
    const float (*c2w)[4] = c2ws[v];
    const float tan_fovx = s_fov[v].x, tan_fovy = s_fov[v].y;
    const float focal_x = s_fov[v].z, focal_y = s_fov[v].w;
    float4* rec = at_view(rec0, geom_stride, v);
    u32 my_tiles = 0;

    // ---- phase 1: geometry (one Gaussian per thread) ---------------------------------------------------------------
    int out_radius = 0;
    float out_mx = 0.f, out_my = 0.f;
    u32 out_key = 0xFFFFFFFFu;
    ushort4 out_rect = make_ushort4(0, 0, 0, 0);
    bool emit = false;                  // survives every cull -> gets a record and a colour
    float pupil_shift = 0.f;
    float2 point_image = {0.f, 0.f};
    float3 conic = {0.f, 0.f, 0.f};
    float depth_z = 0.f;
    const float* vm = cam.view;
    const float* pm = cam.proj;

    if (idx < P) {
        const float3 p_view = xform4x3(p_orig, vm);

        // Near-plane culling only (auxiliary.h:154).
        bool alive = !(p_view.z <= 0.2f);
        if (!alive && prefiltered) {
            printf("Point is filtered although prefiltered is set. This shouldn't happen!");
            __trap();
        }

        if (alive) {
            float4 p_hom = xform4x4(p_orig, pm);

            // Entrance-pupil shift from the (unshifted) projected direction, forward.cu:392-424.
            float2 ab = {p_hom.x / p_hom.w, p_hom.y / p_hom.w};
            float radius = sqrt(ab.x * ab.x + ab.y * ab.y);
            float theta = atan(radius);
            float theta2 = theta * theta;
            float theta3 = theta2 * theta;
            float theta5 = theta2 * theta3;
            float theta7 = theta2 * theta5;
            pupil_shift = cam.shift[0] * theta3 + cam.shift[1] * theta5 + cam.shift[2] * theta7;
            p_hom.w += pupil_shift;

            float p_w = 1.0f / (p_hom.w + 0.0000001f);
            float3 p_proj = {p_hom.x * p_w, p_hom.y * p_w, p_hom.z * p_w};

            // EWA 2D covariance, forward.cu:99-164 (t.z carries the pupil shift).
            float3 t = xform4x3(p_orig, vm);
            t.z += pupil_shift;
            const float limx = 1.3f * tan_fovx;
            const float limy = 1.3f * tan_fovy;
            const float txtz = t.x / t.z;
            const float tytz = t.y / t.z;
            t.x = min(limx, max(-limx, txtz)) * t.z;
            t.y = min(limy, max(-limy, tytz)) * t.z;

            M3 J = m3_cols(focal_x / t.z, 0.0f, -(focal_x * t.x) / (t.z * t.z),
                           0.0f, focal_y / t.z, -(focal_y * t.y) / (t.z * t.z),
                           0.0f, 0.0f, 0.0f);
            M3 Wm = m3_cols(vm[0], vm[4], vm[8], vm[1], vm[5], vm[9], vm[2], vm[6], vm[10]);
            M3 T = m3_mul(Wm, J);
            M3 Vrk = m3_cols(cov6[0], cov6[1], cov6[2], cov6[1], cov6[3], cov6[4], cov6[2], cov6[4], cov6[5]);
            M3 cov = m3_mul(m3_mul(m3_transpose(T), m3_transpose(Vrk)), T);
            float cov_x = cov.c[0][0] + 0.3f;
            float cov_y = cov.c[0][1];
            float cov_z = cov.c[1][1] + 0.3f;

            float det = (cov_x * cov_z - cov_y * cov_y);
            if (det != 0.0f) {
                float det_inv = 1.f / det;
                conic = {cov_z * det_inv, -cov_y * det_inv, cov_x * det_inv};

                float mid = 0.5f * (cov_x + cov_z);
                float lambda1 = mid + sqrt(max(0.1f, mid * mid - det));
                float lambda2 = mid - sqrt(max(0.1f, mid * mid - det));
                float my_radius = ceil(3.f * sqrt(max(lambda1, lambda2)));
                point_image = {ndc_to_pix(p_proj.x, W), ndc_to_pix(p_proj.y, H)};

                uint2 rmin, rmax;
                tile_rect(point_image.x, point_image.y, (int)my_radius, grid_x, grid_y, rmin, rmax);
                u32 area = (rmax.x - rmin.x) * (rmax.y - rmin.y);
                if (area != 0) {
                    emit = true;
                    out_radius = (int)my_radius;
                    out_mx = p_hom.z;
                    out_my = p_hom.w;
                    depth_z = p_view.z;
                    out_key = __float_as_uint(p_view.z);
                    out_rect = make_ushort4((unsigned short)rmin.x, (unsigned short)rmin.y,
                                            (unsigned short)(rmax.x - rmin.x), (unsigned short)(rmax.y - rmin.y));
                    my_tiles = area;
                }
            }
        }
        const size_t o = (size_t)v * P + idx;
        radii[o] = out_radius;
        mean2Dx[o] = out_mx;
        mean2Dy[o] = out_my;
        at_view(depth_key0, geom_stride, v)[idx] = out_key;
        at_view(rect_out0, geom_stride, v)[idx] = out_rect;
    }

    // ---- phase 2: colour.  SH rows of the surviving Gaussians are staged per warp with coalesced 128-bit loads ------
    // (a later view of a batch only fetches the rows no earlier view needed)
    float3 rgb = {0.f, 0.f, 0.f};
    u32 clamp_mask = 0;
    if (HAS_SH) {
        const int row_len = 3 * M;
        const int stride = sh_stride(M);
        float* tile = sh_tiles + (size_t)warp * 32 * stride;
        const int warp_first = blockIdx.x * PRE_THREADS + warp * 32;
        const int rows = min(32, P - warp_first);
        const u32 need = __ballot_sync(0xffffffffu, emit) & ~sh_loaded;
        if (need != 0 && rows > 0) {
            const float* gsh = shs + (size_t)warp_first * row_len;
            if (sh_vec4) sh_tile_load<4>(tile, stride, gsh, row_len, rows, need, lane);
            else sh_tile_load<1>(tile, stride, gsh, row_len, rows, need, lane);
            sh_loaded |= need;
        }
        __syncwarp();
        if (emit) {
            // forward.cu:20-96: direction from the pupil-shifted camera-space point mapped back to world.
            float pc0 = vm[0] * p_orig.x + vm[4] * p_orig.y + vm[8] * p_orig.z + vm[12] * 1.0f;
            float pc1 = vm[1] * p_orig.x + vm[5] * p_orig.y + vm[9] * p_orig.z + vm[13] * 1.0f;
            float pc2 = vm[2] * p_orig.x + vm[6] * p_orig.y + vm[10] * p_orig.z + vm[14] * 1.0f;
            float pc3 = vm[3] * p_orig.x + vm[7] * p_orig.y + vm[11] * p_orig.z + vm[15] * 1.0f;
            pc2 += pupil_shift;
            float3 pos;
            pos.x = c2w[0][0] * pc0 + c2w[0][1] * pc1 + c2w[0][2] * pc2 + c2w[0][3] * pc3;
            pos.y = c2w[1][0] * pc0 + c2w[1][1] * pc1 + c2w[1][2] * pc2 + c2w[1][3] * pc3;
            pos.z = c2w[2][0] * pc0 + c2w[2][1] * pc1 + c2w[2][2] * pc2 + c2w[2][3] * pc3;
            float3 dir = {pos.x - cam.campos[0], pos.y - cam.campos[1], pos.z - cam.campos[2]};
            float len = sqrt(dir.x * dir.x + dir.y * dir.y + dir.z * dir.z);
            dir.x = dir.x / len;
            dir.y = dir.y / len;
            dir.z = dir.z / len;

            const float* sh = tile + lane * stride;
            float res[3];
            const float x = dir.x, y = dir.y, z = dir.z;
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                float r = BAGS_SH_C0 * sh[ch];
                if (D > 0) {
                    r = r - BAGS_SH_C1 * y * sh[3 + ch] + BAGS_SH_C1 * z * sh[6 + ch] - BAGS_SH_C1 * x * sh[9 + ch];
                    if (D > 1) {
                        float xx = x * x, yy = y * y, zz = z * z;
                        float xy = x * y, yz = y * z, xz = x * z;
                        r = r + BAGS_SH_C2_0 * xy * sh[12 + ch] + BAGS_SH_C2_1 * yz * sh[15 + ch] +
                            BAGS_SH_C2_2 * (2.0f * zz - xx - yy) * sh[18 + ch] + BAGS_SH_C2_3 * xz * sh[21 + ch] +
                            BAGS_SH_C2_4 * (xx - yy) * sh[24 + ch];
                        if (D > 2) {
                            r = r + BAGS_SH_C3_0 * y * (3.0f * xx - yy) * sh[27 + ch] +
                                BAGS_SH_C3_1 * xy * z * sh[30 + ch] +
                                BAGS_SH_C3_2 * y * (4.0f * zz - xx - yy) * sh[33 + ch] +
                                BAGS_SH_C3_3 * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[36 + ch] +
                                BAGS_SH_C3_4 * x * (4.0f * zz - xx - yy) * sh[39 + ch] +
                                BAGS_SH_C3_5 * z * (xx - yy) * sh[42 + ch] +
                                BAGS_SH_C3_6 * x * (xx - 3.0f * yy) * sh[45 + ch];
                        }
                    }
                }
                r += 0.5f;
                if (r < 0) clamp_mask |= (1u << ch);
                res[ch] = max(r, 0.0f);
            }
            rgb = {res[0], res[1], res[2]};
        }
    } else if (emit) {
        rgb = {colors_precomp[3 * idx], colors_precomp[3 * idx + 1], colors_precomp[3 * idx + 2]};
    }
    if (emit) {
        rec[3 * (size_t)idx + 0] = make_float4(point_image.x, point_image.y, depth_z, __uint_as_float(clamp_mask));
        rec[3 * (size_t)idx + 1] = make_float4(conic.x, conic.y, conic.z, opacities[idx]);
        rec[3 * (size_t)idx + 2] = make_float4(rgb.x, rgb.y, rgb.z, 0.f);
    }

    // Warp-reduce the view's tile-instance count (summed per CTA after the view loop).
    u32 s = my_tiles;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) warp_sums[v][warp] = s;
    }   // views

    // One 64-bit atomic per CTA and view replaces the reference's InclusiveSum + blocking 4-byte copy as the source of
    // num_rendered.
    __syncthreads();
    if (threadIdx.x < B) {
        unsigned long long tot = 0;
#pragma unroll
        for (int w = 0; w < PRE_THREADS / 32; w++) tot += warp_sums[threadIdx.x][w];
        if (tot) atomicAdd(at_view(total_tiles, geom_stride, threadIdx.x), tot);
    }
}

// a.viewmatrix / projmatrix / cam_pos / shift_factors: [nviews][16 | 16 | 3 | 3]; radii / mean2Dx / mean2Dy:
// [nviews][P]; g: view 0's state, the others `geom_stride` bytes apart; fov: per-view projection constants (NULL:
// the single view described by `a`).
int launch_preprocess_fwd(const FwdArgs& a, GeomState& g, int* radii, float* mean2Dx, float* mean2Dy,
                          cudaStream_t st, int nviews, const ViewFov* fov_in, size_t geom_stride) {
    const int gx = tiles_x(a.W), gy = tiles_y(a.H);
    const int blocks = (a.P + PRE_THREADS - 1) / PRE_THREADS;
    const bool has_sh = a.colors_precomp == nullptr;
    const bool has_cov = a.cov3D_precomp != nullptr;
    const size_t smem = has_sh ? (size_t)PRE_WARPS * 32 * sh_stride(a.M) * sizeof(float) : 0;
    if (smem > 200 * 1024) {
        set_error("preprocess: %d SH coefficients per Gaussian do not fit the staging tile", a.M);
        return BAGS_E_INVALID;
    }
    if (nviews < 1 || nviews > MAX_BATCH) {
        set_error("preprocess: %d views in one launch (1..%d supported)", nviews, MAX_BATCH);
        return BAGS_E_INVALID;
    }
    ViewFov fov;
    if (fov_in != nullptr) {
        fov = *fov_in;
    } else {
        fov.tan_fovx[0] = a.tan_fovx; fov.tan_fovy[0] = a.tan_fovy;
        fov.focal_x[0] = a.focal_x; fov.focal_y[0] = a.focal_y;
    }
    if (nviews == 1) geom_stride = 0;
    if (smem > 48 * 1024) {
        BAGS_CUDA_OK(cudaFuncSetAttribute(preprocess_fwd_kernel<true, true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        BAGS_CUDA_OK(cudaFuncSetAttribute(preprocess_fwd_kernel<true, false>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const int vec4 = has_sh && ((3 * a.M) % 4 == 0) && (((size_t)a.shs & 15) == 0);
#define BAGS_PRE_ARGS                                                                                              \
    a.P, a.D, a.M, a.means3D, a.scales, a.scale_modifier, a.rotations, a.opacities, a.shs, a.cov3D_precomp,        \
        a.colors_precomp, a.viewmatrix, a.projmatrix, a.cam_pos, a.shift_factors, a.W, a.H, fov, nviews, gx, gy,   \
        a.prefiltered, radii, mean2Dx, mean2Dy, g.rec, g.key_a, g.rect, g.counters, geom_stride, vec4
    if (has_sh && has_cov)
        preprocess_fwd_kernel<true, true><<<blocks, PRE_THREADS, smem, st>>>(BAGS_PRE_ARGS);
    else if (has_sh)
        preprocess_fwd_kernel<true, false><<<blocks, PRE_THREADS, smem, st>>>(BAGS_PRE_ARGS);
    else if (has_cov)
        preprocess_fwd_kernel<false, true><<<blocks, PRE_THREADS, smem, st>>>(BAGS_PRE_ARGS);
    else
        preprocess_fwd_kernel<false, false><<<blocks, PRE_THREADS, smem, st>>>(BAGS_PRE_ARGS);
#undef BAGS_PRE_ARGS
    BAGS_LAUNCH_OK("preprocess_fwd_kernel");
    return 0;
}

// ---- markVisible (rasterizer_impl.cu:54-66) --------------------------------------------------------------------
__global__ void mark_visible_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ viewmatrix,
                                    uint8_t* __restrict__ present) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P) return;
    float3 p = {means3D[3 * idx], means3D[3 * idx + 1], means3D[3 * idx + 2]};
    float3 pv = xform4x3(p, viewmatrix);
    present[idx] = (pv.z <= 0.2f) ? 0 : 1;
}

// ---- MCMC relocation (utils.cu:6-36) ---------------------------------------------------------------------------
__global__ void relocation_kernel(int P, const float* __restrict__ opacity_old, const float* __restrict__ scale_old,
                                  const int* __restrict__ N, const float* __restrict__ binoms, int n_max,
                                  float* __restrict__ opacity_new, float* __restrict__ scale_new) {
    int idx = threadIdx.x + blockIdx.x * blockDim.x;
    if (idx >= P) return;
    int N_idx = N[idx];
    float denom_sum = 0.0f;
    float o_new = 1.0f - powf(1.0f - opacity_old[idx], 1.0f / N_idx);
    opacity_new[idx] = o_new;
    for (int i = 1; i <= N_idx; ++i) {
        for (int k = 0; k <= (i - 1); ++k) {
            float bin_coeff = binoms[(i - 1) * n_max + k];
            // utils.cu:29 calls pow/sqrt with integer arguments, which promote to double; the product is rounded to float
            float term = (float)((pow(-1.0, (double)k) / sqrt((double)(k + 1))) * pow((double)o_new, (double)(k + 1)));
            denom_sum += (bin_coeff * term);
        }
    }
    float coeff = (opacity_old[idx] / denom_sum);
    for (int i = 0; i < 3; ++i) scale_new[idx * 3 + i] = coeff * scale_old[idx * 3 + i];
}

}  // namespace bags

using namespace bags;

extern "C" int bags_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                                 uint8_t* present, void* stream) {
    (void)projmatrix;
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) {
        set_error("bags_mark_visible: invalid argument");
        return BAGS_E_INVALID;
    }
    if (P == 0) return 0;
    mark_visible_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, means3D, viewmatrix, present);
    BAGS_LAUNCH_OK("mark_visible_kernel");
    return 0;
}

extern "C" int bags_compute_relocation(int P, const float* opacity_old, const float* scale_old, const int* N,
                                       const float* binoms, int n_max, float* opacity_new, float* scale_new,
                                       void* stream) {
    if (P < 0 || (P > 0 && (!opacity_old || !scale_old || !N || !binoms || !opacity_new || !scale_new))) {
        set_error("bags_compute_relocation: invalid argument");
        return BAGS_E_INVALID;
    }
    if (P == 0) return 0;
    relocation_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream>>>(P, opacity_old, scale_old, N, binoms, n_max,
                                                                         opacity_new, scale_new);
    BAGS_LAUNCH_OK("relocation_kernel");
    return 0;
}
