// Fused backward preprocess for sm_100a: the reference's computeCov2DCUDA (K8) + preprocessCUDA (K9) + the four
// .mean(0) pose reductions + 16 zero-filled scratch tensors, in one pass over the Gaussians.
//
// Restates 3dgs-pose/cuda_rasterizer/backward.cu:37-170 (SH), :175-373 (2D covariance), :377-440 (scale/rotation),
// :548-922 (projection, depth, pose terms) and rasterize_points.cu:170-185,274 (zeros + .mean(0) + covariance+1e-4),
// quirks included:
//   * the 2D covariance is recomputed WITHOUT the entrance-pupil shift, and the SH direction likewise (Q2);
//   * dL_dmeans3D's projection term carries the extra integer factors (W/2) and (H/2) on top of the NDC-unit
//     dL_dmean2D (Q9); the projection-matrix gradient does not (Q10: rows 2,6,10,14 stay zero);
//   * dL_dshift's projection term re-uses theta^3 in the first summand of all three components and is evaluated in
//     FP64 through the (-1.) literals (Q11);
//   * pose / intrinsic gradients are means over all P rows, culled rows counting as zero (Q12);
//   * the reference's dL_dviewmatrix is a data race in which every visible thread overwrites row 0
//     (backward.cu:351-363); after .mean(0) it is (the 8 terms of one arbitrary visible Gaussian)/P.  We make the
//     deterministic choice "highest-index visible Gaussian" (Q13).
//
// Every dense output row is written here (zeros for culled Gaussians), so the binding allocates with empty().
// The per-Gaussian P x (16+16+3+3) pose scratch of the reference does not exist: pose terms are reduced per warp /
// CTA and committed with one atomic per CTA and term.
#include "../../include/bags_b200.h"
#include "common.cuh"
#include "math.cuh"

namespace bags {

constexpr int PB_THREADS = 128;
constexpr int PB_WARPS = PB_THREADS / 32;
// pose_acc layout (doubles): [0..15] sum dL_dproj, [16..18] sum dL_dcampos, [19..21] sum dL_dshift,
// [22] (as u32) 1 + highest visible index.  Sums are carried in FP64 from the warp butterfly on: the terms are
// signed and cancel heavily (condition numbers ~100 on wide-splat scenes), so the summation itself adds no error.
constexpr int PA_PROJ = 0, PA_CAMPOS = 16, PA_SHIFT = 19, PA_WINNER = 22, PA_NRED = 22;

// Row stride (floats) of the per-warp SH staging tile: odd, so that lane-per-row access is bank-conflict free.
__host__ __device__ inline int sh_tile_stride(int M) { return (3 * M) | 1; }

// Cooperative, coalesced copy between the per-warp SH tile (32 rows, one per Gaussian of the warp) and the
// contiguous global block of the same 32 Gaussians.  `rows` = number of valid Gaussians (tail warp), `mask` selects
// rows (loads and accumulating stores; plain stores are dense).  VEC = 4 needs 3M % 4 == 0 and a 16-byte aligned
// base.  SH_ACCUM adds the tile rows selected by `mask` onto what global memory holds (fp32 round-to-nearest adds,
// the same operation as autograd's `grad += new`) and leaves the other rows untouched.
enum { SH_LOAD = 0, SH_STORE = 1, SH_ACCUM = 2, SH_STORE_MASKED = 3, SH_LOAD_RW = 4 };
template <int VEC, int MODE>
__device__ __forceinline__ void sh_tile_copy(float* tile, int stride, float* gptr, int row_len, int rows, u32 mask,
                                             int lane) {
    const int total = rows * row_len;
    for (int f = lane * VEC; f < total; f += 32 * VEC) {
        const int row = f / row_len;
        const int col = f - row * row_len;
        float* t = tile + row * stride + col;
        if (MODE == SH_LOAD) {
            if (!((mask >> row) & 1u)) continue;
            if (VEC == 4) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(gptr + f));
                t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
            } else {
                t[0] = __ldg(gptr + f);
            }
        } else if (MODE == SH_ACCUM) {
            if (!((mask >> row) & 1u)) continue;
            if (VEC == 4) {
                float4 v = *reinterpret_cast<const float4*>(gptr + f);
                v.x = __fadd_rn(v.x, t[0]); v.y = __fadd_rn(v.y, t[1]);
                v.z = __fadd_rn(v.z, t[2]); v.w = __fadd_rn(v.w, t[3]);
                *reinterpret_cast<float4*>(gptr + f) = v;
            } else {
                gptr[f] = __fadd_rn(gptr[f], t[0]);
            }
        } else if (MODE == SH_LOAD_RW) {      // like SH_LOAD, of memory this kernel also writes (no read-only path)
            if (!((mask >> row) & 1u)) continue;
            if (VEC == 4) {
                const float4 v = *reinterpret_cast<const float4*>(gptr + f);
                t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
            } else {
                t[0] = gptr[f];
            }
        } else {
            if (MODE == SH_STORE_MASKED && !((mask >> row) & 1u)) continue;
            if (VEC == 4) {
                *reinterpret_cast<float4*>(gptr + f) = make_float4(t[0], t[1], t[2], t[3]);
            } else {
                gptr[f] = t[0];
            }
        }
    }
}

// Scales that turn the raw moments accumulated by the blend backward into the reference's per-splat gradients.
struct GradScale { float sx, sy, sc, trace; };
__device__ __forceinline__ GradScale grad_scale(const float4 conic_opacity, const int W, const int H) {
    GradScale g;
    g.sx = conic_opacity.w * (0.5f * (float)W);
    g.sy = conic_opacity.w * (0.5f * (float)H);
    g.sc = -0.5f * conic_opacity.w;
    g.trace = conic_opacity.x + conic_opacity.z;
    return g;
}

struct Cov2DBack {
    float dL_dcov[6];
    float dL_dT00, dL_dT01, dL_dT02, dL_dT10, dL_dT11, dL_dT12;
    float dL_dtx, dL_dty, dL_dtz;
    float tz, tz2;
};

// backward.cu:195-327 of the reference, for one Gaussian.
__device__ __forceinline__ Cov2DBack cov2d_backward(const float3 mean, const float* cov3D, float h_x, float h_y,
                                                    float tan_fovx, float tan_fovy, const float* view_matrix,
                                                    const float3 dL_dconic) {
    Cov2DBack o;
    float3 t = xform4x3(mean, view_matrix);
    const float limx = 1.3f * tan_fovx;
    const float limy = 1.3f * tan_fovy;
    const float txtz = t.x / t.z;
    const float tytz = t.y / t.z;
    t.x = min(limx, max(-limx, txtz)) * t.z;
    t.y = min(limy, max(-limy, tytz)) * t.z;
    const float x_grad_mul = txtz < -limx || txtz > limx ? 0 : 1;
    const float y_grad_mul = tytz < -limy || tytz > limy ? 0 : 1;

    M3 J = m3_cols(h_x / t.z, 0.0f, -(h_x * t.x) / (t.z * t.z),
                   0.0f, h_y / t.z, -(h_y * t.y) / (t.z * t.z),
                   0.0f, 0.0f, 0.0f);
    M3 Wm = m3_cols(view_matrix[0], view_matrix[4], view_matrix[8], view_matrix[1], view_matrix[5], view_matrix[9],
                    view_matrix[2], view_matrix[6], view_matrix[10]);
    M3 Vrk = m3_cols(cov3D[0], cov3D[1], cov3D[2], cov3D[1], cov3D[3], cov3D[4], cov3D[2], cov3D[4], cov3D[5]);
    M3 T = m3_mul(Wm, J);
    M3 cov2D = m3_mul(m3_mul(m3_transpose(T), m3_transpose(Vrk)), T);

    const float a = cov2D.c[0][0] + 0.3f;
    const float b = cov2D.c[0][1];
    const float c = cov2D.c[1][1] + 0.3f;

    const float denom = a * c - b * b;
    float dL_da = 0, dL_db = 0, dL_dc = 0;
    const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);

    if (denom2inv != 0) {
        dL_da = denom2inv * (-c * c * dL_dconic.x + 2 * b * c * dL_dconic.y + (denom - a * c) * dL_dconic.z);
        dL_dc = denom2inv * (-a * a * dL_dconic.z + 2 * a * b * dL_dconic.y + (denom - a * c) * dL_dconic.x);
        dL_db = denom2inv * 2 * (b * c * dL_dconic.x - (denom + 2 * b * b) * dL_dconic.y + a * b * dL_dconic.z);

        o.dL_dcov[0] = (T.c[0][0] * T.c[0][0] * dL_da + T.c[0][0] * T.c[1][0] * dL_db + T.c[1][0] * T.c[1][0] * dL_dc);
        o.dL_dcov[3] = (T.c[0][1] * T.c[0][1] * dL_da + T.c[0][1] * T.c[1][1] * dL_db + T.c[1][1] * T.c[1][1] * dL_dc);
        o.dL_dcov[5] = (T.c[0][2] * T.c[0][2] * dL_da + T.c[0][2] * T.c[1][2] * dL_db + T.c[1][2] * T.c[1][2] * dL_dc);
        o.dL_dcov[1] = 2 * T.c[0][0] * T.c[0][1] * dL_da + (T.c[0][0] * T.c[1][1] + T.c[0][1] * T.c[1][0]) * dL_db +
                       2 * T.c[1][0] * T.c[1][1] * dL_dc;
        o.dL_dcov[2] = 2 * T.c[0][0] * T.c[0][2] * dL_da + (T.c[0][0] * T.c[1][2] + T.c[0][2] * T.c[1][0]) * dL_db +
                       2 * T.c[1][0] * T.c[1][2] * dL_dc;
        o.dL_dcov[4] = 2 * T.c[0][2] * T.c[0][1] * dL_da + (T.c[0][1] * T.c[1][2] + T.c[0][2] * T.c[1][1]) * dL_db +
                       2 * T.c[1][1] * T.c[1][2] * dL_dc;
    } else {
#pragma unroll
        for (int i = 0; i < 6; i++) o.dL_dcov[i] = 0;
    }

    o.dL_dT00 = 2 * (T.c[0][0] * Vrk.c[0][0] + T.c[0][1] * Vrk.c[0][1] + T.c[0][2] * Vrk.c[0][2]) * dL_da +
                (T.c[1][0] * Vrk.c[0][0] + T.c[1][1] * Vrk.c[0][1] + T.c[1][2] * Vrk.c[0][2]) * dL_db;
    o.dL_dT01 = 2 * (T.c[0][0] * Vrk.c[1][0] + T.c[0][1] * Vrk.c[1][1] + T.c[0][2] * Vrk.c[1][2]) * dL_da +
                (T.c[1][0] * Vrk.c[1][0] + T.c[1][1] * Vrk.c[1][1] + T.c[1][2] * Vrk.c[1][2]) * dL_db;
    o.dL_dT02 = 2 * (T.c[0][0] * Vrk.c[2][0] + T.c[0][1] * Vrk.c[2][1] + T.c[0][2] * Vrk.c[2][2]) * dL_da +
                (T.c[1][0] * Vrk.c[2][0] + T.c[1][1] * Vrk.c[2][1] + T.c[1][2] * Vrk.c[2][2]) * dL_db;
    o.dL_dT10 = 2 * (T.c[1][0] * Vrk.c[0][0] + T.c[1][1] * Vrk.c[0][1] + T.c[1][2] * Vrk.c[0][2]) * dL_dc +
                (T.c[0][0] * Vrk.c[0][0] + T.c[0][1] * Vrk.c[0][1] + T.c[0][2] * Vrk.c[0][2]) * dL_db;
    o.dL_dT11 = 2 * (T.c[1][0] * Vrk.c[1][0] + T.c[1][1] * Vrk.c[1][1] + T.c[1][2] * Vrk.c[1][2]) * dL_dc +
                (T.c[0][0] * Vrk.c[1][0] + T.c[0][1] * Vrk.c[1][1] + T.c[0][2] * Vrk.c[1][2]) * dL_db;
    o.dL_dT12 = 2 * (T.c[1][0] * Vrk.c[2][0] + T.c[1][1] * Vrk.c[2][1] + T.c[1][2] * Vrk.c[2][2]) * dL_dc +
                (T.c[0][0] * Vrk.c[2][0] + T.c[0][1] * Vrk.c[2][1] + T.c[0][2] * Vrk.c[2][2]) * dL_db;

    const float dL_dJ00 = Wm.c[0][0] * o.dL_dT00 + Wm.c[0][1] * o.dL_dT01 + Wm.c[0][2] * o.dL_dT02;
    const float dL_dJ02 = Wm.c[2][0] * o.dL_dT00 + Wm.c[2][1] * o.dL_dT01 + Wm.c[2][2] * o.dL_dT02;
    const float dL_dJ11 = Wm.c[1][0] * o.dL_dT10 + Wm.c[1][1] * o.dL_dT11 + Wm.c[1][2] * o.dL_dT12;
    const float dL_dJ12 = Wm.c[2][0] * o.dL_dT10 + Wm.c[2][1] * o.dL_dT11 + Wm.c[2][2] * o.dL_dT12;

    const float tz = 1.f / t.z;
    const float tz2 = tz * tz;
    const float tz3 = tz2 * tz;
    o.tz = tz;
    o.tz2 = tz2;

    o.dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
    o.dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
    o.dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t.x) * tz3 * dL_dJ02 +
               (2 * h_y * t.y) * tz3 * dL_dJ12;
    return o;
}

__device__ __forceinline__ void load_cov3d(int idx, const float* cov3D_precomp, const float* scales,
                                           const float* rotations, float scale_modifier, float* cov6) {
    if (cov3D_precomp) {
#pragma unroll
        for (int k = 0; k < 6; k++) cov6[k] = cov3D_precomp[6 * (size_t)idx + k];
    } else {
        cov3d_from_scale_rot(scales[3 * (size_t)idx], scales[3 * (size_t)idx + 1], scales[3 * (size_t)idx + 2],
                             scale_modifier, rotations[4 * (size_t)idx], rotations[4 * (size_t)idx + 1],
                             rotations[4 * (size_t)idx + 2], rotations[4 * (size_t)idx + 3], cov6);
    }
}

// Everything the backward computes for ONE visible Gaussian of ONE view (K8 + K9 + SH + scale / rotation), shared by
// the single-view kernel and the batched one.  sh_in: the Gaussian's SH row; dsh_out: where dL_dSH goes -- SH_ADD false:
// written (may alias sh_in: in place), SH_ADD true: added with __fadd_rn onto what the row holds (the batched kernel sums
// the views there).  pose[]: this Gaussian's terms of the pose / intrinsic sums.
struct GaussGrad {
    float mean3D[3], mean2D[2], dens[2], color[3], opacity, covariance, cov3D[6], scale[3], rot[4];
};
template <bool SH_ADD>
__device__ __forceinline__ void sh_put(float* p, const float v) {
    if (SH_ADD) *p = __fadd_rn(*p, v);
    else *p = v;
}
template <bool SH_ADD>
__device__ __forceinline__ void gaussian_backward(const int idx, const int D, const int M, const int W, const int H,
                                                  const float* __restrict__ means3D, const float* __restrict__ shs,
                                                  const float* __restrict__ scales,
                                                  const float* __restrict__ rotations, const float scale_modifier,
                                                  const float* __restrict__ cov3D_precomp, const float* view,
                                                  const float* proj, const float* campos, const float tan_fovx,
                                                  const float tan_fovy, const float focal_x, const float focal_y,
                                                  const float4* __restrict__ rec, const float* __restrict__ gacc,
                                                  const float* sh_in, float* dsh_out, GaussGrad& out,
                                                  float (&pose)[PA_NRED]) {
    const size_t i = (size_t)idx;
#pragma unroll
    for (int k = 0; k < 3; k++) out.scale[k] = 0.f;     // stay zero with a precomputed 3D covariance
#pragma unroll
    for (int k = 0; k < 4; k++) out.rot[k] = 0.f;
    const float4* ga = reinterpret_cast<const float4*>(gacc + i * GACC_STRIDE);
    const float4 g0 = ga[0], g1 = ga[1], g2 = ga[2];
    const float g_cov = gacc[i * GACC_STRIDE + GA_COV];
    // The blend backward accumulates raw per-splat moments (see blend.cu); the per-splat constants of
    // backward.cu:1002-1003,1098-1116 are applied here, once per Gaussian: opacity, ddelx_dx = 0.5 W,
    // ddely_dy = 0.5 H, the -0.5 of the conic terms, and conic.x + conic.z per contributing pixel.
    const GradScale gs = grad_scale(rec[3 * i + 1], W, H);
    float dL_dcolor[3] = {g0.x, g0.y, g0.z};
    const float dL_ddepth = g0.w;
    const float dmx = -gs.sx * g1.x, dmy = -gs.sy * g1.y;          // dL_dmean2D (NDC units)
    const float3 dL_dconic = {gs.sc * g2.x, gs.sc * g2.y, gs.sc * g2.z};
    out.mean2D[0] = dmx; out.mean2D[1] = dmy;
    out.dens[0] = gs.sx * g1.z; out.dens[1] = gs.sy * g1.w;
    out.color[0] = g0.x; out.color[1] = g0.y; out.color[2] = g0.z;
    out.opacity = g2.w;
    out.covariance = g_cov * gs.trace + 0.0001f;

    const float3 m = {means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]};
    float cov6[6];
    load_cov3d(idx, cov3D_precomp, scales, rotations, scale_modifier, cov6);

    // ---- K8: conic -> cov2D -> cov3D / mean ------------------------------------------------------------
    const Cov2DBack cb = cov2d_backward(m, cov6, focal_x, focal_y, tan_fovx, tan_fovy, view, dL_dconic);
#pragma unroll
    for (int k = 0; k < 6; k++) out.cov3D[k] = cb.dL_dcov[k];

    const float4 p_hom = xform4x4(m, proj);
    const float2 ab = {p_hom.x / p_hom.w, p_hom.y / p_hom.w};
    const float theta = atan(sqrt(ab.x * ab.x + ab.y * ab.y));
    const float theta2 = theta * theta;
    const float theta3 = theta2 * theta;
    const float theta5 = theta2 * theta3;
    const float theta7 = theta2 * theta5;
    float dshift[3] = {cb.dL_dtz * theta3, cb.dL_dtz * theta5, cb.dL_dtz * theta7};

    float3 dL_dmean = xformVec4x3T({cb.dL_dtx, cb.dL_dty, cb.dL_dtz}, view);
    float dmean[3] = {dL_dmean.x, dL_dmean.y, dL_dmean.z};

    // ---- K9: projection part ---------------------------------------------------------------------------
    const float4 m_hom = p_hom;
    const float m_w = 1.0f / (m_hom.w + 0.0000001f);
    const float scaling_radial = 1.;
    const int image_width = W, image_height = H;
    const float mul1 = (proj[0] * m.x + proj[4] * m.y + proj[8] * m.z + proj[12]) * m_w * m_w;
    const float mul2 = (proj[1] * m.x + proj[5] * m.y + proj[9] * m.z + proj[13]) * m_w * m_w;
    float pm[3];
    pm[0] = scaling_radial * (image_width / 2) * (proj[0] * m_w - proj[3] * mul1) * dmx +
            scaling_radial * (image_height / 2) * (proj[1] * m_w - proj[3] * mul2) * dmy;
    pm[1] = scaling_radial * (image_width / 2) * (proj[4] * m_w - proj[7] * mul1) * dmx +
            scaling_radial * (image_height / 2) * (proj[5] * m_w - proj[7] * mul2) * dmy;
    pm[2] = scaling_radial * (image_width / 2) * (proj[8] * m_w - proj[11] * mul1) * dmx +
            scaling_radial * (image_height / 2) * (proj[9] * m_w - proj[11] * mul2) * dmy;

    // Q11: the first summand uses theta3 in all three components; FP64 through the (-1.) literal.
    dshift[0] += (float)(dmx * (image_width / 2) * m_hom.x * (-1. / (m_w * m_w)) * theta3 +
                         dmy * (image_height / 2) * m_hom.y * (-1. / (m_w * m_w)) * theta3);
    dshift[1] += (float)(dmx * (image_width / 2) * m_hom.x * (-1. / (m_w * m_w)) * theta3 +
                         dmy * (image_height / 2) * m_hom.y * (-1. / (m_w * m_w)) * theta5);
    dshift[2] += (float)(dmx * (image_width / 2) * m_hom.x * (-1. / (m_w * m_w)) * theta3 +
                         dmy * (image_height / 2) * m_hom.y * (-1. / (m_w * m_w)) * theta7);

    pose[PA_PROJ + 0] = scaling_radial * m_w * m.x * dmx;
    pose[PA_PROJ + 4] = scaling_radial * m_w * m.y * dmx;
    pose[PA_PROJ + 8] = scaling_radial * m_w * m.z * dmx;
    pose[PA_PROJ + 12] = scaling_radial * m_w * dmx;
    pose[PA_PROJ + 1] = scaling_radial * m_w * m.x * dmy;
    pose[PA_PROJ + 5] = scaling_radial * m_w * m.y * dmy;
    pose[PA_PROJ + 9] = scaling_radial * m_w * m.z * dmy;
    pose[PA_PROJ + 13] = scaling_radial * m_w * dmy;
    pose[PA_PROJ + 3] = scaling_radial * (m_w * m_w * (-1.) * m_hom.x * m.x * dmx +
                                          m_w * m_w * (-1.) * m_hom.y * m.x * dmy);
    pose[PA_PROJ + 7] = scaling_radial * (m_w * m_w * (-1.) * m_hom.x * m.y * dmx +
                                          m_w * m_w * (-1.) * m_hom.y * m.y * dmy);
    pose[PA_PROJ + 11] = scaling_radial * (m_w * m_w * (-1.) * m_hom.x * m.z * dmx +
                                           m_w * m_w * (-1.) * m_hom.y * m.z * dmy);
    pose[PA_PROJ + 15] = scaling_radial * (m_w * m_w * (-1.) * m_hom.x * dmx +
                                           m_w * m_w * (-1.) * m_hom.y * dmy);

    dmean[0] += pm[0];
    dmean[1] += pm[1];
    dmean[2] += pm[2];
    // depth -> mean (backward.cu:910-912)
    dmean[0] += view[2] * dL_ddepth;
    dmean[1] += view[6] * dL_ddepth;
    dmean[2] += view[10] * dL_ddepth;

    // ---- SH backward (backward.cu:37-170) --------------------------------------------------------------
    if (shs) {
        const u32 clamp_mask = __float_as_uint(rec[3 * i].w);
        const float3 dir_orig = {m.x - campos[0], m.y - campos[1], m.z - campos[2]};
        const float len = sqrt(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
        const float x = dir_orig.x / len, y = dir_orig.y / len, z = dir_orig.z / len;
        const float* sh = sh_in;
        float* dsh = dsh_out;            // may alias sh_in: every slot is read before it is overwritten
        float dRGB[3];
#pragma unroll
        for (int ch = 0; ch < 3; ch++) dRGB[ch] = dL_dcolor[ch] * ((clamp_mask >> ch) & 1u ? 0 : 1);

        float ddir[3] = {0.f, 0.f, 0.f};  // dL/d(dir) accumulated over channels
        const float xx = x * x, yy = y * y, zz = z * z;
        const float xy = x * y, yz = y * z, xz = x * z;
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            const float g = dRGB[ch];
            float dx = 0.f, dy = 0.f, dz = 0.f;
            // read this channel's coefficients first, then overwrite the same slots with their gradients
            float c[16];
#pragma unroll
            for (int k = 1; k < 16; k++) c[k] = (k < (D + 1) * (D + 1)) ? sh[k * 3 + ch] : 0.f;
            sh_put<SH_ADD>(dsh + 0 * 3 + ch, BAGS_SH_C0 * g);
            if (D > 0) {
                sh_put<SH_ADD>(dsh + 1 * 3 + ch, (-BAGS_SH_C1 * y) * g);
                sh_put<SH_ADD>(dsh + 2 * 3 + ch, (BAGS_SH_C1 * z) * g);
                sh_put<SH_ADD>(dsh + 3 * 3 + ch, (-BAGS_SH_C1 * x) * g);
                dx = -BAGS_SH_C1 * c[3];
                dy = -BAGS_SH_C1 * c[1];
                dz = BAGS_SH_C1 * c[2];
                if (D > 1) {
                    sh_put<SH_ADD>(dsh + 4 * 3 + ch, (BAGS_SH_C2_0 * xy) * g);
                    sh_put<SH_ADD>(dsh + 5 * 3 + ch, (BAGS_SH_C2_1 * yz) * g);
                    sh_put<SH_ADD>(dsh + 6 * 3 + ch, (BAGS_SH_C2_2 * (2.f * zz - xx - yy)) * g);
                    sh_put<SH_ADD>(dsh + 7 * 3 + ch, (BAGS_SH_C2_3 * xz) * g);
                    sh_put<SH_ADD>(dsh + 8 * 3 + ch, (BAGS_SH_C2_4 * (xx - yy)) * g);
                    const float s4 = c[4], s5 = c[5], s6 = c[6], s7 = c[7], s8 = c[8];
                    dx += BAGS_SH_C2_0 * y * s4 + BAGS_SH_C2_2 * 2.f * -x * s6 + BAGS_SH_C2_3 * z * s7 +
                          BAGS_SH_C2_4 * 2.f * x * s8;
                    dy += BAGS_SH_C2_0 * x * s4 + BAGS_SH_C2_1 * z * s5 + BAGS_SH_C2_2 * 2.f * -y * s6 +
                          BAGS_SH_C2_4 * 2.f * -y * s8;
                    dz += BAGS_SH_C2_1 * y * s5 + BAGS_SH_C2_2 * 2.f * 2.f * z * s6 + BAGS_SH_C2_3 * x * s7;
                    if (D > 2) {
                        sh_put<SH_ADD>(dsh + 9 * 3 + ch, (BAGS_SH_C3_0 * y * (3.f * xx - yy)) * g);
                        sh_put<SH_ADD>(dsh + 10 * 3 + ch, (BAGS_SH_C3_1 * xy * z) * g);
                        sh_put<SH_ADD>(dsh + 11 * 3 + ch, (BAGS_SH_C3_2 * y * (4.f * zz - xx - yy)) * g);
                        sh_put<SH_ADD>(dsh + 12 * 3 + ch, (BAGS_SH_C3_3 * z * (2.f * zz - 3.f * xx - 3.f * yy)) * g);
                        sh_put<SH_ADD>(dsh + 13 * 3 + ch, (BAGS_SH_C3_4 * x * (4.f * zz - xx - yy)) * g);
                        sh_put<SH_ADD>(dsh + 14 * 3 + ch, (BAGS_SH_C3_5 * z * (xx - yy)) * g);
                        sh_put<SH_ADD>(dsh + 15 * 3 + ch, (BAGS_SH_C3_6 * x * (xx - 3.f * yy)) * g);
                        const float s9 = c[9], s10 = c[10], s11 = c[11], s12 = c[12], s13 = c[13],
                                    s14 = c[14], s15 = c[15];
                        dx += (BAGS_SH_C3_0 * s9 * 3.f * 2.f * xy + BAGS_SH_C3_1 * s10 * yz +
                               BAGS_SH_C3_2 * s11 * -2.f * xy + BAGS_SH_C3_3 * s12 * -3.f * 2.f * xz +
                               BAGS_SH_C3_4 * s13 * (-3.f * xx + 4.f * zz - yy) +
                               BAGS_SH_C3_5 * s14 * 2.f * xz + BAGS_SH_C3_6 * s15 * 3.f * (xx - yy));
                        dy += (BAGS_SH_C3_0 * s9 * 3.f * (xx - yy) + BAGS_SH_C3_1 * s10 * xz +
                               BAGS_SH_C3_2 * s11 * (-3.f * yy + 4.f * zz - xx) +
                               BAGS_SH_C3_3 * s12 * -3.f * 2.f * yz + BAGS_SH_C3_4 * s13 * -2.f * xy +
                               BAGS_SH_C3_5 * s14 * -2.f * yz + BAGS_SH_C3_6 * s15 * -3.f * 2.f * xy);
                        dz += (BAGS_SH_C3_1 * s10 * xy + BAGS_SH_C3_2 * s11 * 4.f * 2.f * yz +
                               BAGS_SH_C3_3 * s12 * 3.f * (2.f * zz - xx - yy) +
                               BAGS_SH_C3_4 * s13 * 4.f * 2.f * xz + BAGS_SH_C3_5 * s14 * (xx - yy));
                    }
                }
            }
            ddir[0] += dx * g;
            ddir[1] += dy * g;
            ddir[2] += dz * g;
        }
        // coefficients above the active degree get exact zeros (reference: torch::zeros)
        const int active = (D + 1) * (D + 1);
        for (int k = active; k < M; k++) {
            dsh[k * 3 + 0] = 0.f;
            dsh[k * 3 + 1] = 0.f;
            dsh[k * 3 + 2] = 0.f;
        }
        // normalisation of the direction (auxiliary.h:107-117)
        const float3 v = dir_orig;
        const float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
        const float invsum32 = 1.0f / sqrt(sum2 * sum2 * sum2);
        float3 dm;
        dm.x = ((+sum2 - v.x * v.x) * ddir[0] - v.y * v.x * ddir[1] - v.z * v.x * ddir[2]) * invsum32;
        dm.y = (-v.x * v.y * ddir[0] + (sum2 - v.y * v.y) * ddir[1] - v.z * v.y * ddir[2]) * invsum32;
        dm.z = (-v.x * v.z * ddir[0] - v.y * v.z * ddir[1] + (sum2 - v.z * v.z) * ddir[2]) * invsum32;
        dmean[0] += dm.x;
        dmean[1] += dm.y;
        dmean[2] += dm.z;
        pose[PA_CAMPOS + 0] = -dm.x;
        pose[PA_CAMPOS + 1] = -dm.y;
        pose[PA_CAMPOS + 2] = -dm.z;
    }

    // ---- cov3D -> scale / rotation (backward.cu:377-440) ----------------------------------------------
    if (scales) {
        const float sx0 = scales[3 * i], sy0 = scales[3 * i + 1], sz0 = scales[3 * i + 2];
        const float r = rotations[4 * i], x = rotations[4 * i + 1], y = rotations[4 * i + 2],
                    z = rotations[4 * i + 3];
        M3 R = m3_cols(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                       2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                       2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
        M3 S = m3_cols(1.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, 1.0f);
        const float s0 = scale_modifier * sx0, s1 = scale_modifier * sy0, s2 = scale_modifier * sz0;
        S.c[0][0] = s0;
        S.c[1][1] = s1;
        S.c[2][2] = s2;
        M3 Mm = m3_mul(S, R);
        const float* dc = out.cov3D;
        M3 dL_dSigma = m3_cols(dc[0], 0.5f * dc[1], 0.5f * dc[2], 0.5f * dc[1], dc[3], 0.5f * dc[4],
                               0.5f * dc[2], 0.5f * dc[4], dc[5]);
        M3 dL_dM = m3_mul(m3_scale(2.0f, Mm), dL_dSigma);
        M3 Rt = m3_transpose(R);
        M3 dL_dMt = m3_transpose(dL_dM);
        out.scale[0] = Rt.c[0][0] * dL_dMt.c[0][0] + Rt.c[0][1] * dL_dMt.c[0][1] + Rt.c[0][2] * dL_dMt.c[0][2];
        out.scale[1] = Rt.c[1][0] * dL_dMt.c[1][0] + Rt.c[1][1] * dL_dMt.c[1][1] + Rt.c[1][2] * dL_dMt.c[1][2];
        out.scale[2] = Rt.c[2][0] * dL_dMt.c[2][0] + Rt.c[2][1] * dL_dMt.c[2][1] + Rt.c[2][2] * dL_dMt.c[2][2];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            dL_dMt.c[0][k] *= s0;
            dL_dMt.c[1][k] *= s1;
            dL_dMt.c[2][k] *= s2;
        }
        out.rot[0] = 2 * z * (dL_dMt.c[0][1] - dL_dMt.c[1][0]) + 2 * y * (dL_dMt.c[2][0] - dL_dMt.c[0][2]) +
                     2 * x * (dL_dMt.c[1][2] - dL_dMt.c[2][1]);
        out.rot[1] = 2 * y * (dL_dMt.c[1][0] + dL_dMt.c[0][1]) + 2 * z * (dL_dMt.c[2][0] + dL_dMt.c[0][2]) +
                     2 * r * (dL_dMt.c[1][2] - dL_dMt.c[2][1]) - 4 * x * (dL_dMt.c[2][2] + dL_dMt.c[1][1]);
        out.rot[2] = 2 * x * (dL_dMt.c[1][0] + dL_dMt.c[0][1]) + 2 * r * (dL_dMt.c[2][0] - dL_dMt.c[0][2]) +
                     2 * z * (dL_dMt.c[1][2] + dL_dMt.c[2][1]) - 4 * y * (dL_dMt.c[2][2] + dL_dMt.c[0][0]);
        out.rot[3] = 2 * r * (dL_dMt.c[0][1] - dL_dMt.c[1][0]) + 2 * x * (dL_dMt.c[2][0] + dL_dMt.c[0][2]) +
                     2 * y * (dL_dMt.c[1][2] + dL_dMt.c[2][1]) - 4 * z * (dL_dMt.c[1][1] + dL_dMt.c[0][0]);
    }

    out.mean3D[0] = dmean[0];
    out.mean3D[1] = dmean[1];
    out.mean3D[2] = dmean[2];
    pose[PA_SHIFT + 0] = dshift[0];
    pose[PA_SHIFT + 1] = dshift[1];
    pose[PA_SHIFT + 2] = dshift[2];
}

// ACCUM = false: `acc_mask` is ignored (every dense row is written).  ACCUM = true: the outputs selected by acc_mask
// are accumulated (separate instantiation, so the plain path keeps its register count).
// MINB: resident CTAs per SM the register allocation is held to (5 -> 96 registers, 6 -> 80 with 36 bytes of spills).
template <bool ACCUM, int MINB>
__global__ void __launch_bounds__(PB_THREADS, MINB)
preprocess_bwd_kernel(int P, int D, int M, const float* __restrict__ means3D, const int* __restrict__ radii,
                      const float* __restrict__ shs, const float* __restrict__ scales,
                      const float* __restrict__ rotations, float scale_modifier,
                      const float* __restrict__ cov3D_precomp, const float* __restrict__ viewmatrix,
                      const float* __restrict__ projmatrix, const float* __restrict__ campos_ptr, int W, int H,
                      float tan_fovx, float tan_fovy, float focal_x, float focal_y, const float4* __restrict__ rec,
                      const float* __restrict__ gacc, double* __restrict__ pose_acc, BwdOut o, int sh_vec4,
                      u32 acc_mask) {
    const u32 acc = ACCUM ? acc_mask : 0u;
    extern __shared__ float sh_tiles[];           // [PB_WARPS][32][sh_tile_stride(M)]
    __shared__ float view[16];
    __shared__ float proj[16];
    __shared__ float campos[3];
    __shared__ double s_red[PB_WARPS][PA_NRED];
    __shared__ u32 s_win[PB_WARPS];
    if (threadIdx.x < 16) {
        view[threadIdx.x] = viewmatrix[threadIdx.x];
        proj[threadIdx.x] = projmatrix[threadIdx.x];
    }
    if (threadIdx.x < 3) campos[threadIdx.x] = campos_ptr[threadIdx.x];
    __syncthreads();

    const int idx = blockIdx.x * PB_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    float pose[PA_NRED];
#pragma unroll
    for (int k = 0; k < PA_NRED; k++) pose[k] = 0.f;
    u32 winner = 0;

    // ---- stage the SH rows of this warp's visible Gaussians in shared memory (coalesced 128-bit loads) ----------
    const int sh_len = 3 * M;
    const int stride = sh_tile_stride(M);
    float* tile = sh_tiles + (size_t)warp * 32 * stride;
    float* my_sh = tile + lane * stride;          // this lane's row: SH in, dL_dSH out (in place)
    const int warp_first = blockIdx.x * PB_THREADS + warp * 32;
    const int rows = min(32, P - warp_first);
    const bool lane_visible = (idx < P) && (radii[idx] > 0);
    const u32 vis_mask = __ballot_sync(0xffffffffu, lane_visible);
    if (shs != nullptr && rows > 0) {
        float* gsh = const_cast<float*>(shs) + (size_t)warp_first * sh_len;
        if (sh_vec4) sh_tile_copy<4, SH_LOAD>(tile, stride, gsh, sh_len, rows, vis_mask, lane);
        else sh_tile_copy<1, SH_LOAD>(tile, stride, gsh, sh_len, rows, vis_mask, lane);
        __syncwarp();
    }

    if (idx < P) {
        const size_t i = (size_t)idx;
        const bool visible = lane_visible;
        float out_mean3D[3] = {0.f, 0.f, 0.f};
        float out_mean2D[2] = {0.f, 0.f};
        float out_dens[2] = {0.f, 0.f};
        float out_color[3] = {0.f, 0.f, 0.f};
        float out_opacity = 0.f;
        float out_covariance = 0.0001f;   // zeros + 1e-4 (rasterize_points.cu:274)
        float out_cov3D[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float out_scale[3] = {0.f, 0.f, 0.f};
        float out_rot[4] = {0.f, 0.f, 0.f, 0.f};

        if (visible) {
            winner = (u32)idx + 1u;
            GaussGrad out;
            gaussian_backward<false>(idx, D, M, W, H, means3D, shs, scales, rotations, scale_modifier, cov3D_precomp,
                                     view, proj, campos, tan_fovx, tan_fovy, focal_x, focal_y, rec, gacc, my_sh, my_sh,
                                     out, pose);
#pragma unroll
            for (int k = 0; k < 3; k++) { out_mean3D[k] = out.mean3D[k]; out_color[k] = out.color[k]; out_scale[k] = out.scale[k]; }
#pragma unroll
            for (int k = 0; k < 2; k++) { out_mean2D[k] = out.mean2D[k]; out_dens[k] = out.dens[k]; }
#pragma unroll
            for (int k = 0; k < 6; k++) out_cov3D[k] = out.cov3D[k];
#pragma unroll
            for (int k = 0; k < 4; k++) out_rot[k] = out.rot[k];
            out_opacity = out.opacity;
            out_covariance = out.covariance;
        } else if (o.dL_dsh != nullptr && M > 0 && !(acc & BAGS_ACC_SH)) {
            for (int k = 0; k < sh_len; k++) my_sh[k] = 0.f;
        }

        // dense outputs, every row -- except the ones in accumulate mode (include/bags_b200.h, BAGS_ACC_*): those add
        // the rows of visible Gaussians onto the caller's buffer and leave the rows of culled ones alone
        if (!(acc & BAGS_ACC_MEANS3D)) {
#pragma unroll
            for (int k = 0; k < 3; k++) o.dL_dmean3D[3 * i + k] = out_mean3D[k];
        } else if (visible) {
#pragma unroll
            for (int k = 0; k < 3; k++) o.dL_dmean3D[3 * i + k] = __fadd_rn(o.dL_dmean3D[3 * i + k], out_mean3D[k]);
        }
        o.dL_dmean2D[3 * i + 0] = out_mean2D[0];
        o.dL_dmean2D[3 * i + 1] = out_mean2D[1];
        o.dL_dmean2D[3 * i + 2] = 0.f;
        o.dL_dmean2D_densify[3 * i + 0] = out_dens[0];
        o.dL_dmean2D_densify[3 * i + 1] = out_dens[1];
        o.dL_dmean2D_densify[3 * i + 2] = 0.f;
#pragma unroll
        for (int k = 0; k < 3; k++) o.dL_dcolor[3 * i + k] = out_color[k];
        if (!(acc & BAGS_ACC_OPACITY)) o.dL_dopacity[i] = out_opacity;
        else if (visible) o.dL_dopacity[i] = __fadd_rn(o.dL_dopacity[i], out_opacity);
        o.covariance[i] = out_covariance;
#pragma unroll
        for (int k = 0; k < 6; k++) o.dL_dcov3D[6 * i + k] = out_cov3D[k];
        if (!(acc & BAGS_ACC_SCALE)) {
#pragma unroll
            for (int k = 0; k < 3; k++) o.dL_dscale[3 * i + k] = out_scale[k];
        } else if (visible) {
#pragma unroll
            for (int k = 0; k < 3; k++) o.dL_dscale[3 * i + k] = __fadd_rn(o.dL_dscale[3 * i + k], out_scale[k]);
        }
        if (!(acc & BAGS_ACC_ROT)) {
#pragma unroll
            for (int k = 0; k < 4; k++) o.dL_drot[4 * i + k] = out_rot[k];
        } else if (visible) {
#pragma unroll
            for (int k = 0; k < 4; k++) o.dL_drot[4 * i + k] = __fadd_rn(o.dL_drot[4 * i + k], out_rot[k]);
        }
    }

    // ---- dL_dSH: the warp's 32 rows leave as one contiguous, fully coalesced block ------------------------------
    if (o.dL_dsh != nullptr && M > 0 && rows > 0) {
        __syncwarp();
        float* gdst = o.dL_dsh + (size_t)warp_first * sh_len;
        if (acc & BAGS_ACC_SH) {
            if (sh_vec4) sh_tile_copy<4, SH_ACCUM>(tile, stride, gdst, sh_len, rows, vis_mask, lane);
            else sh_tile_copy<1, SH_ACCUM>(tile, stride, gdst, sh_len, rows, vis_mask, lane);
        } else {
            if (sh_vec4) sh_tile_copy<4, SH_STORE>(tile, stride, gdst, sh_len, rows, 0xffffffffu, lane);
            else sh_tile_copy<1, SH_STORE>(tile, stride, gdst, sh_len, rows, 0xffffffffu, lane);
        }
    }

    // ---- pose terms: FP64 warp butterfly -> CTA -> one atomic per term -------------------------------------------
#pragma unroll
    for (int k = 0; k < PA_NRED; k++) {
        double v = (double)pose[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) s_red[warp][k] = v;
    }
    const u32 wwin = __reduce_max_sync(0xffffffffu, winner);
    if (lane == 0) s_win[warp] = wwin;
    __syncthreads();
    if (threadIdx.x < PA_NRED) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < PB_WARPS; w++) v += s_red[w][threadIdx.x];
        if (v != 0.0) atomicAdd(&pose_acc[threadIdx.x], v);
    }
    if (threadIdx.x == 32) {
        u32 m = 0;
#pragma unroll
        for (int w = 0; w < PB_WARPS; w++) m = max(m, s_win[w]);
        if (m) atomicMax(reinterpret_cast<u32*>(pose_acc + PA_WINNER), m);
    }
}

// One small CTA: divide the sums by P and evaluate the view-matrix terms of the winning Gaussian.
// (batched backward: blockIdx.x = view; the per-view inputs / outputs are stacked, the states `*_stride` bytes apart)
__global__ void pose_finalize_kernel(int P, const float* __restrict__ means3D, const float* __restrict__ scales,
                                     const float* __restrict__ rotations, float scale_modifier,
                                     const float* __restrict__ cov3D_precomp, const float* __restrict__ viewmatrix,
                                     const __grid_constant__ ViewFov fov, int W, int H,
                                     const float4* __restrict__ rec, const float* __restrict__ gacc,
                                     const double* __restrict__ pose_acc, BwdOut o, const size_t geom_stride,
                                     const size_t gacc_stride) {
    const unsigned vw = blockIdx.x;
    __shared__ float4 s_fov[MAX_BATCH];
    fov_to_shared(fov, s_fov);
    __syncthreads();
    const float tan_fovx = s_fov[vw].x, tan_fovy = s_fov[vw].y;
    const float focal_x = s_fov[vw].z, focal_y = s_fov[vw].w;
    viewmatrix += 16 * vw;
    rec = at_view(rec, geom_stride, vw);
    gacc = at_view(gacc, gacc_stride, vw);
    pose_acc += POSE_ACC_DOUBLES * vw;
    o.dL_dprojmatrix += 16 * vw;
    o.dL_dviewmatrix += 16 * vw;
    o.dL_dcampos += 3 * vw;
    o.dL_dshift += 3 * vw;
    const int t = threadIdx.x;
    const float invP = (float)P;
    // sum rounded to fp32 once, then the reference's fp32 division by P (torch mean)
    if (t < 16) o.dL_dprojmatrix[t] = (float)pose_acc[PA_PROJ + t] / invP;
    if (t < 3) {
        o.dL_dcampos[t] = (float)pose_acc[PA_CAMPOS + t] / invP;
        o.dL_dshift[t] = (float)pose_acc[PA_SHIFT + t] / invP;
    }
    if (t == 0) {
        float vm[16];
#pragma unroll
        for (int k = 0; k < 16; k++) vm[k] = 0.f;
        const u32 win = *reinterpret_cast<const u32*>(pose_acc + PA_WINNER);
        if (win) {
            const int idx = (int)win - 1;
            const size_t i = (size_t)idx;
            const float3 mean = {means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]};
            float cov6[6];
            load_cov3d(idx, cov3D_precomp, scales, rotations, scale_modifier, cov6);
            const float* ga = gacc + i * GACC_STRIDE;
            const GradScale gs = grad_scale(rec[3 * i + 1], W, H);
            const float3 dL_dconic = {gs.sc * ga[GA_CA], gs.sc * ga[GA_CB], gs.sc * ga[GA_CC]};
            const float dL_ddepth = ga[GA_DEPTH];
            const float* view_matrix = viewmatrix;
            const Cov2DBack cb = cov2d_backward(mean, cov6, focal_x, focal_y, tan_fovx, tan_fovy, view_matrix,
                                                dL_dconic);
            const float h_x = focal_x, h_y = focal_y, tz = cb.tz, tz2 = cb.tz2;
            const float dL_dT00 = cb.dL_dT00, dL_dT01 = cb.dL_dT01, dL_dT02 = cb.dL_dT02;
            const float dL_dT10 = cb.dL_dT10, dL_dT11 = cb.dL_dT11, dL_dT12 = cb.dL_dT12;
            // backward.cu:351-363
            vm[0] = h_x * tz * dL_dT00;
            vm[1] = h_x * tz * dL_dT01;
            vm[4] = h_y * tz * dL_dT10;
            vm[5] = h_y * tz * dL_dT11;
            vm[2] = -(view_matrix[0] * dL_dT00 + view_matrix[1] * dL_dT01) * mean.x * h_x * tz2 -
                    (view_matrix[4] * dL_dT10 + view_matrix[5] * dL_dT11 + view_matrix[6] * dL_dT12) * mean.x * h_y * tz2 +
                    (tz - view_matrix[2] * mean.x) * h_x * tz2 * dL_dT02;
            vm[6] = -(view_matrix[0] * dL_dT00 + view_matrix[1] * dL_dT01 + view_matrix[2] * dL_dT02) * mean.y * h_x * tz2 -
                    (view_matrix[4] * dL_dT10 + view_matrix[5] * dL_dT11) * mean.y * h_y * tz2 +
                    (tz - view_matrix[6] * mean.y) * h_y * tz2 * dL_dT12;
            vm[10] = -((view_matrix[0] * dL_dT00 + view_matrix[1] * dL_dT01 + view_matrix[2] * dL_dT02) * h_x +
                       (view_matrix[4] * dL_dT10 + view_matrix[5] * dL_dT11 + view_matrix[6] * dL_dT12) * h_y) *
                     mean.z * tz2;
            vm[14] = -((view_matrix[0] * dL_dT00 + view_matrix[1] * dL_dT01 + view_matrix[2] * dL_dT02) * h_x +
                       (view_matrix[4] * dL_dT10 + view_matrix[5] * dL_dT11 + view_matrix[6] * dL_dT12) * h_y) *
                     tz2;
            vm[2] += mean.x * dL_ddepth;
            vm[6] += mean.y * dL_ddepth;
            vm[10] += mean.z * dL_ddepth;
            vm[14] += dL_ddepth;
        }
#pragma unroll
        for (int k = 0; k < 16; k++) o.dL_dviewmatrix[k] = vm[k] / invP;
    }
}

static int g_pbwd_minb = 6;   // measured at C3: 157.5 us vs 163.4 us with 5
int set_pbwd_min_blocks(int value) {
    const int prev = g_pbwd_minb;
    if (value == 5 || value == 6) g_pbwd_minb = value;
    return prev;
}

int launch_preprocess_bwd(const BwdArgs& a, const GeomState& g, const float* gacc, double* pose_acc, const BwdOut& o,
                          unsigned accumulate_mask, cudaStream_t st) {
    const int blocks = (a.P + PB_THREADS - 1) / PB_THREADS;
    const size_t smem = (size_t)PB_WARPS * 32 * sh_tile_stride(a.M) * sizeof(float);
    if (smem > 200 * 1024) {
        set_error("preprocess_bwd: %d SH coefficients per Gaussian do not fit the staging tile", a.M);
        return BAGS_E_INVALID;
    }
    auto kern = g_pbwd_minb == 6
                    ? (accumulate_mask ? preprocess_bwd_kernel<true, 6> : preprocess_bwd_kernel<false, 6>)
                    : (accumulate_mask ? preprocess_bwd_kernel<true, 5> : preprocess_bwd_kernel<false, 5>);
    if (smem > 48 * 1024)
        BAGS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int vec4 = ((3 * a.M) % 4 == 0) && (((size_t)a.shs & 15) == 0) && (((size_t)o.dL_dsh & 15) == 0);
    kern<<<blocks, PB_THREADS, smem, st>>>(
        a.P, a.D, a.M, a.means3D, a.radii, a.shs, a.scales, a.rotations, a.scale_modifier, a.cov3D_precomp,
        a.viewmatrix, a.projmatrix, a.campos, a.W, a.H, a.tan_fovx, a.tan_fovy, a.focal_x, a.focal_y, g.rec, gacc,
        pose_acc, o, vec4, (u32)accumulate_mask);
    BAGS_LAUNCH_OK("preprocess_bwd_kernel");
    ViewFov fov;
    fov.tan_fovx[0] = a.tan_fovx; fov.tan_fovy[0] = a.tan_fovy; fov.focal_x[0] = a.focal_x; fov.focal_y[0] = a.focal_y;
    pose_finalize_kernel<<<1, 32, 0, st>>>(a.P, a.means3D, a.scales, a.rotations, a.scale_modifier, a.cov3D_precomp,
                                           a.viewmatrix, fov, a.W, a.H, g.rec, gacc, pose_acc, o, 0, 0);
    BAGS_LAUNCH_OK("pose_finalize_kernel");
    return 0;
}


// =================================================================================================================
// Batched backward preprocess (bags_rasterize_backward_batch): B views of the same Gaussians in one pass.
// Per Gaussian, everything view-independent is touched ONCE: its SH row comes into shared memory once, and the
// gradients w.r.t. its parameters (mean, SH, opacity, scale, rotation, and colour / 3D covariance when those are the
// inputs) are summed over the views in registers / a second shared-memory tile and leave once -- written, or (sinks:
// acc_mask) added onto the caller's buffer, i.e. buffer + g_0 + g_1 + ... in view order with fp32 round-to-nearest adds:
// exactly the sequence of `grad += fresh` the views would perform one after the other.  Per-view results (screen-space
// gradients, pose / intrinsic sums) go to stacked per-view outputs.
// [first, first + count) is the slice of Gaussians this launch covers: the caller may cut the pass into slices and
// start exchanging the finished rows of one slice (multi-GPU gradient all-reduce) while the next one runs.
// =================================================================================================================

template <bool ACCUM>
__global__ void __launch_bounds__(PB_THREADS, 4)
preprocess_bwd_batch_kernel(int P, int first, int count, int D, int M, const float* __restrict__ means3D,
                            const int* __restrict__ radii, const float* __restrict__ shs,
                            const float* __restrict__ scales, const float* __restrict__ rotations,
                            float scale_modifier, const float* __restrict__ cov3D_precomp,
                            const float* __restrict__ viewmatrix, const float* __restrict__ projmatrix,
                            const float* __restrict__ campos_ptr, int W, int H, const __grid_constant__ ViewFov fov, const int B,
                            const float4* __restrict__ rec0, const size_t geom_stride,
                            const float* __restrict__ gacc0, const size_t gacc_stride, double* __restrict__ pose_acc,
                            BwdBatchOut o, int sh_vec4, u32 acc_mask) {
    const u32 acc = ACCUM ? acc_mask : 0u;
    extern __shared__ float sh_tiles[];           // [2][PB_WARPS][32][sh_tile_stride(M)]: SH rows | summed dL_dSH
    __shared__ float views[MAX_BATCH][16];
    __shared__ float projs[MAX_BATCH][16];
    __shared__ float camposs[MAX_BATCH][3];
    __shared__ double s_red[PB_WARPS][PA_NRED];
    __shared__ u32 s_win[PB_WARPS];
    __shared__ float4 s_fov[MAX_BATCH];
    fov_to_shared(fov, s_fov);
    for (int t = threadIdx.x; t < 16 * B; t += PB_THREADS) {
        views[t >> 4][t & 15] = viewmatrix[t];
        projs[t >> 4][t & 15] = projmatrix[t];
    }
    if (threadIdx.x < 3 * B) camposs[threadIdx.x / 3][threadIdx.x % 3] = campos_ptr[threadIdx.x];
    __syncthreads();

    const int end = first + count;
    const int idx = first + blockIdx.x * PB_THREADS + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool valid = idx < end;
    const size_t i = (size_t)idx;

    const int sh_len = 3 * M;
    const int stride = sh_tile_stride(M);
    float* tile_in = sh_tiles + (size_t)warp * 32 * stride;
    float* tile_acc = sh_tiles + (size_t)(PB_WARPS + warp) * 32 * stride;
    const int warp_first = first + blockIdx.x * PB_THREADS + warp * 32;
    const int rows = max(0, min(32, end - warp_first));
    const bool has_sh = shs != nullptr && o.dL_dsh != nullptr && M > 0;

    // which views see this Gaussian
    u32 visbits = 0;
    if (valid)
        for (int v = 0; v < B; v++) visbits |= (radii[(size_t)v * P + i] > 0 ? 1u : 0u) << v;
    const bool any = visbits != 0u;
    const u32 any_mask = __ballot_sync(0xffffffffu, any);

    if (has_sh && rows > 0) {
        float* gsh = const_cast<float*>(shs) + (size_t)warp_first * sh_len;
        float* gdst = o.dL_dsh + (size_t)warp_first * sh_len;
        if (sh_vec4) sh_tile_copy<4, SH_LOAD>(tile_in, stride, gsh, sh_len, rows, any_mask, lane);
        else sh_tile_copy<1, SH_LOAD>(tile_in, stride, gsh, sh_len, rows, any_mask, lane);
        if (acc & BAGS_ACC_SH) {
            if (sh_vec4) sh_tile_copy<4, SH_LOAD_RW>(tile_acc, stride, gdst, sh_len, rows, any_mask, lane);
            else sh_tile_copy<1, SH_LOAD_RW>(tile_acc, stride, gdst, sh_len, rows, any_mask, lane);
        } else {
            for (int f = lane; f < 32 * stride; f += 32) tile_acc[f] = 0.f;
        }
        __syncwarp();
    }

    // sums over the views (sinks: starting from what the caller's buffer holds)
    float s_mean3D[3] = {0.f, 0.f, 0.f}, s_color[3] = {0.f, 0.f, 0.f}, s_opacity = 0.f, s_scale[3] = {0.f, 0.f, 0.f};
    float s_rot[4] = {0.f, 0.f, 0.f, 0.f}, s_cov3D[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (ACCUM && any) {
        if (acc & BAGS_ACC_MEANS3D)
#pragma unroll
            for (int k = 0; k < 3; k++) s_mean3D[k] = o.dL_dmean3D[3 * i + k];
        if (acc & BAGS_ACC_OPACITY) s_opacity = o.dL_dopacity[i];
        if (acc & BAGS_ACC_SCALE)
#pragma unroll
            for (int k = 0; k < 3; k++) s_scale[k] = o.dL_dscale[3 * i + k];
        if (acc & BAGS_ACC_ROT)
#pragma unroll
            for (int k = 0; k < 4; k++) s_rot[k] = o.dL_drot[4 * i + k];
    }

    for (int v = 0; v < B; v++) {
        float pose[PA_NRED];
#pragma unroll
        for (int k = 0; k < PA_NRED; k++) pose[k] = 0.f;
        u32 winner = 0;
        if (valid) {
            const bool visible = (visbits >> v) & 1u;
            float m2d[2] = {0.f, 0.f}, dens[2] = {0.f, 0.f}, covar = 0.0001f;
            if (visible) {
                winner = (u32)idx + 1u;
                GaussGrad g;
                gaussian_backward<true>(idx, D, M, W, H, means3D, has_sh ? shs : nullptr, scales, rotations,
                                        scale_modifier, cov3D_precomp, views[v], projs[v], camposs[v], s_fov[v].x,
                                        s_fov[v].y, s_fov[v].z, s_fov[v].w,
                                        at_view(rec0, geom_stride, v), at_view(gacc0, gacc_stride, v),
                                        tile_in + lane * stride, tile_acc + lane * stride, g, pose);
                m2d[0] = g.mean2D[0]; m2d[1] = g.mean2D[1];
                dens[0] = g.dens[0]; dens[1] = g.dens[1];
                covar = g.covariance;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    s_mean3D[k] = __fadd_rn(s_mean3D[k], g.mean3D[k]);
                    s_color[k] = __fadd_rn(s_color[k], g.color[k]);
                    s_scale[k] = __fadd_rn(s_scale[k], g.scale[k]);
                }
#pragma unroll
                for (int k = 0; k < 4; k++) s_rot[k] = __fadd_rn(s_rot[k], g.rot[k]);
#pragma unroll
                for (int k = 0; k < 6; k++) s_cov3D[k] = __fadd_rn(s_cov3D[k], g.cov3D[k]);
                s_opacity = __fadd_rn(s_opacity, g.opacity);
            }
            const size_t ov = (size_t)v * P + i;
            o.dL_dmean2D[3 * ov + 0] = m2d[0];
            o.dL_dmean2D[3 * ov + 1] = m2d[1];
            o.dL_dmean2D[3 * ov + 2] = 0.f;
            o.dL_dmean2D_densify[3 * ov + 0] = dens[0];
            o.dL_dmean2D_densify[3 * ov + 1] = dens[1];
            o.dL_dmean2D_densify[3 * ov + 2] = 0.f;
            if (o.covariance) o.covariance[ov] = covar;
        }
        // pose terms of view v: FP64 warp butterfly -> CTA -> one atomic per term
#pragma unroll
        for (int k = 0; k < PA_NRED; k++) {
            double x = (double)pose[k];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
            if (lane == 0) s_red[warp][k] = x;
        }
        const u32 wwin = __reduce_max_sync(0xffffffffu, winner);
        if (lane == 0) s_win[warp] = wwin;
        __syncthreads();
        double* pa = pose_acc + (size_t)v * POSE_ACC_DOUBLES;
        if (threadIdx.x < PA_NRED) {
            double x = 0.0;
#pragma unroll
            for (int w = 0; w < PB_WARPS; w++) x += s_red[w][threadIdx.x];
            if (x != 0.0) atomicAdd(&pa[threadIdx.x], x);
        }
        if (threadIdx.x == 32) {
            u32 m = 0;
#pragma unroll
            for (int w = 0; w < PB_WARPS; w++) m = max(m, s_win[w]);
            if (m) atomicMax(reinterpret_cast<u32*>(pa + PA_WINNER), m);
        }
        __syncthreads();
    }

    if (valid) {
        // sinks: only the rows some view saw are rewritten (the others keep what the buffer holds)
        if (!(acc & BAGS_ACC_MEANS3D) || any)
#pragma unroll
            for (int k = 0; k < 3; k++) o.dL_dmean3D[3 * i + k] = s_mean3D[k];
        if (!(acc & BAGS_ACC_OPACITY) || any) o.dL_dopacity[i] = s_opacity;
        if (!(acc & BAGS_ACC_SCALE) || any)
#pragma unroll
            for (int k = 0; k < 3; k++) o.dL_dscale[3 * i + k] = s_scale[k];
        if (!(acc & BAGS_ACC_ROT) || any)
#pragma unroll
            for (int k = 0; k < 4; k++) o.dL_drot[4 * i + k] = s_rot[k];
        if (o.dL_dcolor)
#pragma unroll
            for (int k = 0; k < 3; k++) o.dL_dcolor[3 * i + k] = s_color[k];
        if (o.dL_dcov3D)
#pragma unroll
            for (int k = 0; k < 6; k++) o.dL_dcov3D[6 * i + k] = s_cov3D[k];
    }
    if (has_sh && rows > 0) {
        __syncwarp();
        float* gdst = o.dL_dsh + (size_t)warp_first * sh_len;
        if (acc & BAGS_ACC_SH) {
            if (sh_vec4) sh_tile_copy<4, SH_STORE_MASKED>(tile_acc, stride, gdst, sh_len, rows, any_mask, lane);
            else sh_tile_copy<1, SH_STORE_MASKED>(tile_acc, stride, gdst, sh_len, rows, any_mask, lane);
        } else {
            if (sh_vec4) sh_tile_copy<4, SH_STORE>(tile_acc, stride, gdst, sh_len, rows, 0xffffffffu, lane);
            else sh_tile_copy<1, SH_STORE>(tile_acc, stride, gdst, sh_len, rows, 0xffffffffu, lane);
        }
    }
}

// Slice [first, first + count) of the batched backward preprocess; `finalize`: this is the last slice (the per-view pose /
// intrinsic gradients are completed).  a.viewmatrix / projmatrix / campos: [B][16 | 16 | 3]; a.radii: [B][P]; pose_acc:
// [B][POSE_ACC_DOUBLES] zeroed before the first slice; pose outputs of `po`: [B][16 | 16 | 3 | 3].
int launch_preprocess_bwd_batch(const BwdArgs& a, int B, const ViewFov& fov, const GeomState& g, size_t geom_stride,
                                const float* gacc, size_t gacc_stride, double* pose_acc, const BwdOut& po,
                                const BwdBatchOut& o, unsigned accumulate_mask, int first, int count, bool finalize,
                                cudaStream_t st) {
    if (B < 1 || B > MAX_BATCH) {
        set_error("preprocess_bwd: %d views in one launch (1..%d supported)", B, MAX_BATCH);
        return BAGS_E_INVALID;
    }
    if (count > 0) {
        const int blocks = (count + PB_THREADS - 1) / PB_THREADS;
        const size_t smem = (size_t)2 * PB_WARPS * 32 * sh_tile_stride(a.M) * sizeof(float);
        if (smem > 200 * 1024) {
            set_error("preprocess_bwd: %d SH coefficients per Gaussian do not fit the staging tiles", a.M);
            return BAGS_E_INVALID;
        }
        auto kern = accumulate_mask ? preprocess_bwd_batch_kernel<true> : preprocess_bwd_batch_kernel<false>;
        if (smem > 48 * 1024)
            BAGS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // slices start at arbitrary rows: the vector path needs 16-byte aligned rows, i.e. 3M a multiple of 4
        const int vec4 = ((3 * a.M) % 4 == 0) && (((size_t)a.shs & 15) == 0) && (((size_t)o.dL_dsh & 15) == 0);
        kern<<<blocks, PB_THREADS, smem, st>>>(a.P, first, count, a.D, a.M, a.means3D, a.radii, a.shs, a.scales,
                                               a.rotations, a.scale_modifier, a.cov3D_precomp, a.viewmatrix,
                                               a.projmatrix, a.campos, a.W, a.H, fov, B, g.rec, geom_stride, gacc,
                                               gacc_stride, pose_acc, o, vec4, (u32)accumulate_mask);
        BAGS_LAUNCH_OK("preprocess_bwd_batch_kernel");
    }
    if (finalize) {
        pose_finalize_kernel<<<B, 32, 0, st>>>(a.P, a.means3D, a.scales, a.rotations, a.scale_modifier, a.cov3D_precomp,
                                               a.viewmatrix, fov, a.W, a.H, g.rec, gacc, pose_acc, po, geom_stride,
                                               gacc_stride);
        BAGS_LAUNCH_OK("pose_finalize_kernel");
    }
    return 0;
}

}  // namespace bags
