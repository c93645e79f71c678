// Small column-major 3x3 / 4x4 helpers used by the per-Gaussian kernels.
//
// Parity note: radii, tile rectangles and depth keys must match the reference bit for bit, and those bits are decided
// by the *order* of fp32 operations (and by where nvcc contracts mul+add into fma).  The reference evaluates its
// covariance algebra through generic column-major matrix products -- including the multiplications by the literal
// zeros of S and J (forward.cu:115-156,172-194 of the reference).  These helpers therefore keep the same textbook
// product formula  R[j][i] = A[0][i]*B[j][0] + A[1][i]*B[j][1] + A[2][i]*B[j][2]  (summed left to right) instead of
// a hand-simplified form, so that the compiler sees the same expression trees.  The kernels that use them are
// bandwidth-bound, so the redundant flops are free.
#pragma once
#include "common.cuh"

namespace bags {

struct M3 {
    float c[3][3];  // c[col][row]
};

__device__ __forceinline__ M3 m3_cols(float a0, float a1, float a2, float b0, float b1, float b2, float c0, float c1,
                                      float c2) {
    M3 m;
    m.c[0][0] = a0; m.c[0][1] = a1; m.c[0][2] = a2;
    m.c[1][0] = b0; m.c[1][1] = b1; m.c[1][2] = b2;
    m.c[2][0] = c0; m.c[2][1] = c1; m.c[2][2] = c2;
    return m;
}

__device__ __forceinline__ M3 m3_mul(const M3& a, const M3& b) {
    M3 r;
#pragma unroll
    for (int j = 0; j < 3; j++) {
#pragma unroll
        for (int i = 0; i < 3; i++) {
            r.c[j][i] = a.c[0][i] * b.c[j][0] + a.c[1][i] * b.c[j][1] + a.c[2][i] * b.c[j][2];
        }
    }
    return r;
}

__device__ __forceinline__ M3 m3_transpose(const M3& a) {
    M3 r;
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++) r.c[j][i] = a.c[i][j];
    return r;
}

__device__ __forceinline__ M3 m3_scale(float s, const M3& a) {
    M3 r;
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
        for (int i = 0; i < 3; i++) r.c[j][i] = a.c[j][i] * s;
    return r;
}

// p' = M p with the reference's flat layout m[4*c + r] (auxiliary.h:58-77).
__device__ __forceinline__ float3 xform4x3(const float3& p, const float* m) {
    float3 t = {
        m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
        m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
        m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14],
    };
    return t;
}

__device__ __forceinline__ float4 xform4x4(const float3& p, const float* m) {
    float4 t = {
        m[0] * p.x + m[4] * p.y + m[8] * p.z + m[12],
        m[1] * p.x + m[5] * p.y + m[9] * p.z + m[13],
        m[2] * p.x + m[6] * p.y + m[10] * p.z + m[14],
        m[3] * p.x + m[7] * p.y + m[11] * p.z + m[15],
    };
    return t;
}

__device__ __forceinline__ float3 xformVec4x3T(const float3& p, const float* m) {
    float3 t = {
        m[0] * p.x + m[1] * p.y + m[2] * p.z,
        m[4] * p.x + m[5] * p.y + m[6] * p.z,
        m[8] * p.x + m[9] * p.y + m[10] * p.z,
    };
    return t;
}

// Inverse of the 4x4 whose column c is (v[c], v[4+c], v[8+c], v[12+c]) -- i.e. the matrix the reference builds
// from `viewmatrix` before inverting it per Gaussian (forward.cu:26-31).  Cofactor expansion, fp32.
// out[c][r] column-major.
__device__ inline void inverse_w2c(const float* v, float out[4][4]) {
    float m[4][4];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        m[c][0] = v[c];
        m[c][1] = v[4 + c];
        m[c][2] = v[8 + c];
        m[c][3] = v[12 + c];
    }
    float c00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
    float c02 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
    float c03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
    float c04 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
    float c06 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
    float c07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
    float c08 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
    float c10 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
    float c11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
    float c12 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
    float c14 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
    float c15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
    float c16 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
    float c18 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
    float c19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
    float c20 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
    float c22 = m[1][0] * m[3][1] - m[3][0] * m[1][1];
    float c23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];

    float f0[4] = {c00, c00, c02, c03};
    float f1[4] = {c04, c04, c06, c07};
    float f2[4] = {c08, c08, c10, c11};
    float f3[4] = {c12, c12, c14, c15};
    float f4[4] = {c16, c16, c18, c19};
    float f5[4] = {c20, c20, c22, c23};
    float v0[4] = {m[1][0], m[0][0], m[0][0], m[0][0]};
    float v1[4] = {m[1][1], m[0][1], m[0][1], m[0][1]};
    float v2[4] = {m[1][2], m[0][2], m[0][2], m[0][2]};
    float v3[4] = {m[1][3], m[0][3], m[0][3], m[0][3]};
    const float sa[4] = {+1.f, -1.f, +1.f, -1.f};
    const float sb[4] = {-1.f, +1.f, -1.f, +1.f};
    float inv[4][4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        inv[0][k] = (v1[k] * f0[k] - v2[k] * f1[k] + v3[k] * f2[k]) * sa[k];
        inv[1][k] = (v0[k] * f0[k] - v2[k] * f3[k] + v3[k] * f4[k]) * sb[k];
        inv[2][k] = (v0[k] * f1[k] - v1[k] * f3[k] + v3[k] * f5[k]) * sa[k];
        inv[3][k] = (v0[k] * f2[k] - v1[k] * f4[k] + v2[k] * f5[k]) * sb[k];
    }
    float d0 = m[0][0] * inv[0][0], d1 = m[0][1] * inv[1][0], d2 = m[0][2] * inv[2][0], d3 = m[0][3] * inv[3][0];
    float det = (d0 + d1) + (d2 + d3);
    float ood = 1.0f / det;
#pragma unroll
    for (int c = 0; c < 4; c++)
#pragma unroll
        for (int r = 0; r < 4; r++) out[c][r] = inv[c][r] * ood;
}

// SH basis constants (public values of the real spherical-harmonics basis, bands 0..3).
#define BAGS_SH_C0 0.28209479177387814f
#define BAGS_SH_C1 0.4886025119029199f
#define BAGS_SH_C2_0 1.0925484305920792f
#define BAGS_SH_C2_1 -1.0925484305920792f
#define BAGS_SH_C2_2 0.31539156525252005f
#define BAGS_SH_C2_3 -1.0925484305920792f
#define BAGS_SH_C2_4 0.5462742152960396f
#define BAGS_SH_C3_0 -0.5900435899266435f
#define BAGS_SH_C3_1 2.890611442640554f
#define BAGS_SH_C3_2 -0.4570457994644658f
#define BAGS_SH_C3_3 0.3731763325901154f
#define BAGS_SH_C3_4 -0.4570457994644658f
#define BAGS_SH_C3_5 1.445305721320277f
#define BAGS_SH_C3_6 -0.5900435899266435f

// 3D covariance from scale + (un-normalised, forward.cu:178) quaternion.  Sigma = (S R)^T (S R); upper triangle out.
__device__ __forceinline__ void cov3d_from_scale_rot(float sx, float sy, float sz, float mod, float qr, float qx,
                                                     float qy, float qz, float* cov6) {
    M3 S = m3_cols(1.0f, 0.0f, 0.0f, 0.0f, 1.0f, 0.0f, 0.0f, 0.0f, 1.0f);
    S.c[0][0] = mod * sx;
    S.c[1][1] = mod * sy;
    S.c[2][2] = mod * sz;
    float r = qr, x = qx, y = qy, z = qz;
    M3 R = m3_cols(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                   2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                   2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
    M3 Mm = m3_mul(S, R);
    M3 Sig = m3_mul(m3_transpose(Mm), Mm);
    cov6[0] = Sig.c[0][0];
    cov6[1] = Sig.c[0][1];
    cov6[2] = Sig.c[0][2];
    cov6[3] = Sig.c[1][1];
    cov6[4] = Sig.c[1][2];
    cov6[5] = Sig.c[2][2];
}

// ((v + 1) * S - 1) / 2 evaluated in double exactly like auxiliary.h:41-44 (the literals there are doubles).
__device__ __forceinline__ float ndc_to_pix(float v, int S) { return ((v + 1.0) * S - 1.0) * 0.5; }

// Tile rectangle of a splat (auxiliary.h:46-56): truncate toward zero, clamp to the grid.
__device__ __forceinline__ void tile_rect(float px, float py, int max_radius, int gx, int gy, uint2& rmin,
                                          uint2& rmax) {
    rmin.x = min((u32)gx, (u32)max((int)0, (int)((px - max_radius) / TILE_X)));
    rmin.y = min((u32)gy, (u32)max((int)0, (int)((py - max_radius) / TILE_Y)));
    rmax.x = min((u32)gx, (u32)max((int)0, (int)((px + max_radius + TILE_X - 1) / TILE_X)));
    rmax.y = min((u32)gy, (u32)max((int)0, (int)((py + max_radius + TILE_Y - 1) / TILE_Y)));
}

}  // namespace bags
