// distCUDA2 replacement for sm_100a: mean squared distance to the 3 nearest neighbours of every point.
//
// Semantics follow simple-knn/simple_knn.cu:132-221 of the reference exactly: for every point, the three smallest
// values of  dist = dx*dx + dy*dy + dz*dz  (fp32, same expression => same fma contraction) over all OTHER points
// (self excluded by position, coincident points count with distance 0), result (b0 + b1 + b2) / 3.0f written at the
// point's original index.  The reference reaches that result by Morton-sorting the points, boxing every 1024
// consecutive points and letting each point loop over ALL boxes (O(P^2/1024) box tests: 3.4e10 at 6 M points) with
// two blocking device->host copies for the bounding box and five thrust allocations.  Because the result is the
// exact 3-NN multiset it does not depend on the search structure, so a different one is used here:
//
//   bbox (atomic float min/max, stays on the device) -> 30-bit Morton codes -> the hand-written onesweep sort
//   (binning.cu) -> points gathered into Morton order as float4 {x,y,z,id} -> a 32-ary implicit box hierarchy
//   (32 points / 1 Ki / 32 Ki points per box) -> per-point pruned descent, seeded with the point's own 32-point leaf.
//
// A box is skipped only if its fp32 box distance exceeds the current third-best distance, and the box distance is a
// lower bound of every fp32 point distance inside (each operation is monotone), so the search is exact.
#include "../../include/bags_b200.h"
#include "common.cuh"

#include <float.h>

namespace bags {

struct Box {
    float3 lo, hi;
};

constexpr int KNN_THREADS = 256;
constexpr int FAN = 32;

struct KnnScratch {
    float* bbox;      // [8]: min xyz, max xyz
    u32* code_a;      // [P]
    u32* code_b;      // [P]
    u32* idx_a;       // [P]
    u32* idx_b;       // [P]
    float4* pts;      // [P] Morton-ordered {x,y,z,bits(original index)}
    Box* l0;          // [n0] 32 points each
    Box* l1;          // [n1] 32 leaves each
    Box* l2;          // [n2] 32 l1 boxes each
    SortScratch sort;
    int n0, n1, n2;
    static size_t bytes(int P) {
        size_t n = (size_t)(P > 0 ? P : 1);
        size_t n0 = (n + FAN - 1) / FAN, n1 = (n0 + FAN - 1) / FAN, n2 = (n1 + FAN - 1) / FAN;
        return align_up(8 * sizeof(float)) + 4 * align_up(n * sizeof(u32)) + align_up(n * sizeof(float4)) +
               align_up(n0 * sizeof(Box)) + align_up(n1 * sizeof(Box)) + align_up(n2 * sizeof(Box)) +
               SortScratch::bytes(P) + 256;
    }
    static KnnScratch from(char* p, int P) {
        KnnScratch s;
        p = (char*)align_up((size_t)p);
        size_t n = (size_t)(P > 0 ? P : 1);
        s.n0 = (int)((n + FAN - 1) / FAN);
        s.n1 = (s.n0 + FAN - 1) / FAN;
        s.n2 = (s.n1 + FAN - 1) / FAN;
        carve(p, s.bbox, 8);
        carve(p, s.code_a, n);
        carve(p, s.code_b, n);
        carve(p, s.idx_a, n);
        carve(p, s.idx_b, n);
        carve(p, s.pts, n);
        carve(p, s.l0, (size_t)s.n0);
        carve(p, s.l1, (size_t)s.n1);
        carve(p, s.l2, (size_t)s.n2);
        s.sort = SortScratch::from(p, P);
        return s;
    }
};

__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
    if (v >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void knn_bbox_init_kernel(float* bbox) {
    // the reference reduces with init {0,0,0} (simple_knn.cu:191-200): the box always contains the origin
    if (threadIdx.x < 8) bbox[threadIdx.x] = 0.f;
}

__global__ void __launch_bounds__(KNN_THREADS)
knn_bbox_kernel(int P, const float* __restrict__ pts, float* __restrict__ bbox) {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float v = pts[3 * (size_t)i + k];
            lo[k] = fminf(lo[k], v);
            hi[k] = fmaxf(hi[k], v);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (lo[k] <= hi[k]) {
                atomic_min_float(&bbox[k], lo[k]);
                atomic_max_float(&bbox[4 + k], hi[k]);
            }
        }
    }
}

__device__ __forceinline__ u32 spread10(u32 x) {
    x = (x | (x << 16)) & 0x030000FF;
    x = (x | (x << 8)) & 0x0300F00F;
    x = (x | (x << 4)) & 0x030C30C3;
    x = (x | (x << 2)) & 0x09249249;
    return x;
}

__global__ void knn_morton_kernel(int P, const float* __restrict__ pts, const float* __restrict__ bbox,
                                  u32* __restrict__ codes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float3 lo = {bbox[0], bbox[1], bbox[2]};
    const float3 hi = {bbox[4], bbox[5], bbox[6]};
    const float x = pts[3 * (size_t)i], y = pts[3 * (size_t)i + 1], z = pts[3 * (size_t)i + 2];
    const u32 cx = spread10((u32)(((x - lo.x) / (hi.x - lo.x)) * ((1 << 10) - 1)));
    const u32 cy = spread10((u32)(((y - lo.y) / (hi.y - lo.y)) * ((1 << 10) - 1)));
    const u32 cz = spread10((u32)(((z - lo.z) / (hi.z - lo.z)) * ((1 << 10) - 1)));
    codes[i] = cx | (cy << 1) | (cz << 2);
}

// Gather into Morton order and build the 32-point leaf boxes (one warp per leaf).
__global__ void __launch_bounds__(KNN_THREADS)
knn_gather_kernel(int P, const float* __restrict__ pts, const u32* __restrict__ order, float4* __restrict__ sorted,
                  Box* __restrict__ l0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float3 lo = {FLT_MAX, FLT_MAX, FLT_MAX}, hi = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (i < P) {
        const u32 id = order[i];
        const float x = pts[3 * (size_t)id], y = pts[3 * (size_t)id + 1], z = pts[3 * (size_t)id + 2];
        sorted[i] = make_float4(x, y, z, __uint_as_float(id));
        lo = {x, y, z};
        hi = {x, y, z};
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo.x = fminf(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, o));
        lo.y = fminf(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, o));
        lo.z = fminf(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, o));
        hi.x = fmaxf(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, o));
        hi.y = fmaxf(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, o));
        hi.z = fmaxf(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, o));
    }
    const int leaf = i >> 5;
    if ((threadIdx.x & 31) == 0 && leaf * FAN < P) {
        Box b;
        b.lo = lo;
        b.hi = hi;
        l0[leaf] = b;
    }
}

// Parent boxes: one warp merges 32 children.
__global__ void __launch_bounds__(KNN_THREADS)
knn_merge_kernel(int n_child, const Box* __restrict__ child, Box* __restrict__ parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float3 lo = {FLT_MAX, FLT_MAX, FLT_MAX}, hi = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (i < n_child) {
        lo = child[i].lo;
        hi = child[i].hi;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo.x = fminf(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, o));
        lo.y = fminf(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, o));
        lo.z = fminf(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, o));
        hi.x = fmaxf(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, o));
        hi.y = fmaxf(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, o));
        hi.z = fmaxf(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, o));
    }
    const int p = i >> 5;
    if ((threadIdx.x & 31) == 0 && p * FAN < n_child) {
        Box b;
        b.lo = lo;
        b.hi = hi;
        parent[p] = b;
    }
}

// simple_knn.cu:119-130
__device__ __forceinline__ float box_dist2(const Box& box, const float3& p) {
    float3 diff = {0, 0, 0};
    if (p.x < box.lo.x || p.x > box.hi.x) diff.x = min(abs(p.x - box.lo.x), abs(p.x - box.hi.x));
    if (p.y < box.lo.y || p.y > box.hi.y) diff.y = min(abs(p.y - box.lo.y), abs(p.y - box.hi.y));
    if (p.z < box.lo.z || p.z > box.hi.z) diff.z = min(abs(p.z - box.lo.z), abs(p.z - box.hi.z));
    return diff.x * diff.x + diff.y * diff.y + diff.z * diff.z;
}

// simple_knn.cu:132-146
__device__ __forceinline__ void update3(const float3& ref, const float4& q, float* knn) {
    float3 d = {q.x - ref.x, q.y - ref.y, q.z - ref.z};
    float dist = d.x * d.x + d.y * d.y + d.z * d.z;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        if (knn[j] > dist) {
            float t = knn[j];
            knn[j] = dist;
            dist = t;
        }
    }
}

__global__ void __launch_bounds__(KNN_THREADS)
knn_query_kernel(int P, const float4* __restrict__ pts, const Box* __restrict__ l0, const Box* __restrict__ l1,
                 const Box* __restrict__ l2, int n0, int n1, int n2, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const float4 me = pts[i];
    const float3 p = {me.x, me.y, me.z};
    float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};

    // seed with the point's own leaf (same 32 points for the whole warp)
    const int my_leaf = i >> 5;
    {
        const int b = my_leaf * FAN, e = min(P, b + FAN);
        for (int k = b; k < e; k++) {
            if (k == i) continue;
            update3(p, pts[k], best);
        }
    }
    // pruned descent through the 32-ary hierarchy
    for (int c2 = 0; c2 < n2; c2++) {
        if (box_dist2(l2[c2], p) > best[2]) continue;
        const int e1 = min(n1, (c2 + 1) * FAN);
        for (int c1 = c2 * FAN; c1 < e1; c1++) {
            if (box_dist2(l1[c1], p) > best[2]) continue;
            const int e0 = min(n0, (c1 + 1) * FAN);
            for (int c0 = c1 * FAN; c0 < e0; c0++) {
                if (c0 == my_leaf) continue;
                if (box_dist2(l0[c0], p) > best[2]) continue;
                const int b = c0 * FAN, e = min(P, b + FAN);
                for (int k = b; k < e; k++) update3(p, pts[k], best);
            }
        }
    }
    out[__float_as_uint(me.w)] = (best[0] + best[1] + best[2]) / 3.0f;
}

}  // namespace bags

using namespace bags;

extern "C" size_t bags_knn_scratch_bytes(int P) { return KnnScratch::bytes(P); }

extern "C" int bags_knn_mean_dist2(int P, const float* points, float* mean_dist2, char* scratch, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (P < 0 || (P > 0 && (!points || !mean_dist2 || !scratch))) {
        set_error("bags_knn_mean_dist2: invalid argument");
        return BAGS_E_INVALID;
    }
    if (P == 0) return 0;
    if (P >= (1 << 30)) {
        set_error("bags_knn_mean_dist2: P=%d exceeds the 30-bit sort bookkeeping", P);
        return BAGS_E_OVERFLOW;
    }
    KnnScratch s = KnnScratch::from(scratch, P);
    knn_bbox_init_kernel<<<1, 32, 0, st>>>(s.bbox);
    BAGS_LAUNCH_OK("knn_bbox_init_kernel");
    const int rb = min((P + KNN_THREADS - 1) / KNN_THREADS, 148 * 8);
    knn_bbox_kernel<<<rb, KNN_THREADS, 0, st>>>(P, points, s.bbox);
    BAGS_LAUNCH_OK("knn_bbox_kernel");
    const int pb = (P + KNN_THREADS - 1) / KNN_THREADS;
    knn_morton_kernel<<<pb, KNN_THREADS, 0, st>>>(P, points, s.bbox, s.code_a);
    BAGS_LAUNCH_OK("knn_morton_kernel");
    u32 *fk, *order;
    int rc = launch_sort_pairs(P, nullptr, s.code_a, s.idx_a, s.code_b, s.idx_b, 0, 30, s.sort, true, SORT_FRESH, st, &fk, &order);
    if (rc) return rc;
    knn_gather_kernel<<<pb, KNN_THREADS, 0, st>>>(P, points, order, s.pts, s.l0);
    BAGS_LAUNCH_OK("knn_gather_kernel");
    knn_merge_kernel<<<(s.n0 + KNN_THREADS - 1) / KNN_THREADS, KNN_THREADS, 0, st>>>(s.n0, s.l0, s.l1);
    BAGS_LAUNCH_OK("knn_merge_kernel");
    knn_merge_kernel<<<(s.n1 + KNN_THREADS - 1) / KNN_THREADS, KNN_THREADS, 0, st>>>(s.n1, s.l1, s.l2);
    BAGS_LAUNCH_OK("knn_merge_kernel");
    knn_query_kernel<<<pb, KNN_THREADS, 0, st>>>(P, s.pts, s.l0, s.l1, s.l2, s.n0, s.n1, s.n2, mean_dist2);
    BAGS_LAUNCH_OK("knn_query_kernel");
    return 0;
}
