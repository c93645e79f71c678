// Fused densification compaction for sm_100a: clone / split / prune of
// scene/gaussian_model.py:334-447 as ONE gather over all parameter tensors and their Adam state.
//
// The reference's densify_and_prune (gaussian_model.py:433-447) is three passes of whole-model tensor surgery:
//   densify_and_clone  -> densification_postfix: torch.cat of every parameter and of both Adam moments (:354-373,:375-393)
//   densify_and_split  -> the same cat, then prune_points of the split parents: boolean indexing of every parameter and
//                         both moments (:315-348)
//   prune_points       -> boolean indexing of everything once more
// i.e. for each of the 6 parameter groups x {param, exp_avg, exp_avg_sq}: two concatenations + two masked copies of the
// full tensor (each with its own nonzero / host synchronisation), ~70 full-size passes per densification.
//
// The order of the surviving rows is fully determined by three per-row decisions taken on the ORIGINAL P rows:
//   keep  (not a split parent, not pruned) | clone (selected by densify_and_clone, copy not pruned)
//   child (selected by densify_and_split, children not pruned; N children per parent)
// final layout = [kept originals (ascending row) | clones (ascending) | children r = 0 (ascending) | ... | r = N-1].
// The caller evaluates the masks with the reference's own torch expressions (they are O(P) element-wise work) and
// exclusive-scans them; this kernel then moves every row of every tensor exactly once:
//   kept row    : parameter + exp_avg + exp_avg_sq copied
//   clone row   : parameter copied, moments zero                       (torch.cat(..., zeros_like(extension)))
//   child row r : parameter copied, except
//                   xyz      <- R(q / |q|) . sample[r * n_split + rank] + xyz        (gaussian_model.py:407-409)
//                   scaling  <- log(exp(scaling) * (1 / (0.8 N)))                    (:410; torch divides by the Python
//                               scalar 0.8 N as a multiplication by its fp32 reciprocal)
//                 moments zero
// Gathers are bit-exact; the two computed child fields agree with torch's bmm / log-exp to fp32 rounding.
#include "../../include/bags_b200.h"
#include "common.cuh"

namespace bags {

constexpr int DENSIFY_MAX_TENSORS = 8;
constexpr int DENSIFY_THREADS = 256;

struct DensifyBatch {
    const float* src[DENSIFY_MAX_TENSORS];
    float* dst[DENSIFY_MAX_TENSORS];
    const float* src_m[DENSIFY_MAX_TENSORS];
    float* dst_m[DENSIFY_MAX_TENSORS];
    const float* src_v[DENSIFY_MAX_TENSORS];
    float* dst_v[DENSIFY_MAX_TENSORS];
    int width[DENSIFY_MAX_TENSORS];
    int role[DENSIFY_MAX_TENSORS];
    unsigned int block_start[DENSIFY_MAX_TENSORS + 1];
    int count;
};

__device__ __forceinline__ void put(float* dst, float* dst_m, float* dst_v, const long long o, const float val,
                                    const float m, const float v) {
    dst[o] = val;
    if (dst_m) dst_m[o] = m;
    if (dst_v) dst_v[o] = v;
}

// plan: [P][4] int32 = exclusive scans of {keep, clone, child, split-parent}; flags: bit 0 keep, bit 1 clone, bit 2
// child, bit 3 split parent.  totals = {kept, clones, children per repeat, split parents}.
__global__ void __launch_bounds__(DENSIFY_THREADS)
densify_gather_kernel(const DensifyBatch b, const long long P, const int N, const int4* __restrict__ plan,
                      const uint8_t* __restrict__ flags, const int T0, const int T1, const int T2, const int n_split,
                      const float* __restrict__ rotation, const float* __restrict__ samples, const float inv_split) {
    int t = 0;
#pragma unroll
    for (int k = 1; k < DENSIFY_MAX_TENSORS; k++)
        if (k < b.count && blockIdx.x >= b.block_start[k]) t = k;
    const int w = b.width[t];
    const long long e = (long long)(blockIdx.x - b.block_start[t]) * DENSIFY_THREADS + threadIdx.x;   // element of src
    if (e >= P * w) return;
    const long long i = e / w;
    const int col = (int)(e - i * w);
    const uint8_t f = flags[i];
    if ((f & 7u) == 0u) return;
    const int4 pl = plan[i];
    const float val = b.src[t][e];
    float* dst = b.dst[t];
    float* dst_m = b.dst_m[t];
    float* dst_v = b.dst_v[t];
    if (f & 1u) {
        const long long o = (long long)pl.x * w + col;
        put(dst, dst_m, dst_v, o, val, dst_m ? b.src_m[t][e] : 0.f, dst_v ? b.src_v[t][e] : 0.f);
    }
    if (f & 2u) put(dst, dst_m, dst_v, ((long long)T0 + pl.y) * w + col, val, 0.f, 0.f);
    if (f & 4u) {
        const int role = b.role[t];
        float rrow[3] = {0.f, 0.f, 0.f};
        if (role == 1) {
            // row `col` of build_rotation(rotation[i]) (utils/general_utils.py:130-151)
            const float qr = rotation[4 * i], qx = rotation[4 * i + 1], qy = rotation[4 * i + 2],
                        qz = rotation[4 * i + 3];
            const float norm = sqrtf(qr * qr + qx * qx + qy * qy + qz * qz);
            const float r = qr / norm, x = qx / norm, y = qy / norm, z = qz / norm;
            if (col == 0) { rrow[0] = 1 - 2 * (y * y + z * z); rrow[1] = 2 * (x * y - r * z); rrow[2] = 2 * (x * z + r * y); }
            else if (col == 1) { rrow[0] = 2 * (x * y + r * z); rrow[1] = 1 - 2 * (x * x + z * z); rrow[2] = 2 * (y * z - r * x); }
            else { rrow[0] = 2 * (x * z - r * y); rrow[1] = 2 * (y * z + r * x); rrow[2] = 1 - 2 * (x * x + y * y); }
        }
        for (int r = 0; r < N; r++) {
            float cv = val;
            if (role == 1) {
                const float* s = samples + 3 * ((long long)r * n_split + pl.w);
                cv = (rrow[0] * s[0] + rrow[1] * s[1] + rrow[2] * s[2]) + val;
            } else if (role == 2) {
                cv = logf(expf(val) * inv_split);
            }
            put(dst, dst_m, dst_v, ((long long)T0 + T1 + (long long)r * T2 + pl.z) * w + col, cv, 0.f, 0.f);
        }
    }
}

}  // namespace bags

using namespace bags;

extern "C" int bags_densify_gather(long long P, int N, const int32_t* plan, const uint8_t* flags, int kept, int clones,
                                   int children, int split_parents, const float* rotation, const float* samples,
                                   int ntensors, const bags_densify_tensor* tensors, void* stream) {
    if (P < 0 || N < 1 || ntensors < 0 || (ntensors > 0 && !tensors) || kept < 0 || clones < 0 || children < 0 ||
        split_parents < 0 || (P > 0 && (!plan || !flags)) || (children > 0 && (!rotation || !samples))) {
        set_error("bags_densify_gather: invalid argument");
        return BAGS_E_INVALID;
    }
    if ((long long)kept + clones + (long long)N * children >= (1ll << 31)) {
        set_error("bags_densify_gather: more than 2^31 output rows");
        return BAGS_E_OVERFLOW;
    }
    if (P == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const float inv_split = 1.0f / (float)(0.8 * N);   // torch: x / 1.6 on a CUDA tensor = x * float(1 / float(1.6))
    int i = 0;
    while (i < ntensors) {
        DensifyBatch b;
        b.count = 0;
        unsigned long long blocks = 0;
        b.block_start[0] = 0;
        while (i < ntensors && b.count < DENSIFY_MAX_TENSORS) {
            const bags_densify_tensor& t = tensors[i];
            if (t.width < 1 || !t.src || !t.dst || ((t.src_m != nullptr) != (t.dst_m != nullptr)) ||
                ((t.src_v != nullptr) != (t.dst_v != nullptr)) || t.role < 0 || t.role > 2 ||
                (t.role == 1 && t.width != 3)) {
                set_error("bags_densify_gather: tensor %d: bad descriptor", i);
                return BAGS_E_INVALID;
            }
            const unsigned long long nb =
                (unsigned long long)((P * t.width + DENSIFY_THREADS - 1) / DENSIFY_THREADS);
            if (blocks + nb > 0x7fffffffull) {
                if (b.count == 0) {
                    set_error("bags_densify_gather: tensor needs more than 2^31 blocks");
                    return BAGS_E_OVERFLOW;
                }
                break;
            }
            i++;
            const int k = b.count++;
            b.src[k] = t.src; b.dst[k] = t.dst; b.src_m[k] = t.src_m; b.dst_m[k] = t.dst_m;
            b.src_v[k] = t.src_v; b.dst_v[k] = t.dst_v; b.width[k] = t.width; b.role[k] = t.role;
            blocks += nb;
            b.block_start[k + 1] = (unsigned int)blocks;
        }
        if (b.count == 0) continue;
        densify_gather_kernel<<<(unsigned int)blocks, DENSIFY_THREADS, 0, st>>>(
            b, P, N, reinterpret_cast<const int4*>(plan), flags, kept, clones, children, split_parents, rotation,
            samples, inv_split);
        BAGS_LAUNCH_OK("densify_gather_kernel");
    }
    return 0;
}
