// Internal shared definitions for the sm_100a rasterizer kernels (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <stdio.h>

namespace bags {

typedef uint32_t u32;
typedef uint64_t u64;

// Tile geometry is part of the parity contract (config.h:16-17 of the reference: 16x16 pixel tiles).
constexpr int TILE_X = 16;
constexpr int TILE_Y = 16;
constexpr int TILE_PIX = TILE_X * TILE_Y;

// ---- error plumbing --------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch();
void prof_begin(int stage, cudaStream_t st);
void prof_end(int stage, cudaStream_t st);
struct ProfScope {
    int stage; cudaStream_t st;
    ProfScope(int s, cudaStream_t t) : stage(s), st(t) { prof_begin(s, t); }
    ~ProfScope() { prof_end(stage, st); }
};
#define BAGS_CUDA_OK(expr)                                                                       \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            bags::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,                 \
                            cudaGetErrorString(_e));                                             \
            (void)cudaGetLastError();   /* reported through the return code: do not leave it for the next caller */ \
            return BAGS_E_CUDA;                                                                  \
        }                                                                                        \
    } while (0)
#define BAGS_LAUNCH_OK(name)                                                                     \
    do {                                                                                         \
        cudaError_t _e = cudaGetLastError();                                                     \
        if (_e != cudaSuccess) {                                                                 \
            bags::set_error("launch of %s failed at %s:%d: %s", name, __FILE__, __LINE__,        \
                            cudaGetErrorString(_e));                                             \
            return BAGS_E_CUDA;                                                                  \
        }                                                                                        \
        bags::count_launch();                                                                    \
    } while (0)

// ---- opaque state layouts --------------------------------------------------------------------------------------
// Every sub-array starts on a 256-byte boundary relative to a 256-byte aligned base.
inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

template <typename T>
inline void carve(char*& p, T*& out, size_t count) {
    out = reinterpret_cast<T*>(p);
    p += align_up(count * sizeof(T));
}

// Radix-sort bookkeeping shared by the depth sort (P items) and the tile sort (R items).
constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 16;                         // keys per thread of a throughput-bound pass (several waves)
constexpr int SORT_ITEMS_SMALL = 8;                   // ... of a latency-bound one (all CTAs resident at once)
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;  // 4096 keys per CTA
constexpr int SORT_TILE_SMALL = SORT_THREADS * SORT_ITEMS_SMALL;   // 2048; the scratch is sized for this one
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int SORT_MAX_PASSES = 4;

struct SortScratch {
    u32* ghist;        // [SORT_MAX_PASSES][RADIX] global digit histograms
    u32* tile_counter; // [SORT_MAX_PASSES] dynamic tile ids
    u32* status;       // [SORT_MAX_PASSES][ntiles][RADIX] decoupled look-back words
    size_t zero_bytes; // bytes from ghist to the end that must be zeroed before a sort
    int ntiles;
    static size_t bytes(int n) {
        size_t nt = (size_t)(n + SORT_TILE_SMALL - 1) / SORT_TILE_SMALL;
        if (nt == 0) nt = 1;
        return align_up(SORT_MAX_PASSES * RADIX * sizeof(u32)) + align_up(SORT_MAX_PASSES * sizeof(u32)) +
               align_up((size_t)SORT_MAX_PASSES * nt * RADIX * sizeof(u32));
    }
    static SortScratch from(char*& p, int n) {
        SortScratch s;
        char* p0 = p;
        size_t nt = (size_t)(n + SORT_TILE_SMALL - 1) / SORT_TILE_SMALL;
        if (nt == 0) nt = 1;
        s.ntiles = (int)nt;      // capacity in tiles (stride of a pass's status words)
        carve(p, s.ghist, SORT_MAX_PASSES * RADIX);
        carve(p, s.tile_counter, SORT_MAX_PASSES);
        carve(p, s.status, (size_t)SORT_MAX_PASSES * nt * RADIX);
        s.zero_bytes = (size_t)(p - p0);
        return s;
    }
};

// Per-Gaussian state kept from forward to backward.
//   rec[3g+0] = {pix.x, pix.y, depth, bits(clamped mask)}   rec[3g+1] = {conic.x, conic.y, conic.z, opacity}
//   rec[3g+2] = {r, g, b, 0}
struct GeomState {
    float4* rec;       // [3P]
    ushort4* rect;     // [P] {min.x, min.y, width, height} in tiles; width*height == tiles_touched
    // forward-only temporaries (depth sort ping-pong; after the 4 passes the result is back in key_a / idx_a)
    u32* key_a;        // [P] float bits of view-space depth, 0xFFFFFFFF when culled
    u32* key_b;        // [P]
    u32* idx_a;        // [P] Gaussian ids in (depth, id) order after the sort (culled ones last)
    u32* idx_b;        // [P]
    unsigned long long* counters;  // [4]: [0] = total tile instances, [1..3] spare
    u32* emit_status;  // [ceil(P/EMIT_THREADS)] per-CTA tile-instance totals of the emit scan
    SortScratch sort;
    static size_t bytes(int P);
    static GeomState from(char* p, int P);
};

constexpr int EMIT_THREADS = 1024;  // Gaussians per emit CTA (= per chained-scan tile)

// Tile-sort ping-pong: emit fills (tile_buf[0], id_buf[0]); after `npass` passes the sorted data is in buffer npass&1.
struct BinningState {
    u32* tile_buf[2];   // [R] tile id of every instance
    u32* id_buf[2];     // [R] Gaussian id of every instance
    u32* point_list;    // = id_buf[npass & 1]: ids sorted by (tile, depth, id) -- the only part backward needs
    u32* tile_sorted;   // = tile_buf[npass & 1]
    SortScratch sort;
    static size_t bytes(int R);
    static BinningState from(char* p, int R, int npass);
};
// number of 8-bit passes needed to sort tile ids < ntiles
inline int tile_sort_bits(int ntiles) {
    int bits = 1;
    while ((1 << bits) < ntiles) bits++;
    return bits;
}
inline int tile_sort_passes(int ntiles) { return (tile_sort_bits(ntiles) + RADIX_BITS - 1) / RADIX_BITS; }

struct ImageState {
    float* final_T;   // [N]
    u32* n_contrib;   // [N]
    uint2* ranges;    // [tiles]
    // longest-first launch order of the blend CTAs (see tile_order_kernel): CTA i of a blend grid works on tile order[i]
    u32* fwd_order;   // [tiles] tiles by descending list length
    u32* bwd_cost;    // [tiles] largest n_contrib of the tile's pixels (written by the forward blend)
    u32* bwd_order;   // [tiles] tiles by descending bwd_cost
    static size_t bytes(int width, int height);
    static ImageState from(char* p, int width, int height);
};

// ---- batched launches (bags_rasterize_forward_batch / _backward_batch) ---------------------------------------------
// B views of the same Gaussians are one launch set: every per-view kernel gets the view as blockIdx.y (binning) or
// blockIdx.z (blend), and the views' private states are B consecutive copies of the single-view layouts, so view v's
// arrays sit `v * stride` bytes behind view 0's.  All strides are 0 for a single view.
constexpr int MAX_BATCH = 8;
struct ViewStride {
    size_t geom = 0, bin = 0, img = 0;
};
#ifdef __CUDACC__
template <typename T>
__device__ __forceinline__ T* at_view(T* p, const size_t stride_bytes, const unsigned view) {
    return reinterpret_cast<T*>(reinterpret_cast<size_t>(p) + stride_bytes * view);
}
#endif
// per-view projection constants of a batch (kernel parameter, by value)
struct ViewFov {
    float tan_fovx[MAX_BATCH], tan_fovy[MAX_BATCH], focal_x[MAX_BATCH], focal_y[MAX_BATCH];
};
#ifdef __CUDACC__
// Kernel parameters can only be indexed with compile-time constants without a local-memory copy: thread v of the CTA
// moves view v's constants to shared memory ({tan_fovx, tan_fovy, focal_x, focal_y}).
__device__ __forceinline__ void fov_to_shared(const ViewFov& fov, float4* s_fov) {
#pragma unroll
    for (int v = 0; v < MAX_BATCH; v++)
        if (threadIdx.x == v) s_fov[v] = make_float4(fov.tan_fovx[v], fov.tan_fovy[v], fov.focal_x[v], fov.focal_y[v]);
}
#endif

inline int tiles_x(int w) { return (w + TILE_X - 1) / TILE_X; }
inline int tiles_y(int h) { return (h + TILE_Y - 1) / TILE_Y; }

// ---- kernel launchers (defined in the .cu files) ---------------------------------------------------------------
struct FwdArgs {
    int P, D, M, W, H;
    const float* background;
    const float* means3D;
    const float* shs;
    const float* colors_precomp;
    const float* opacities;
    const float* scales;
    float scale_modifier;
    const float* rotations;
    const float* cov3D_precomp;
    const float* viewmatrix;
    const float* projmatrix;
    const float* cam_pos;
    const float* shift_factors;
    float tan_fovx, tan_fovy, focal_x, focal_y;
    int prefiltered;
};

int launch_preprocess_fwd(const FwdArgs& a, GeomState& g, int* radii, float* mean2Dx, float* mean2Dy,
                          cudaStream_t st, int nviews = 1, const ViewFov* fov = nullptr, size_t geom_stride = 0);
int launch_unpack(int P, int R, int W, int H, const GeomState& g, const BinningState& b, const ImageState& img,
                  u32* tiles_touched, u32* depth_bits, float* xy, float* conic_opacity, float* rgb, u32* point_list,
                  u64* point_list_keys, u32* ranges, u32* n_contrib, float* final_T, cudaStream_t st);
// n_dev / R_dev: device location of the element count when the host only knows a capacity (NULL: the count is exact)
// hist_mode of launch_sort_pairs: SORT_FRESH zeroes the bookkeeping and counts the digits itself; SORT_PREPARED: the
// caller has zeroed the bookkeeping (sort_prepare, or as part of a larger memset); SORT_HIST_DONE: ... and a producer
// kernel has already filled the digit histograms of keys_in (no histogram launch).
enum { SORT_FRESH = 0, SORT_PREPARED = 1, SORT_HIST_DONE = 2 };
int sort_prepare(SortScratch& sc, cudaStream_t st);
int launch_sort_pairs(int n, const unsigned long long* n_dev, u32* keys_in, u32* vals_in, u32* keys_out,
                      u32* vals_out, int begin_bit, int end_bit, SortScratch& sc, bool iota_vals, int hist_mode,
                      cudaStream_t st, u32** final_keys, u32** final_vals, int nviews = 1, size_t data_stride = 0,
                      size_t count_stride = 0);
// nviews > 1: one launch for the views of a batch (pointers are view 0's; see ViewStride)
int launch_emit_scan(int P, const GeomState& g, const u32* sorted_idx, u32* offsets, cudaStream_t st, int nviews = 1,
                     const ViewStride& vs = ViewStride());
int launch_emit_fill(int P, int R, const unsigned long long* R_dev, const GeomState& g, const u32* sorted_idx,
                     const u32* offsets, BinningState& b, int grid_x, int hist_passes, cudaStream_t st,
                     int nviews = 1, const ViewStride& vs = ViewStride());
int set_sort_lookback(int value);
int set_sort_small_tiles(int value);
int set_sort_hwmatch(int value);
int set_pbwd_min_blocks(int value);
int launch_ranges(int R, const unsigned long long* R_dev, int ntiles, const u32* tile_sorted, uint2* ranges,
                  cudaStream_t st, int nviews = 1, const ViewStride& vs = ViewStride());
// order[i] = i-th tile in descending `cost` (cost = list length when `ranges` is given, else cost_in[tile]); one launch
// for the views of a batch
int launch_tile_order(int ntiles, const uint2* ranges, const u32* cost_in, u32* order, cudaStream_t st, int nviews = 1,
                      const ViewStride& vs = ViewStride());
int launch_blend_fwd(int W, int H, const GeomState& g, const BinningState& b, ImageState& img, const float* bg,
                     float* out_color, float* out_depth, float* out_alpha, cudaStream_t st, int nviews = 1,
                     const ViewStride& vs = ViewStride());
int launch_blend_bwd(int W, int H, const GeomState& g, const BinningState& b, const ImageState& img, const float* bg,
                     const float* dL_dpix, const float* dL_dpix_depth, const float* dL_dalpha, float* gacc,
                     cudaStream_t st, int nviews = 1, const ViewStride& vs = ViewStride(), size_t gacc_stride = 0);

struct BwdArgs {
    int P, D, M, W, H;
    const float* means3D;
    const float* shs;
    const float* colors_precomp;
    const float* scales;
    float scale_modifier;
    const float* rotations;
    const float* cov3D_precomp;
    const float* viewmatrix;
    const float* projmatrix;
    const float* campos;
    float tan_fovx, tan_fovy, focal_x, focal_y;
    const int* radii;
};
struct BwdOut {
    float* dL_dmean2D;
    float* dL_dmean2D_densify;
    float* covariance;
    float* dL_dopacity;
    float* dL_dcolor;
    float* dL_dmean3D;
    float* dL_dcov3D;
    float* dL_dsh;
    float* dL_dscale;
    float* dL_drot;
    float* dL_dprojmatrix;
    float* dL_dviewmatrix;
    float* dL_dcampos;
    float* dL_dshift;
};
// gacc: [P][16] accumulated by blend_bwd; pose_acc: [32] doubles (sums + winner), zeroed by the caller.
int launch_preprocess_bwd(const BwdArgs& a, const GeomState& g, const float* gacc, double* pose_acc,
                          const BwdOut& o, unsigned accumulate_mask, cudaStream_t st);

constexpr int POSE_ACC_DOUBLES = 32;   // pose_acc doubles per view
// outputs of the batched backward preprocess (see preprocess_bwd.cu)
struct BwdBatchOut {
    float* dL_dmean2D;          // [B][P][3]
    float* dL_dmean2D_densify;  // [B][P][3]
    float* covariance;          // [B][P] or NULL
    float* dL_dopacity;         // [P]        summed over the views (as are all the following)
    float* dL_dcolor;           // [P][3] or NULL
    float* dL_dmean3D;          // [P][3]
    float* dL_dcov3D;           // [P][6] or NULL
    float* dL_dsh;              // [P][M][3] or NULL
    float* dL_dscale;           // [P][3]
    float* dL_drot;             // [P][4]
};
int launch_preprocess_bwd_batch(const BwdArgs& a, int B, const ViewFov& fov, const GeomState& g, size_t geom_stride,
                                const float* gacc, size_t gacc_stride, double* pose_acc, const BwdOut& po,
                                const BwdBatchOut& o, unsigned accumulate_mask, int first, int count, bool finalize,
                                cudaStream_t st);

constexpr int GACC_STRIDE = 16;
// gacc record layout (floats)
enum { GA_R = 0, GA_G = 1, GA_B = 2, GA_DEPTH = 3, GA_MX = 4, GA_MY = 5, GA_AMX = 6, GA_AMY = 7,
       GA_CA = 8, GA_CB = 9, GA_CC = 10, GA_OP = 11, GA_COV = 12 };

}  // namespace bags
