// C ABI entry points (include/bags_b200.h): state layouts, host orchestration of the forward / backward pipelines,
// and the test-only state unpacker.  Host logic follows CudaRasterizer::Rasterizer::forward / backward
// (3dgs-pose/cuda_rasterizer/rasterizer_impl.cu:198-349,353-475) and RasterizeGaussians[Backward]CUDA
// (3dgs-pose/rasterize_points.cu:36-276) of the reference; see DESIGN.md for what is fused or reordered.
#include "../../include/bags_b200.h"
#include "common.cuh"

#include <stdarg.h>
#include <string.h>
#include <mutex>
#include <vector>

namespace bags {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- measurement hooks -------------------------------------------------------------------------------------------
static unsigned long long g_launches = 0;
void count_launch() { __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED); }

struct ProfRec { int stage; cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::mutex g_prof_mu;             // the hooks may be reached from several host threads
static std::vector<ProfRec> g_prof_recs;
static std::vector<cudaEvent_t> g_prof_pool;
static std::vector<ProfRec> g_prof_open;

static cudaEvent_t prof_event() {
    if (!g_prof_pool.empty()) {
        cudaEvent_t e = g_prof_pool.back();
        g_prof_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

void prof_begin(int stage, cudaStream_t st) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRec r;
    r.stage = stage;
    r.a = prof_event();
    r.b = prof_event();
    cudaEventRecord(r.a, st);
    g_prof_open.push_back(r);
}

void prof_end(int stage, cudaStream_t st) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (size_t i = g_prof_open.size(); i-- > 0;) {
        if (g_prof_open[i].stage == stage) {
            ProfRec r = g_prof_open[i];
            g_prof_open.erase(g_prof_open.begin() + i);
            cudaEventRecord(r.b, st);
            g_prof_recs.push_back(r);
            return;
        }
    }
}

// ---- state layouts -----------------------------------------------------------------------------------------------
static inline size_t emit_tiles(int P) { return (size_t)(P + EMIT_THREADS - 1) / EMIT_THREADS + 1; }

size_t GeomState::bytes(int P) {
    size_t n = (size_t)P;
    return align_up(3 * n * sizeof(float4)) + align_up(n * sizeof(ushort4)) + 4 * align_up(n * sizeof(u32)) +
           align_up(4 * sizeof(unsigned long long)) + align_up(emit_tiles(P) * sizeof(u32)) + SortScratch::bytes(P) +
           256;
}

GeomState GeomState::from(char* p, int P) {
    GeomState g;
    p = (char*)align_up((size_t)p);
    size_t n = (size_t)P;
    carve(p, g.rec, 3 * n);
    carve(p, g.rect, n);
    carve(p, g.key_a, n);
    carve(p, g.key_b, n);
    carve(p, g.idx_a, n);
    carve(p, g.idx_b, n);
    carve(p, g.counters, 4);
    carve(p, g.emit_status, emit_tiles(P));
    g.sort = SortScratch::from(p, P);
    return g;
}

size_t BinningState::bytes(int R) {
    size_t n = (size_t)(R > 0 ? R : 1);
    return 4 * align_up(n * sizeof(u32)) + SortScratch::bytes(R) + 256;
}

BinningState BinningState::from(char* p, int R, int npass) {
    BinningState b;
    p = (char*)align_up((size_t)p);
    size_t n = (size_t)(R > 0 ? R : 1);
    // id buffers first: the one holding point_list is what backward keeps reading
    carve(p, b.id_buf[0], n);
    carve(p, b.id_buf[1], n);
    carve(p, b.tile_buf[0], n);
    carve(p, b.tile_buf[1], n);
    b.sort = SortScratch::from(p, R);
    b.point_list = b.id_buf[npass & 1];
    b.tile_sorted = b.tile_buf[npass & 1];
    return b;
}

size_t ImageState::bytes(int width, int height) {
    size_t N = (size_t)width * height;
    size_t T = (size_t)tiles_x(width) * tiles_y(height);
    return 2 * align_up(N * 4) + align_up(T * sizeof(uint2)) + 3 * align_up(T * sizeof(u32)) + 256;
}

ImageState ImageState::from(char* p, int width, int height) {
    ImageState s;
    p = (char*)align_up((size_t)p);
    size_t N = (size_t)width * height;
    size_t T = (size_t)tiles_x(width) * tiles_y(height);
    carve(p, s.final_T, N);
    carve(p, s.n_contrib, N);
    carve(p, s.ranges, T);
    carve(p, s.fwd_order, T);
    carve(p, s.bwd_cost, T);
    carve(p, s.bwd_order, T);
    return s;
}

// ---- host resources of the synchronous forward: a pinned word for num_rendered and an event per call in flight ------
// Slots are pooled per device under a mutex and owned by one call at a time, so concurrent forwards from several host
// threads on one GPU (a viewer or evaluation thread next to the training thread) never share a word or an event.
struct SyncSlot {
    unsigned long long* pinned = nullptr;
    cudaEvent_t ev = nullptr;
    int dev = -1;
};
static std::mutex g_slot_mu;
static std::vector<SyncSlot> g_slot_pool;

static int acquire_slot(SyncSlot* out) {
    int dev = 0;
    BAGS_CUDA_OK(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lk(g_slot_mu);
        for (size_t i = 0; i < g_slot_pool.size(); i++) {
            if (g_slot_pool[i].dev == dev) {
                *out = g_slot_pool[i];
                g_slot_pool.erase(g_slot_pool.begin() + i);
                return 0;
            }
        }
    }
    SyncSlot s;
    s.dev = dev;
    BAGS_CUDA_OK(cudaMallocHost((void**)&s.pinned, 64));
    BAGS_CUDA_OK(cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
    *out = s;
    return 0;
}

static void release_slot(const SyncSlot& s) {
    std::lock_guard<std::mutex> lk(g_slot_mu);
    g_slot_pool.push_back(s);
}
struct SlotGuard {
    SyncSlot s;
    bool held = false;
    ~SlotGuard() { if (held) release_slot(s); }
};

// ---- unpack kernels (test-only) ----------------------------------------------------------------------------------
__global__ void unpack_geom_kernel(int P, const float4* __restrict__ rec, const ushort4* __restrict__ rect,
                                   u32* tiles_touched, u32* depth_bits, float* xy, float* conic_opacity,
                                   float* rgb) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    ushort4 r = rect[i];
    u32 tt = (u32)r.z * (u32)r.w;
    if (tiles_touched) tiles_touched[i] = tt;
    float4 q0 = make_float4(0, 0, 0, 0), q1 = q0, q2 = q0;
    if (tt) {
        q0 = rec[3 * (size_t)i];
        q1 = rec[3 * (size_t)i + 1];
        q2 = rec[3 * (size_t)i + 2];
    }
    if (depth_bits) depth_bits[i] = __float_as_uint(q0.z);
    if (xy) { xy[2 * i] = q0.x; xy[2 * i + 1] = q0.y; }
    if (conic_opacity) {
        conic_opacity[4 * i] = q1.x; conic_opacity[4 * i + 1] = q1.y;
        conic_opacity[4 * i + 2] = q1.z; conic_opacity[4 * i + 3] = q1.w;
    }
    if (rgb) { rgb[3 * i] = q2.x; rgb[3 * i + 1] = q2.y; rgb[3 * i + 2] = q2.z; }
}

// R_cap slots; only the first min(R_cap, *count) hold instances (asynchronous forward: R_cap is the caller's capacity)
__global__ void unpack_binning_kernel(int R_cap, const unsigned long long* __restrict__ count,
                                      const u32* __restrict__ point_list, const u32* __restrict__ tile_sorted,
                                      const float4* __restrict__ rec, u32* out_list, u64* out_keys) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long c = *count;
    const int R = c < (unsigned long long)R_cap ? (int)c : R_cap;
    if (i >= R_cap) return;
    if (i >= R) {
        if (out_list) out_list[i] = 0u;
        if (out_keys) out_keys[i] = 0ull;
        return;
    }
    u32 id = point_list[i];
    if (out_list) out_list[i] = id;
    if (out_keys) out_keys[i] = ((u64)tile_sorted[i] << 32) | (u64)__float_as_uint(rec[3 * (size_t)id].z);
}

int launch_unpack(int P, int R, int W, int H, const GeomState& g, const BinningState& b, const ImageState& img,
                  u32* tiles_touched, u32* depth_bits, float* xy, float* conic_opacity, float* rgb, u32* point_list,
                  u64* point_list_keys, u32* ranges, u32* n_contrib, float* final_T, cudaStream_t st) {
    if (P > 0 && (tiles_touched || depth_bits || xy || conic_opacity || rgb)) {
        unpack_geom_kernel<<<(P + 255) / 256, 256, 0, st>>>(P, g.rec, g.rect, tiles_touched, depth_bits, xy,
                                                            conic_opacity, rgb);
        BAGS_LAUNCH_OK("unpack_geom_kernel");
    }
    if (R > 0 && (point_list || point_list_keys)) {
        unpack_binning_kernel<<<(R + 255) / 256, 256, 0, st>>>(R, g.counters, b.point_list, b.tile_sorted, g.rec,
                                                               point_list, point_list_keys);
        BAGS_LAUNCH_OK("unpack_binning_kernel");
    }
    size_t N = (size_t)W * H, T = (size_t)tiles_x(W) * tiles_y(H);
    if (ranges) BAGS_CUDA_OK(cudaMemcpyAsync(ranges, img.ranges, T * sizeof(uint2), cudaMemcpyDeviceToDevice, st));
    if (n_contrib) BAGS_CUDA_OK(cudaMemcpyAsync(n_contrib, img.n_contrib, N * 4, cudaMemcpyDeviceToDevice, st));
    if (final_T) BAGS_CUDA_OK(cudaMemcpyAsync(final_T, img.final_T, N * 4, cudaMemcpyDeviceToDevice, st));
    return 0;
}

}  // namespace bags

using namespace bags;

extern "C" const char* bags_last_error(void) { return g_err; }
extern "C" unsigned long long bags_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
extern "C" int bags_profile_enable(int on) {
    g_prof_on = on != 0;
    return 0;
}
extern "C" int bags_profile_read(double* ms, unsigned long long* calls) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : g_prof_recs) {
        BAGS_CUDA_OK(cudaEventSynchronize(r.b));
        float t = 0.f;
        BAGS_CUDA_OK(cudaEventElapsedTime(&t, r.a, r.b));
        if (r.stage >= 0 && r.stage < BAGS_STAGE_COUNT) {
            if (ms) ms[r.stage] += t;
            if (calls) calls[r.stage] += 1;
        }
        g_prof_pool.push_back(r.a);
        g_prof_pool.push_back(r.b);
    }
    g_prof_recs.clear();
    return 0;
}
extern "C" int bags_version(void) { return 100; }

extern "C" size_t bags_geom_bytes(int P) { return GeomState::bytes(P < 0 ? 0 : P); }
extern "C" size_t bags_image_bytes(int width, int height) { return ImageState::bytes(width, height); }
extern "C" size_t bags_binning_bytes(int P, int num_rendered, int width, int height) {
    (void)P; (void)width; (void)height;
    return BinningState::bytes(num_rendered);
}
extern "C" size_t bags_backward_scratch_bytes(int P) {
    return align_up((size_t)(P < 0 ? 0 : P) * GACC_STRIDE * sizeof(float)) + 256 + 256;
}

#define BAGS_REQUIRE(cond, msg)                      \
    do {                                             \
        if (!(cond)) {                               \
            set_error("%s: %s", __func__, msg);      \
            return BAGS_E_INVALID;                   \
        }                                            \
    } while (0)

// Shared body of the two forward entry points.
//   capacity < 0  (bags_rasterize_forward): the reference's protocol -- num_rendered is read back on the host (one
//                 blocking wait per view, rasterizer_impl.cu:291), the binning buffer has exactly that many slots.
//   capacity >= 0 (bags_rasterize_forward_async): nothing waits.  The binning buffer, the grids of the emit / tile-sort /
//                 range kernels and the sort bookkeeping are sized by `capacity`; the kernels read the actual count from
//                 the geometry state, where the preprocess kernel left it, and clamp it to the capacity.  The count is
//                 also copied to *count_out (device or pinned host memory) for the caller to look at whenever it likes.
static int forward_impl(
    bags_alloc_fn geom_alloc, void* geom_user, bags_alloc_fn binning_alloc, void* binning_user,
    bags_alloc_fn image_alloc, void* image_user, int P, int D, int M, const float* background, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp, const float* opacities, const float* scales,
    float scale_modifier, const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
    const float* projmatrix, const float* cam_pos, const float* shift_factors, float tan_fovx, float tan_fovy,
    int prefiltered, float* out_color, float* out_depth, float* out_alpha, int* radii, float* mean2Dx, float* mean2Dy,
    int debug, int capacity, unsigned long long* count_out, cudaStream_t st) {
    const bool async = capacity >= 0;
    BAGS_REQUIRE(P >= 0 && width > 0 && height > 0, "negative P or empty image");
    BAGS_REQUIRE(out_color && out_depth && out_alpha, "null output image");
    BAGS_REQUIRE(geom_alloc && binning_alloc && image_alloc, "null allocation callback");
    BAGS_REQUIRE(!async || capacity < (1 << 30), "capacity does not fit the 30-bit look-back words");
    const size_t N = (size_t)width * height;
    if (P == 0) {
        // reference: kernels skipped, zero-initialised outputs returned (rasterize_points.cu:66-83)
        BAGS_CUDA_OK(cudaMemsetAsync(out_color, 0, 3 * N * sizeof(float), st));
        BAGS_CUDA_OK(cudaMemsetAsync(out_depth, 0, N * sizeof(float), st));
        BAGS_CUDA_OK(cudaMemsetAsync(out_alpha, 0, N * sizeof(float), st));
        if (async && count_out) BAGS_CUDA_OK(cudaMemsetAsync(count_out, 0, sizeof(unsigned long long), st));
        return 0;
    }
    BAGS_REQUIRE(background && means3D && opacities && viewmatrix && projmatrix && cam_pos && shift_factors,
                 "null required input");
    BAGS_REQUIRE(radii && mean2Dx && mean2Dy, "null per-Gaussian output");
    BAGS_REQUIRE((shs != nullptr) != (colors_precomp != nullptr), "provide exactly one of shs / colors_precomp");
    BAGS_REQUIRE((scales != nullptr && rotations != nullptr) != (cov3D_precomp != nullptr),
                 "provide exactly one of scales+rotations / cov3D_precomp");
    BAGS_REQUIRE(shs == nullptr || (M >= (D + 1) * (D + 1) && D >= 0 && D <= 3), "SH degree / coefficient count");
    BAGS_REQUIRE(tiles_x(width) <= 65535 && tiles_y(height) <= 65535, "image too large for 16-bit tile coordinates");

    int rc;
    SlotGuard slot;
    if (!async) {
        rc = acquire_slot(&slot.s);
        if (rc) return rc;
        slot.held = true;
    }

    char* geom_ptr = geom_alloc(GeomState::bytes(P), geom_user);
    char* img_ptr = image_alloc(ImageState::bytes(width, height), image_user);
    if (!geom_ptr || !img_ptr) {
        set_error("allocation callback returned NULL");
        return BAGS_E_ALLOC;
    }
    GeomState g = GeomState::from(geom_ptr, P);
    ImageState img = ImageState::from(img_ptr, width, height);
    const int ntiles = tiles_x(width) * tiles_y(height);
    const int npass = tile_sort_passes(ntiles);

    // Asynchronous mode: everything that could fail on the host is done before the first launch
    char* bin_ptr = nullptr;
    if (async) {
        bin_ptr = binning_alloc(BinningState::bytes(capacity), binning_user);
        if (!bin_ptr) {
            set_error("binning allocation callback returned NULL");
            return BAGS_E_ALLOC;
        }
    }

    FwdArgs a;
    a.P = P; a.D = D; a.M = M; a.W = width; a.H = height;
    a.background = background; a.means3D = means3D; a.shs = shs; a.colors_precomp = colors_precomp;
    a.opacities = opacities; a.scales = scales; a.scale_modifier = scale_modifier; a.rotations = rotations;
    a.cov3D_precomp = cov3D_precomp; a.viewmatrix = viewmatrix; a.projmatrix = projmatrix; a.cam_pos = cam_pos;
    a.shift_factors = shift_factors; a.tan_fovx = tan_fovx; a.tan_fovy = tan_fovy;
    a.focal_y = height / (2.0f * tan_fovy);
    a.focal_x = width / (2.0f * tan_fovx);
    a.prefiltered = prefiltered;

    // counters + emit totals + the depth sort's bookkeeping are contiguous: one memset
    BAGS_CUDA_OK(cudaMemsetAsync(g.counters, 0, ((char*)g.sort.ghist - (char*)g.counters) + g.sort.zero_bytes, st));
    {
        ProfScope ps(BAGS_STAGE_PREPROCESS, st);
        rc = launch_preprocess_fwd(a, g, radii, mean2Dx, mean2Dy, st);
    }
    if (rc) return rc;

    if (async) {
        if (count_out)
            BAGS_CUDA_OK(cudaMemcpyAsync(count_out, g.counters, sizeof(unsigned long long), cudaMemcpyDefault, st));
    } else {
        // num_rendered travels to the host while the depth sort runs
        BAGS_CUDA_OK(cudaMemcpyAsync(slot.s.pinned, g.counters, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                     st));
        BAGS_CUDA_OK(cudaEventRecord(slot.s.ev, st));
    }

    u32 *sorted_keys, *sorted_idx;
    {
        ProfScope ps(BAGS_STAGE_DEPTH_SORT, st);
        rc = launch_sort_pairs(P, nullptr, g.key_a, g.idx_a, g.key_b, g.idx_b, 0, 32, g.sort, true, SORT_PREPARED, st,
                               &sorted_keys, &sorted_idx);
    }
    if (rc) return rc;

    // offsets of every Gaussian's tile instances (depth order) go to the key buffer the sort did not end in; this
    // does not need num_rendered, so it also runs while the host waits
    u32* offsets = (sorted_keys == g.key_a) ? g.key_b : g.key_a;
    prof_begin(BAGS_STAGE_EMIT, st);
    rc = launch_emit_scan(P, g, sorted_idx, offsets, st);
    prof_end(BAGS_STAGE_EMIT, st);
    if (rc) return rc;

    int R = capacity;                                   // slots of the binning buffer
    const unsigned long long* R_dev = async ? g.counters : nullptr;
    if (!async) {
        BAGS_CUDA_OK(cudaEventSynchronize(slot.s.ev));
        const unsigned long long R64 = *slot.s.pinned;
        if (R64 >= (1ull << 30)) {
            set_error("num_rendered = %llu does not fit the 30-bit look-back words", R64);
            return BAGS_E_OVERFLOW;
        }
        R = (int)R64;
        bin_ptr = binning_alloc(BinningState::bytes(R), binning_user);
        if (!bin_ptr) {
            set_error("binning allocation callback returned NULL");
            return BAGS_E_ALLOC;
        }
    }
    BinningState b = BinningState::from(bin_ptr, R, npass);

    if (R > 0) {
        const int hist_passes = npass <= 2 ? npass : 0;     // the fill kernel produces the sort's digit histograms
        {
            ProfScope ps(BAGS_STAGE_EMIT, st);
            if (hist_passes) {
                rc = sort_prepare(b.sort, st);
                if (rc) return rc;
            }
            rc = launch_emit_fill(P, R, R_dev, g, sorted_idx, offsets, b, tiles_x(width), hist_passes, st);
        }
        if (rc) return rc;
        u32 *fk, *fv;
        {
            ProfScope ps(BAGS_STAGE_TILE_SORT, st);
            rc = launch_sort_pairs(R, R_dev, b.tile_buf[0], b.id_buf[0], b.tile_buf[1], b.id_buf[1], 0,
                                   tile_sort_bits(ntiles), b.sort, false,
                                   hist_passes > 0 ? SORT_HIST_DONE : SORT_FRESH, st, &fk, &fv);
        }
        if (rc) return rc;
        if (fv != b.point_list || fk != b.tile_sorted) {
            set_error("internal: tile sort ping-pong parity mismatch");
            return BAGS_E_INVALID;
        }
    }
    {
        ProfScope ps(BAGS_STAGE_RANGES, st);
        rc = launch_ranges(R, R_dev, ntiles, b.tile_sorted, img.ranges, st);
    }
    if (rc) return rc;
    {
        ProfScope ps(BAGS_STAGE_BLEND_FWD, st);
        rc = launch_blend_fwd(width, height, g, b, img, background, out_color, out_depth, out_alpha, st);
    }
    if (rc) return rc;
    if (debug) BAGS_CUDA_OK(cudaStreamSynchronize(st));
    return async ? 0 : R;
}

extern "C" int bags_rasterize_forward(
    bags_alloc_fn geom_alloc, void* geom_user, bags_alloc_fn binning_alloc, void* binning_user,
    bags_alloc_fn image_alloc, void* image_user, int P, int D, int M, const float* background, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp, const float* opacities, const float* scales,
    float scale_modifier, const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
    const float* projmatrix, const float* intrinsic, const float* cam_pos, const float* shift_factors, float tan_fovx,
    float tan_fovy, int prefiltered, float* out_color, float* out_depth, float* out_alpha, int* radii, float* mean2Dx,
    float* mean2Dy, int debug, void* stream) {
    (void)intrinsic;  // passed through by the reference too, never read by live code (forward.cu:352)
    return forward_impl(geom_alloc, geom_user, binning_alloc, binning_user, image_alloc, image_user, P, D, M,
                        background, width, height, means3D, shs, colors_precomp, opacities, scales, scale_modifier,
                        rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos, shift_factors, tan_fovx, tan_fovy,
                        prefiltered, out_color, out_depth, out_alpha, radii, mean2Dx, mean2Dy, debug, -1, nullptr,
                        (cudaStream_t)stream);
}

extern "C" int bags_rasterize_forward_async(
    bags_alloc_fn geom_alloc, void* geom_user, bags_alloc_fn binning_alloc, void* binning_user,
    bags_alloc_fn image_alloc, void* image_user, int P, int D, int M, const float* background, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp, const float* opacities, const float* scales,
    float scale_modifier, const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
    const float* projmatrix, const float* intrinsic, const float* cam_pos, const float* shift_factors, float tan_fovx,
    float tan_fovy, int prefiltered, float* out_color, float* out_depth, float* out_alpha, int* radii, float* mean2Dx,
    float* mean2Dy, int capacity, unsigned long long* num_rendered_out, void* stream) {
    (void)intrinsic;
    BAGS_REQUIRE(capacity >= 0, "negative capacity");
    return forward_impl(geom_alloc, geom_user, binning_alloc, binning_user, image_alloc, image_user, P, D, M,
                        background, width, height, means3D, shs, colors_precomp, opacities, scales, scale_modifier,
                        rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos, shift_factors, tan_fovx, tan_fovy,
                        prefiltered, out_color, out_depth, out_alpha, radii, mean2Dx, mean2Dy, 0, capacity,
                        num_rendered_out, (cudaStream_t)stream);
}

extern "C" int bags_rasterize_backward_accum(
    int P, int D, int M, int R, const float* background, int width, int height, const float* means3D,
    const float* shs, const float* colors_precomp, const float* alphas, const float* scales, float scale_modifier,
    const float* rotations, const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
    const float* intrinsic, const float* campos, float tan_fovx, float tan_fovy, const int* radii,
    const char* geom_buffer, const char* binning_buffer, const char* image_buffer, const float* dL_dpix,
    const float* dL_dpix_depth, const float* dL_dalpha, char* scratch, float* dL_dmean2D, float* dL_dmean2D_densify,
    float* covariance, float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
    float* dL_dscale, float* dL_drot, float* dL_dprojmatrix, float* dL_dviewmatrix, float* dL_dcampos,
    float* dL_dshift, unsigned accumulate_mask, int debug, void* stream) {
    (void)intrinsic; (void)alphas; (void)colors_precomp;
    BAGS_REQUIRE((accumulate_mask & ~(BAGS_ACC_MEANS3D | BAGS_ACC_SH | BAGS_ACC_OPACITY | BAGS_ACC_SCALE |
                                      BAGS_ACC_ROT)) == 0, "unknown bit in accumulate_mask");
    cudaStream_t st = (cudaStream_t)stream;
    BAGS_REQUIRE(P >= 0 && R >= 0 && width > 0 && height > 0, "negative size");
    BAGS_REQUIRE(scratch && dL_dprojmatrix && dL_dviewmatrix && dL_dcampos && dL_dshift, "null pose output/scratch");

    if (P == 0) {
        // The reference launches nothing and returns .mean(0) of empty (0, k) tensors (rasterize_points.cu:274): NaN.
        // (0xFF bytes are a quiet NaN.)
        BAGS_CUDA_OK(cudaMemsetAsync(dL_dprojmatrix, 0xFF, 16 * sizeof(float), st));
        BAGS_CUDA_OK(cudaMemsetAsync(dL_dviewmatrix, 0xFF, 16 * sizeof(float), st));
        BAGS_CUDA_OK(cudaMemsetAsync(dL_dcampos, 0xFF, 3 * sizeof(float), st));
        BAGS_CUDA_OK(cudaMemsetAsync(dL_dshift, 0xFF, 3 * sizeof(float), st));
        return 0;
    }
    char* sp = (char*)align_up((size_t)scratch);
    float* gacc = (float*)sp;
    double* pose_acc = (double*)(sp + align_up((size_t)P * GACC_STRIDE * sizeof(float)));
    prof_begin(BAGS_STAGE_BWD_CLEAR, st);
    BAGS_CUDA_OK(cudaMemsetAsync(sp, 0, align_up((size_t)P * GACC_STRIDE * sizeof(float)) + 256, st));
    prof_end(BAGS_STAGE_BWD_CLEAR, st);

    BwdOut o;
    o.dL_dmean2D = dL_dmean2D; o.dL_dmean2D_densify = dL_dmean2D_densify; o.covariance = covariance;
    o.dL_dopacity = dL_dopacity; o.dL_dcolor = dL_dcolor; o.dL_dmean3D = dL_dmean3D; o.dL_dcov3D = dL_dcov3D;
    o.dL_dsh = dL_dsh; o.dL_dscale = dL_dscale; o.dL_drot = dL_drot; o.dL_dprojmatrix = dL_dprojmatrix;
    o.dL_dviewmatrix = dL_dviewmatrix; o.dL_dcampos = dL_dcampos; o.dL_dshift = dL_dshift;

    BwdArgs a;
    a.P = P; a.D = D; a.M = M; a.W = width; a.H = height;
    a.means3D = means3D; a.shs = shs; a.colors_precomp = colors_precomp; a.scales = scales;
    a.scale_modifier = scale_modifier; a.rotations = rotations; a.cov3D_precomp = cov3D_precomp;
    a.viewmatrix = viewmatrix; a.projmatrix = projmatrix; a.campos = campos; a.tan_fovx = tan_fovx;
    a.tan_fovy = tan_fovy;
    a.focal_y = height / (2.0f * tan_fovy);
    a.focal_x = width / (2.0f * tan_fovx);
    a.radii = radii;

    GeomState g;
    memset(&g, 0, sizeof(g));
    if (P > 0) {
        BAGS_REQUIRE(background && means3D && viewmatrix && projmatrix && campos && radii, "null required input");
        BAGS_REQUIRE(geom_buffer && binning_buffer && image_buffer, "null state buffer");
        // dL_dpix_depth / dL_dalpha may be NULL = "no gradient reaches the depth / weight image" (all zeros): what
        // autograd's undefined gradient means; the reference always receives materialised zero tensors
        BAGS_REQUIRE(dL_dpix, "null upstream colour gradient");
        BAGS_REQUIRE(dL_dmean2D && dL_dmean2D_densify && covariance && dL_dopacity && dL_dcolor && dL_dmean3D &&
                         dL_dcov3D && dL_dscale && dL_drot,
                     "null dense output");
        BAGS_REQUIRE(shs == nullptr || dL_dsh != nullptr, "null dL_dsh with SH input");
        BAGS_REQUIRE((scales != nullptr && rotations != nullptr) != (cov3D_precomp != nullptr),
                     "provide exactly one of scales+rotations / cov3D_precomp");
        g = GeomState::from(const_cast<char*>(geom_buffer), P);
        const int ntiles = tiles_x(width) * tiles_y(height);
        BinningState b = BinningState::from(const_cast<char*>(binning_buffer), R, tile_sort_passes(ntiles));
        ImageState img = ImageState::from(const_cast<char*>(image_buffer), width, height);
        if (R > 0) {
            int rc;
            {
                ProfScope ps(BAGS_STAGE_BLEND_BWD, st);
                rc = launch_blend_bwd(width, height, g, b, img, background, dL_dpix, dL_dpix_depth, dL_dalpha, gacc,
                                      st);
            }
            if (rc) return rc;
        }
    }
    int rc;
    {
        ProfScope ps(BAGS_STAGE_PREPROCESS_BWD, st);
        rc = launch_preprocess_bwd(a, g, gacc, pose_acc, o, accumulate_mask, st);
    }
    if (rc) return rc;
    if (debug) BAGS_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

extern "C" int bags_rasterize_backward(
    int P, int D, int M, int R, const float* background, int width, int height, const float* means3D,
    const float* shs, const float* colors_precomp, const float* alphas, const float* scales, float scale_modifier,
    const float* rotations, const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
    const float* intrinsic, const float* campos, float tan_fovx, float tan_fovy, const int* radii,
    const char* geom_buffer, const char* binning_buffer, const char* image_buffer, const float* dL_dpix,
    const float* dL_dpix_depth, const float* dL_dalpha, char* scratch, float* dL_dmean2D, float* dL_dmean2D_densify,
    float* covariance, float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
    float* dL_dscale, float* dL_drot, float* dL_dprojmatrix, float* dL_dviewmatrix, float* dL_dcampos,
    float* dL_dshift, int debug, void* stream) {
    return bags_rasterize_backward_accum(
        P, D, M, R, background, width, height, means3D, shs, colors_precomp, alphas, scales, scale_modifier, rotations,
        cov3D_precomp, viewmatrix, projmatrix, intrinsic, campos, tan_fovx, tan_fovy, radii, geom_buffer,
        binning_buffer, image_buffer, dL_dpix, dL_dpix_depth, dL_dalpha, scratch, dL_dmean2D, dL_dmean2D_densify,
        covariance, dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot, dL_dprojmatrix,
        dL_dviewmatrix, dL_dcampos, dL_dshift, 0u, debug, stream);
}

// ---- batched launch set (B views of the same Gaussians) -----------------------------------------------------------
static int fill_fov(int B, int width, int height, const float* tan_fovx, const float* tan_fovy, ViewFov* fov) {
    for (int v = 0; v < B; v++) {
        fov->tan_fovx[v] = tan_fovx[v];
        fov->tan_fovy[v] = tan_fovy[v];
        fov->focal_y[v] = height / (2.0f * tan_fovy[v]);
        fov->focal_x[v] = width / (2.0f * tan_fovx[v]);
    }
    for (int v = B; v < MAX_BATCH; v++) {
        fov->tan_fovx[v] = fov->tan_fovy[v] = fov->focal_x[v] = fov->focal_y[v] = 0.f;
    }
    return 0;
}

extern "C" int bags_rasterize_forward_batch(
    int B, bags_alloc_fn geom_alloc, void* geom_user, bags_alloc_fn binning_alloc, void* binning_user,
    bags_alloc_fn image_alloc, void* image_user, int P, int D, int M, const float* background, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp, const float* opacities, const float* scales,
    float scale_modifier, const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
    const float* projmatrix, const float* cam_pos, const float* shift_factors, const float* tan_fovx,
    const float* tan_fovy, int prefiltered, float* out_color, float* out_depth, float* out_alpha, int* radii,
    float* mean2Dx, float* mean2Dy, int capacity, unsigned long long* num_rendered_out, void* stream) {
    static_assert(MAX_BATCH == BAGS_MAX_BATCH, "header / kernel batch limits differ");
    cudaStream_t st = (cudaStream_t)stream;
    BAGS_REQUIRE(B >= 1 && B <= MAX_BATCH, "batch size out of range");
    BAGS_REQUIRE(P >= 0 && width > 0 && height > 0, "negative P or empty image");
    BAGS_REQUIRE(out_color && out_depth && out_alpha, "null output image");
    BAGS_REQUIRE(geom_alloc && binning_alloc && image_alloc, "null allocation callback");
    BAGS_REQUIRE(capacity >= 0 && capacity < (1 << 30), "capacity does not fit the 30-bit look-back words");
    const size_t N = (size_t)width * height;
    if (P == 0) {
        BAGS_CUDA_OK(cudaMemsetAsync(out_color, 0, B * 3 * N * sizeof(float), st));
        BAGS_CUDA_OK(cudaMemsetAsync(out_depth, 0, B * N * sizeof(float), st));
        BAGS_CUDA_OK(cudaMemsetAsync(out_alpha, 0, B * N * sizeof(float), st));
        if (num_rendered_out) BAGS_CUDA_OK(cudaMemsetAsync(num_rendered_out, 0, B * sizeof(unsigned long long), st));
        return 0;
    }
    BAGS_REQUIRE(background && means3D && opacities && viewmatrix && projmatrix && cam_pos && shift_factors &&
                     tan_fovx && tan_fovy, "null required input");
    BAGS_REQUIRE(radii && mean2Dx && mean2Dy, "null per-Gaussian output");
    BAGS_REQUIRE((shs != nullptr) != (colors_precomp != nullptr), "provide exactly one of shs / colors_precomp");
    BAGS_REQUIRE((scales != nullptr && rotations != nullptr) != (cov3D_precomp != nullptr),
                 "provide exactly one of scales+rotations / cov3D_precomp");
    BAGS_REQUIRE(shs == nullptr || (M >= (D + 1) * (D + 1) && D >= 0 && D <= 3), "SH degree / coefficient count");
    BAGS_REQUIRE(tiles_x(width) <= 65535 && tiles_y(height) <= 65535, "image too large for 16-bit tile coordinates");

    ViewStride vs;
    vs.geom = GeomState::bytes(P);
    vs.bin = BinningState::bytes(capacity);
    vs.img = ImageState::bytes(width, height);
    char* geom_ptr = geom_alloc(B * vs.geom, geom_user);
    char* img_ptr = image_alloc(B * vs.img, image_user);
    char* bin_ptr = binning_alloc(B * vs.bin, binning_user);
    if (!geom_ptr || !img_ptr || !bin_ptr) {
        set_error("allocation callback returned NULL");
        return BAGS_E_ALLOC;
    }
    BAGS_REQUIRE((((size_t)geom_ptr | (size_t)img_ptr | (size_t)bin_ptr) & 255) == 0,
                 "batched state buffers must be 256-byte aligned");
    GeomState g = GeomState::from(geom_ptr, P);
    ImageState img = ImageState::from(img_ptr, width, height);
    const int ntiles = tiles_x(width) * tiles_y(height);
    const int npass = tile_sort_passes(ntiles);
    BinningState b = BinningState::from(bin_ptr, capacity, npass);

    ViewFov fov;
    fill_fov(B, width, height, tan_fovx, tan_fovy, &fov);
    FwdArgs a;
    a.P = P; a.D = D; a.M = M; a.W = width; a.H = height;
    a.background = background; a.means3D = means3D; a.shs = shs; a.colors_precomp = colors_precomp;
    a.opacities = opacities; a.scales = scales; a.scale_modifier = scale_modifier; a.rotations = rotations;
    a.cov3D_precomp = cov3D_precomp; a.viewmatrix = viewmatrix; a.projmatrix = projmatrix; a.cam_pos = cam_pos;
    a.shift_factors = shift_factors; a.tan_fovx = tan_fovx[0]; a.tan_fovy = tan_fovy[0];
    a.focal_x = fov.focal_x[0]; a.focal_y = fov.focal_y[0];
    a.prefiltered = prefiltered;

    int rc;
    // counters + emit totals + depth-sort bookkeeping of every view: one strided memset
    BAGS_CUDA_OK(cudaMemset2DAsync(g.counters, vs.geom, 0, ((char*)g.sort.ghist - (char*)g.counters) + g.sort.zero_bytes,
                                   (size_t)B, st));
    {
        ProfScope ps(BAGS_STAGE_PREPROCESS, st);
        rc = launch_preprocess_fwd(a, g, radii, mean2Dx, mean2Dy, st, B, &fov, vs.geom);
    }
    if (rc) return rc;
    if (num_rendered_out)
        BAGS_CUDA_OK(cudaMemcpy2DAsync(num_rendered_out, sizeof(unsigned long long), g.counters, vs.geom,
                                       sizeof(unsigned long long), (size_t)B, cudaMemcpyDefault, st));
    u32 *sorted_keys, *sorted_idx;
    {
        ProfScope ps(BAGS_STAGE_DEPTH_SORT, st);
        rc = launch_sort_pairs(P, nullptr, g.key_a, g.idx_a, g.key_b, g.idx_b, 0, 32, g.sort, true, SORT_PREPARED, st,
                               &sorted_keys, &sorted_idx, B, vs.geom, 0);
    }
    if (rc) return rc;
    u32* offsets = (sorted_keys == g.key_a) ? g.key_b : g.key_a;
    {
        ProfScope ps(BAGS_STAGE_EMIT, st);
        rc = launch_emit_scan(P, g, sorted_idx, offsets, st, B, vs);
    }
    if (rc) return rc;
    const int R = capacity;
    const unsigned long long* R_dev = g.counters;
    if (R > 0) {
        const int hist_passes = npass <= 2 ? npass : 0;
        {
            ProfScope ps(BAGS_STAGE_EMIT, st);
            BAGS_CUDA_OK(cudaMemset2DAsync(b.sort.ghist, vs.bin, 0, b.sort.zero_bytes, (size_t)B, st));
            rc = launch_emit_fill(P, R, R_dev, g, sorted_idx, offsets, b, tiles_x(width), hist_passes, st, B, vs);
        }
        if (rc) return rc;
        u32 *fk, *fv;
        {
            ProfScope ps(BAGS_STAGE_TILE_SORT, st);
            rc = launch_sort_pairs(R, R_dev, b.tile_buf[0], b.id_buf[0], b.tile_buf[1], b.id_buf[1], 0,
                                   tile_sort_bits(ntiles), b.sort, false,
                                   hist_passes > 0 ? SORT_HIST_DONE : SORT_PREPARED, st, &fk, &fv, B, vs.bin, vs.geom);
        }
        if (rc) return rc;
        if (fv != b.point_list || fk != b.tile_sorted) {
            set_error("internal: tile sort ping-pong parity mismatch");
            return BAGS_E_INVALID;
        }
    }
    {
        ProfScope ps(BAGS_STAGE_RANGES, st);
        rc = launch_ranges(R, R_dev, ntiles, b.tile_sorted, img.ranges, st, B, vs);
    }
    if (rc) return rc;
    {
        ProfScope ps(BAGS_STAGE_BLEND_FWD, st);
        rc = launch_blend_fwd(width, height, g, b, img, background, out_color, out_depth, out_alpha, st, B, vs);
    }
    return rc;
}

static inline size_t gacc_view_bytes(int P) { return align_up((size_t)(P < 0 ? 0 : P) * GACC_STRIDE * sizeof(float)); }

extern "C" size_t bags_backward_batch_scratch_bytes(int B, int P) {
    if (B < 1) B = 1;
    return (size_t)B * gacc_view_bytes(P) + align_up((size_t)B * POSE_ACC_DOUBLES * sizeof(double)) + 256;
}

extern "C" int bags_rasterize_backward_batch_blend(int B, int P, int R, const float* background, int width, int height,
                                                   const char* geom_buffer, const char* binning_buffer,
                                                   const char* image_buffer, const float* dL_dpix,
                                                   const float* dL_dpix_depth, const float* dL_dalpha, char* scratch,
                                                   void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    BAGS_REQUIRE(B >= 1 && B <= MAX_BATCH, "batch size out of range");
    BAGS_REQUIRE(P >= 0 && R >= 0 && width > 0 && height > 0, "negative size");
    BAGS_REQUIRE(scratch != nullptr, "null scratch");
    if (P == 0) return 0;
    BAGS_REQUIRE(background && geom_buffer && binning_buffer && image_buffer && dL_dpix, "null required input");
    BAGS_REQUIRE((((size_t)geom_buffer | (size_t)binning_buffer | (size_t)image_buffer) & 255) == 0,
                 "batched state buffers must be 256-byte aligned");
    char* sp = (char*)align_up((size_t)scratch);
    const size_t gstride = gacc_view_bytes(P);
    {
        ProfScope ps(BAGS_STAGE_BWD_CLEAR, st);
        BAGS_CUDA_OK(cudaMemsetAsync(sp, 0, (size_t)B * gstride + (size_t)B * POSE_ACC_DOUBLES * sizeof(double), st));
    }
    if (R == 0) return 0;
    ViewStride vs;
    vs.geom = GeomState::bytes(P);
    vs.bin = BinningState::bytes(R);
    vs.img = ImageState::bytes(width, height);
    GeomState g = GeomState::from(const_cast<char*>(geom_buffer), P);
    const int ntiles = tiles_x(width) * tiles_y(height);
    BinningState b = BinningState::from(const_cast<char*>(binning_buffer), R, tile_sort_passes(ntiles));
    ImageState img = ImageState::from(const_cast<char*>(image_buffer), width, height);
    ProfScope ps(BAGS_STAGE_BLEND_BWD, st);
    return launch_blend_bwd(width, height, g, b, img, background, dL_dpix, dL_dpix_depth, dL_dalpha, (float*)sp, st, B,
                            vs, gstride);
}

extern "C" int bags_rasterize_backward_batch_gaussians(
    int B, int P, int D, int M, int width, int height, const float* means3D, const float* shs, const float* scales,
    float scale_modifier, const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
    const float* projmatrix, const float* campos, const float* tan_fovx, const float* tan_fovy, const int* radii,
    const char* geom_buffer, const char* scratch, float* dL_dmean2D, float* dL_dmean2D_densify, float* covariance,
    float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale,
    float* dL_drot, float* dL_dprojmatrix, float* dL_dviewmatrix, float* dL_dcampos, float* dL_dshift,
    unsigned accumulate_mask, int first, int count, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    BAGS_REQUIRE(B >= 1 && B <= MAX_BATCH, "batch size out of range");
    BAGS_REQUIRE(P >= 0 && width > 0 && height > 0, "negative size");
    BAGS_REQUIRE(first >= 0 && count >= 0 && (long long)first + count <= P, "slice outside [0, P)");
    BAGS_REQUIRE((accumulate_mask & ~(BAGS_ACC_MEANS3D | BAGS_ACC_SH | BAGS_ACC_OPACITY | BAGS_ACC_SCALE |
                                      BAGS_ACC_ROT)) == 0, "unknown bit in accumulate_mask");
    BAGS_REQUIRE(scratch && dL_dprojmatrix && dL_dviewmatrix && dL_dcampos && dL_dshift, "null pose output/scratch");
    if (P == 0) {
        // .mean(0) of empty tensors, like the single-view backward
        BAGS_CUDA_OK(cudaMemsetAsync(dL_dprojmatrix, 0xFF, (size_t)B * 16 * sizeof(float), st));
        BAGS_CUDA_OK(cudaMemsetAsync(dL_dviewmatrix, 0xFF, (size_t)B * 16 * sizeof(float), st));
        BAGS_CUDA_OK(cudaMemsetAsync(dL_dcampos, 0xFF, (size_t)B * 3 * sizeof(float), st));
        BAGS_CUDA_OK(cudaMemsetAsync(dL_dshift, 0xFF, (size_t)B * 3 * sizeof(float), st));
        return 0;
    }
    BAGS_REQUIRE(means3D && viewmatrix && projmatrix && campos && tan_fovx && tan_fovy && radii && geom_buffer,
                 "null required input");
    BAGS_REQUIRE(dL_dmean2D && dL_dmean2D_densify && dL_dopacity && dL_dmean3D && dL_dscale && dL_drot,
                 "null dense output");
    BAGS_REQUIRE(shs == nullptr || dL_dsh != nullptr, "null dL_dsh with SH input");
    BAGS_REQUIRE((scales != nullptr && rotations != nullptr) != (cov3D_precomp != nullptr),
                 "provide exactly one of scales+rotations / cov3D_precomp");
    BAGS_REQUIRE(((size_t)geom_buffer & 255) == 0, "batched state buffers must be 256-byte aligned");
    const char* sp = (const char*)align_up((size_t)scratch);
    const size_t gstride = gacc_view_bytes(P);
    const float* gacc = (const float*)sp;
    double* pose_acc = (double*)(sp + (size_t)B * gstride);
    ViewFov fov;
    fill_fov(B, width, height, tan_fovx, tan_fovy, &fov);
    BwdArgs a;
    a.P = P; a.D = D; a.M = M; a.W = width; a.H = height;
    a.means3D = means3D; a.shs = shs; a.colors_precomp = nullptr; a.scales = scales;
    a.scale_modifier = scale_modifier; a.rotations = rotations; a.cov3D_precomp = cov3D_precomp;
    a.viewmatrix = viewmatrix; a.projmatrix = projmatrix; a.campos = campos; a.tan_fovx = tan_fovx[0];
    a.tan_fovy = tan_fovy[0]; a.focal_x = fov.focal_x[0]; a.focal_y = fov.focal_y[0];
    a.radii = radii;
    BwdOut po;
    memset(&po, 0, sizeof(po));
    po.dL_dprojmatrix = dL_dprojmatrix; po.dL_dviewmatrix = dL_dviewmatrix; po.dL_dcampos = dL_dcampos;
    po.dL_dshift = dL_dshift;
    BwdBatchOut o;
    o.dL_dmean2D = dL_dmean2D; o.dL_dmean2D_densify = dL_dmean2D_densify; o.covariance = covariance;
    o.dL_dopacity = dL_dopacity; o.dL_dcolor = dL_dcolor; o.dL_dmean3D = dL_dmean3D; o.dL_dcov3D = dL_dcov3D;
    o.dL_dsh = dL_dsh; o.dL_dscale = dL_dscale; o.dL_drot = dL_drot;
    GeomState g = GeomState::from(const_cast<char*>(geom_buffer), P);
    ProfScope ps(BAGS_STAGE_PREPROCESS_BWD, st);
    return launch_preprocess_bwd_batch(a, B, fov, g, GeomState::bytes(P), gacc, gstride, pose_acc, po, o,
                                       accumulate_mask, first, count, first + count == P, st);
}

extern "C" int bags_unpack_state(int P, int R, int width, int height, const char* geom_buffer,
                                 const char* binning_buffer, const char* image_buffer, uint32_t* tiles_touched,
                                 uint32_t* depth_bits, float* xy, float* conic_opacity, float* rgb,
                                 uint32_t* point_list, uint64_t* point_list_keys, uint32_t* ranges,
                                 uint32_t* n_contrib, float* final_T, void* stream) {
    BAGS_REQUIRE(P > 0 && R >= 0 && geom_buffer && binning_buffer && image_buffer, "invalid argument");
    GeomState g = GeomState::from(const_cast<char*>(geom_buffer), P);
    const int ntiles = tiles_x(width) * tiles_y(height);
    BinningState b = BinningState::from(const_cast<char*>(binning_buffer), R, tile_sort_passes(ntiles));
    ImageState img = ImageState::from(const_cast<char*>(image_buffer), width, height);
    return launch_unpack(P, R, width, height, g, b, img, tiles_touched, depth_bits, xy, conic_opacity, rgb,
                         point_list, (u64*)point_list_keys, ranges, n_contrib, final_T, (cudaStream_t)stream);
}
