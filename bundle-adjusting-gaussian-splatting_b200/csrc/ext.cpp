// diff_gaussian_rasterization._C -- the thin PyTorch C++ extension above the C ABI (include/bags_b200.h).
//
// Mirrors the reference's pybind module (3dgs-pose/ext.cpp:15-20): the same four functions with the argument order
// and return tuples of RasterizeGaussiansCUDA / RasterizeGaussiansBackwardCUDA / markVisible / ComputeRelocationCUDA
// (3dgs-pose/rasterize_points.cu:36-128, 130-276, 278-297, 299-325).  What this layer does is what that file does:
// allocate the output tensors and the three growable byte buffers through the caching allocator, extract raw
// pointers, pick the current CUDA stream, call the library.  No kernel lives here; everything numerical happens
// behind the C ABI in libbags_b200.so.
//
// Extensions (no reference counterpart): rasterize_gaussians_async (caller-provided binning capacity, no host wait,
// CUDA-graph capturable), gradient sinks on the backward, unpack_state (tests), densify_* (fused compaction).
#include <torch/extension.h>

#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <tuple>

#include "../../include/bags_b200.h"

namespace {

constexpr int NUM_CHANNELS = 3;

// Growable byte buffer handed to the library as an allocation callback (resizeFunctional, rasterize_points.cu:28-34).
struct Resizable {
    torch::Tensor t;
    explicit Resizable(const torch::Device& dev)
        : t(torch::empty({0}, torch::TensorOptions().dtype(torch::kUInt8).device(dev))) {}
    static char* alloc(size_t bytes, void* user) {
        auto* self = static_cast<Resizable*>(user);
        self->t = torch::empty({(int64_t)bytes}, self->t.options());
        return reinterpret_cast<char*>(self->t.data_ptr());
    }
};

// fp32, contiguous, on `dev`; 0-element placeholders (absent optional inputs) pass through untouched
inline torch::Tensor f32c(const torch::Tensor& t, const torch::Device& dev) {
    if (!t.defined() || t.numel() == 0) return t;
    if (t.scalar_type() == torch::kFloat32 && t.device() == dev && t.is_contiguous()) return t;
    return t.to(torch::TensorOptions().dtype(torch::kFloat32).device(dev)).contiguous();
}

template <typename T = float>
inline T* ptr(const torch::Tensor& t) {
    if (!t.defined() || t.numel() == 0) return nullptr;
    return reinterpret_cast<T*>(t.data_ptr());
}

inline void check(int rc, const char* what) {
    TORCH_CHECK(rc >= 0, what, " failed (code ", rc, "): ", bags_last_error());
}

inline void* current_stream(const torch::Device& dev) {
    return (void*)c10::cuda::getCurrentCUDAStream(dev.index()).stream();
}

using FwdTuple = std::tuple<int64_t, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor,
                            torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>;

// capacity < 0: bags_rasterize_forward (host waits for num_rendered); >= 0: bags_rasterize_forward_async.
FwdTuple forward_common(const torch::Tensor& background_, const torch::Tensor& means3D_,
                        const torch::Tensor& shift_factors_, const torch::Tensor& colors_,
                        const torch::Tensor& opacity_, const torch::Tensor& scales_, const torch::Tensor& rotations_,
                        const float scale_modifier, const torch::Tensor& cov3D_precomp_,
                        const torch::Tensor& viewmatrix_, const torch::Tensor& projmatrix_,
                        const torch::Tensor& intrinsic_, const float tan_fovx, const float tan_fovy,
                        const int image_height, const int image_width, const torch::Tensor& sh_, const int degree,
                        const torch::Tensor& campos_, const bool prefiltered, const bool debug, const int64_t capacity,
                        const torch::Tensor* count_out) {
    if (means3D_.ndimension() != 2 || means3D_.size(1) != 3) AT_ERROR("means3D must have dimensions (num_points, 3)");
    TORCH_CHECK(means3D_.is_cuda(), "bags_b200: means3D must be a CUDA tensor (there is no CPU path)");
    const torch::Device dev = means3D_.device();
    c10::cuda::CUDAGuard guard(dev);
    const int P = (int)means3D_.size(0);
    const int H = image_height, W = image_width;

    const torch::Tensor means3D = f32c(means3D_, dev), background = f32c(background_, dev),
                        shift_factors = f32c(shift_factors_, dev), colors = f32c(colors_, dev),
                        opacity = f32c(opacity_, dev), scales = f32c(scales_, dev), rotations = f32c(rotations_, dev),
                        cov3D_precomp = f32c(cov3D_precomp_, dev), viewmatrix = f32c(viewmatrix_, dev),
                        projmatrix = f32c(projmatrix_, dev), intrinsic = f32c(intrinsic_, dev), sh = f32c(sh_, dev),
                        campos = f32c(campos_, dev);

    const auto f32 = torch::TensorOptions().dtype(torch::kFloat32).device(dev);
    // every element of every output is written by the kernels: empty(), not the reference's zeros()
    torch::Tensor out_color = torch::empty({NUM_CHANNELS, H, W}, f32);
    torch::Tensor out_depth = torch::empty({1, H, W}, f32);
    torch::Tensor out_alpha = torch::empty({1, H, W}, f32);
    torch::Tensor radii = torch::empty({P}, f32.dtype(torch::kInt32));
    torch::Tensor means2Dx = torch::empty({P, 1}, f32);
    torch::Tensor means2Dy = torch::empty({P, 1}, f32);
    Resizable geom(dev), binning(dev), img(dev);

    const int M = (sh.defined() && sh.numel() != 0) ? (int)sh.size(1) : 0;
    void* stream = current_stream(dev);
    int64_t rendered = 0;
    if (capacity < 0) {
        const int rc = bags_rasterize_forward(
            Resizable::alloc, &geom, Resizable::alloc, &binning, Resizable::alloc, &img, P, degree, M,
            ptr(background), W, H, ptr(means3D), ptr(sh), ptr(colors), ptr(opacity), ptr(scales), scale_modifier,
            ptr(rotations), ptr(cov3D_precomp), ptr(viewmatrix), ptr(projmatrix), ptr(intrinsic), ptr(campos),
            ptr(shift_factors), tan_fovx, tan_fovy, prefiltered ? 1 : 0, ptr(out_color), ptr(out_depth),
            ptr(out_alpha), ptr<int>(radii), ptr(means2Dx), ptr(means2Dy), debug ? 1 : 0, stream);
        check(rc, "bags_rasterize_forward");
        rendered = rc;
    } else {
        unsigned long long* cnt = nullptr;
        if (count_out != nullptr && count_out->defined() && count_out->numel() > 0) {
            TORCH_CHECK(count_out->scalar_type() == torch::kInt64 && count_out->is_contiguous(),
                        "count_out must be a contiguous int64 tensor (CUDA or pinned host memory)");
            TORCH_CHECK(count_out->is_cuda() || count_out->is_pinned(), "count_out must be CUDA or pinned host memory");
            cnt = reinterpret_cast<unsigned long long*>(count_out->data_ptr());
        }
        const int rc = bags_rasterize_forward_async(
            Resizable::alloc, &geom, Resizable::alloc, &binning, Resizable::alloc, &img, P, degree, M,
            ptr(background), W, H, ptr(means3D), ptr(sh), ptr(colors), ptr(opacity), ptr(scales), scale_modifier,
            ptr(rotations), ptr(cov3D_precomp), ptr(viewmatrix), ptr(projmatrix), ptr(intrinsic), ptr(campos),
            ptr(shift_factors), tan_fovx, tan_fovy, prefiltered ? 1 : 0, ptr(out_color), ptr(out_depth),
            ptr(out_alpha), ptr<int>(radii), ptr(means2Dx), ptr(means2Dy), (int)capacity, cnt, stream);
        check(rc, "bags_rasterize_forward_async");
        rendered = capacity;     // what the backward must be given as R: the slots of the binning buffer
    }
    return std::make_tuple(rendered, out_color, out_depth, out_alpha, radii, geom.t, binning.t, img.t, means2Dx,
                           means2Dy);
}

// RasterizeGaussiansCUDA, rasterize_points.cu:36-128
FwdTuple rasterize_gaussians(const torch::Tensor& background, const torch::Tensor& means3D,
                             const torch::Tensor& shift_factors, const torch::Tensor& colors,
                             const torch::Tensor& opacity, const torch::Tensor& scales,
                             const torch::Tensor& rotations, const float scale_modifier,
                             const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix,
                             const torch::Tensor& projmatrix, const torch::Tensor& intrinsic, const float tan_fovx,
                             const float tan_fovy, const int image_height, const int image_width,
                             const torch::Tensor& sh, const int degree, const torch::Tensor& campos,
                             const bool prefiltered, const bool debug) {
    return forward_common(background, means3D, shift_factors, colors, opacity, scales, rotations, scale_modifier,
                          cov3D_precomp, viewmatrix, projmatrix, intrinsic, tan_fovx, tan_fovy, image_height,
                          image_width, sh, degree, campos, prefiltered, debug, -1, nullptr);
}

FwdTuple rasterize_gaussians_async(const torch::Tensor& background, const torch::Tensor& means3D,
                                   const torch::Tensor& shift_factors, const torch::Tensor& colors,
                                   const torch::Tensor& opacity, const torch::Tensor& scales,
                                   const torch::Tensor& rotations, const float scale_modifier,
                                   const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix,
                                   const torch::Tensor& projmatrix, const torch::Tensor& intrinsic,
                                   const float tan_fovx, const float tan_fovy, const int image_height,
                                   const int image_width, const torch::Tensor& sh, const int degree,
                                   const torch::Tensor& campos, const bool prefiltered, const int64_t capacity,
                                   const torch::Tensor& count_out) {
    TORCH_CHECK(capacity >= 0, "capacity must be >= 0");
    return forward_common(background, means3D, shift_factors, colors, opacity, scales, rotations, scale_modifier,
                          cov3D_precomp, viewmatrix, projmatrix, intrinsic, tan_fovx, tan_fovy, image_height,
                          image_width, sh, degree, campos, prefiltered, false, capacity, &count_out);
}

using BwdTuple = std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor,
                            torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor,
                            torch::Tensor, torch::Tensor>;

// RasterizeGaussiansBackwardCUDA, rasterize_points.cu:130-276.  The sink_* arguments are the gradient-sink extension
// (bags_rasterize_backward_accum): a defined tensor receives `+= gradient` for the visible rows and the tuple slot of
// that gradient comes back undefined (None in Python).
BwdTuple rasterize_gaussians_backward(
    const torch::Tensor& background_, const torch::Tensor& means3D_, const torch::Tensor& radii_,
    const torch::Tensor& colors_, const torch::Tensor& scales_, const torch::Tensor& rotations_,
    const float scale_modifier, const torch::Tensor& cov3D_precomp_, const torch::Tensor& viewmatrix_,
    const torch::Tensor& projmatrix_, const torch::Tensor& intrinsic_, const float tan_fovx, const float tan_fovy,
    const int image_height, const int image_width, const torch::Tensor& dL_dout_color_,
    const c10::optional<torch::Tensor>& dL_dout_depth_, const c10::optional<torch::Tensor>& dL_dout_alpha_,
    const torch::Tensor& sh_, const int degree, const torch::Tensor& campos_, const torch::Tensor& geomBuffer,
    const int64_t R, const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer,
    const torch::Tensor& alphas, const bool debug, const c10::optional<torch::Tensor>& sink_means3D,
    const c10::optional<torch::Tensor>& sink_sh, const c10::optional<torch::Tensor>& sink_opacity,
    const c10::optional<torch::Tensor>& sink_scales, const c10::optional<torch::Tensor>& sink_rotations) {
    (void)image_height; (void)image_width;
    const torch::Device dev = means3D_.device();
    TORCH_CHECK(means3D_.is_cuda(), "bags_b200: means3D must be a CUDA tensor (there is no CPU path)");
    c10::cuda::CUDAGuard guard(dev);
    const int P = (int)means3D_.size(0);
    const int H = (int)dL_dout_color_.size(1), W = (int)dL_dout_color_.size(2);

    const torch::Tensor means3D = f32c(means3D_, dev), background = f32c(background_, dev),
                        colors = f32c(colors_, dev), scales = f32c(scales_, dev), rotations = f32c(rotations_, dev),
                        cov3D_precomp = f32c(cov3D_precomp_, dev), viewmatrix = f32c(viewmatrix_, dev),
                        projmatrix = f32c(projmatrix_, dev), intrinsic = f32c(intrinsic_, dev), sh = f32c(sh_, dev),
                        campos = f32c(campos_, dev), dL_dout_color = f32c(dL_dout_color_, dev);
    // absent (None) depth / weight gradients reach the kernel as NULL = zeros
    const torch::Tensor dL_dout_depth = dL_dout_depth_.has_value() ? f32c(*dL_dout_depth_, dev) : torch::Tensor();
    const torch::Tensor dL_dout_alpha = dL_dout_alpha_.has_value() ? f32c(*dL_dout_alpha_, dev) : torch::Tensor();
    const torch::Tensor radii = radii_.contiguous();

    const int M = (sh.defined() && sh.numel() != 0) ? (int)sh.size(1) : 0;
    const auto f32 = torch::TensorOptions().dtype(torch::kFloat32).device(dev);
    unsigned acc = 0;
    auto out = [&](const c10::optional<torch::Tensor>& sink, unsigned bit, std::vector<int64_t> shape,
                   const char* name, bool& is_sink) -> torch::Tensor {
        if (!sink.has_value() || !sink->defined()) {
            is_sink = false;
            return torch::empty(shape, f32);
        }
        const torch::Tensor& b = *sink;
        TORCH_CHECK(b.is_cuda() && b.device() == dev && b.scalar_type() == torch::kFloat32 && b.is_contiguous() &&
                        b.sizes().vec() == shape,
                    "bags_b200: gradient sink for ", name, " must be a contiguous fp32 CUDA tensor of shape ", shape,
                    " on ", dev);
        acc |= bit;
        is_sink = true;
        return b;
    };
    bool s_means, s_op, s_sh, s_sc, s_rot;
    torch::Tensor dL_dmeans3D = out(sink_means3D, BAGS_ACC_MEANS3D, {P, 3}, "means3D", s_means);
    torch::Tensor dL_dmeans2D = torch::empty({P, 3}, f32);
    torch::Tensor dL_dmeans2D_densify = torch::empty({P, 3}, f32);
    torch::Tensor dL_dcolors = torch::empty({P, NUM_CHANNELS}, f32);
    torch::Tensor covariance = torch::empty({P, 1}, f32);
    torch::Tensor dL_dopacity = out(sink_opacity, BAGS_ACC_OPACITY, {P, 1}, "opacity", s_op);
    torch::Tensor dL_dcov3D = torch::empty({P, 6}, f32);
    torch::Tensor dL_dsh = out(sink_sh, BAGS_ACC_SH, {P, M, 3}, "sh", s_sh);
    torch::Tensor dL_dscales = out(sink_scales, BAGS_ACC_SCALE, {P, 3}, "scales", s_sc);
    torch::Tensor dL_drotations = out(sink_rotations, BAGS_ACC_ROT, {P, 4}, "rotations", s_rot);
    torch::Tensor pose = torch::empty({38}, f32);   // proj 16 | view 16 | campos 3 | shift 3
    torch::Tensor scratch = torch::empty({(int64_t)bags_backward_scratch_bytes(P)},
                                         torch::TensorOptions().dtype(torch::kUInt8).device(dev));
    float* pp = ptr(pose);
    const int rc = bags_rasterize_backward_accum(
        P, degree, M, (int)R, ptr(background), W, H, ptr(means3D), ptr(sh), ptr(colors), ptr(alphas), ptr(scales),
        scale_modifier, ptr(rotations), ptr(cov3D_precomp), ptr(viewmatrix), ptr(projmatrix), ptr(intrinsic),
        ptr(campos), tan_fovx, tan_fovy, ptr<int>(radii), ptr<char>(geomBuffer), ptr<char>(binningBuffer),
        ptr<char>(imageBuffer), ptr(dL_dout_color), ptr(dL_dout_depth), ptr(dL_dout_alpha), ptr<char>(scratch),
        ptr(dL_dmeans2D), ptr(dL_dmeans2D_densify), ptr(covariance), ptr(dL_dopacity), ptr(dL_dcolors),
        ptr(dL_dmeans3D), ptr(dL_dcov3D), ptr(dL_dsh), ptr(dL_dscales), ptr(dL_drotations), pp, pp + 16, pp + 32,
        pp + 35, acc, debug ? 1 : 0, current_stream(dev));
    check(rc, "bags_rasterize_backward_accum");
    const torch::Tensor none;
    return std::make_tuple(dL_dmeans2D, dL_dmeans2D_densify, dL_dcolors, s_op ? none : dL_dopacity,
                           s_means ? none : dL_dmeans3D, dL_dcov3D, s_sh ? none : dL_dsh, s_sc ? none : dL_dscales,
                           s_rot ? none : dL_drotations, pose.slice(0, 32, 35), pose.slice(0, 0, 16),
                           pose.slice(0, 16, 32), pose.slice(0, 35, 38), covariance);
}

// ---- B views of the same Gaussians as one launch set (bags_rasterize_forward_batch / _backward_batch_*) -----------
// viewmatrix / projmatrix: (B, 4, 4); campos / shift_factors: (B, 3); tan_fovx / tan_fovy: B floats.  Returns
// (color (B,3,H,W), depth (B,1,H,W), alpha (B,1,H,W), radii (B,P), geom, binning, image buffers, means2Dx, means2Dy (B,P,1)).
using FwdBatchTuple = std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor,
                                 torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>;
FwdBatchTuple rasterize_gaussians_batch(const torch::Tensor& background_, const torch::Tensor& means3D_,
                                        const torch::Tensor& shift_factors_, const torch::Tensor& colors_,
                                        const torch::Tensor& opacity_, const torch::Tensor& scales_,
                                        const torch::Tensor& rotations_, const float scale_modifier,
                                        const torch::Tensor& cov3D_precomp_, const torch::Tensor& viewmatrix_,
                                        const torch::Tensor& projmatrix_, const std::vector<float>& tan_fovx,
                                        const std::vector<float>& tan_fovy, const int image_height,
                                        const int image_width, const torch::Tensor& sh_, const int degree,
                                        const torch::Tensor& campos_, const bool prefiltered, const int64_t capacity,
                                        const torch::Tensor& count_out) {
    if (means3D_.ndimension() != 2 || means3D_.size(1) != 3) AT_ERROR("means3D must have dimensions (num_points, 3)");
    TORCH_CHECK(means3D_.is_cuda(), "bags_b200: means3D must be a CUDA tensor (there is no CPU path)");
    const torch::Device dev = means3D_.device();
    c10::cuda::CUDAGuard guard(dev);
    const int P = (int)means3D_.size(0);
    const int H = image_height, W = image_width;
    const int64_t B = (int64_t)tan_fovx.size();
    TORCH_CHECK(B >= 1 && B <= BAGS_MAX_BATCH, "bags_b200: a batch holds 1..", BAGS_MAX_BATCH, " views, got ", B);
    TORCH_CHECK((int64_t)tan_fovy.size() == B && viewmatrix_.numel() == 16 * B && projmatrix_.numel() == 16 * B &&
                    campos_.numel() == 3 * B && shift_factors_.numel() == 3 * B,
                "bags_b200: per-view inputs must be stacked: viewmatrix/projmatrix (B,4,4), campos/shift_factors (B,3)");
    TORCH_CHECK(capacity >= 0, "capacity must be >= 0");

    const torch::Tensor means3D = f32c(means3D_, dev), background = f32c(background_, dev),
                        shift_factors = f32c(shift_factors_, dev), colors = f32c(colors_, dev),
                        opacity = f32c(opacity_, dev), scales = f32c(scales_, dev), rotations = f32c(rotations_, dev),
                        cov3D_precomp = f32c(cov3D_precomp_, dev), viewmatrix = f32c(viewmatrix_, dev),
                        projmatrix = f32c(projmatrix_, dev), sh = f32c(sh_, dev), campos = f32c(campos_, dev);
    const auto f32 = torch::TensorOptions().dtype(torch::kFloat32).device(dev);
    torch::Tensor out_color = torch::empty({B, NUM_CHANNELS, H, W}, f32);
    torch::Tensor out_depth = torch::empty({B, 1, H, W}, f32);
    torch::Tensor out_alpha = torch::empty({B, 1, H, W}, f32);
    torch::Tensor radii = torch::empty({B, P}, f32.dtype(torch::kInt32));
    torch::Tensor means2Dx = torch::empty({B, P, 1}, f32);
    torch::Tensor means2Dy = torch::empty({B, P, 1}, f32);
    Resizable geom(dev), binning(dev), img(dev);
    const int M = (sh.defined() && sh.numel() != 0) ? (int)sh.size(1) : 0;
    unsigned long long* cnt = nullptr;
    if (count_out.defined() && count_out.numel() > 0) {
        TORCH_CHECK(count_out.scalar_type() == torch::kInt64 && count_out.is_contiguous() && count_out.numel() >= B,
                    "count_out must be a contiguous int64 tensor with one element per view");
        TORCH_CHECK(count_out.is_cuda() || count_out.is_pinned(), "count_out must be CUDA or pinned host memory");
        cnt = reinterpret_cast<unsigned long long*>(count_out.data_ptr());
    }
    const int rc = bags_rasterize_forward_batch(
        (int)B, Resizable::alloc, &geom, Resizable::alloc, &binning, Resizable::alloc, &img, P, degree, M,
        ptr(background), W, H, ptr(means3D), ptr(sh), ptr(colors), ptr(opacity), ptr(scales), scale_modifier,
        ptr(rotations), ptr(cov3D_precomp), ptr(viewmatrix), ptr(projmatrix), ptr(campos), ptr(shift_factors),
        tan_fovx.data(), tan_fovy.data(), prefiltered ? 1 : 0, ptr(out_color), ptr(out_depth), ptr(out_alpha),
        ptr<int>(radii), ptr(means2Dx), ptr(means2Dy), (int)capacity, cnt, current_stream(dev));
    check(rc, "bags_rasterize_forward_batch");
    return std::make_tuple(out_color, out_depth, out_alpha, radii, geom.t, binning.t, img.t, means2Dx, means2Dy);
}

// Step 1 of the batched backward: blend backward of all views.  Returns the scratch buffer step 2 reads.
torch::Tensor rasterize_gaussians_backward_batch_blend(const torch::Tensor& background_, const int64_t P,
                                                       const int64_t R, const torch::Tensor& dL_dout_color_,
                                                       const c10::optional<torch::Tensor>& dL_dout_depth_,
                                                       const c10::optional<torch::Tensor>& dL_dout_alpha_,
                                                       const torch::Tensor& geomBuffer,
                                                       const torch::Tensor& binningBuffer,
                                                       const torch::Tensor& imageBuffer) {
    const torch::Device dev = geomBuffer.device();
    TORCH_CHECK(geomBuffer.is_cuda(), "bags_b200: CUDA tensors only (there is no CPU path)");
    c10::cuda::CUDAGuard guard(dev);
    TORCH_CHECK(dL_dout_color_.dim() == 4 && dL_dout_color_.size(1) == NUM_CHANNELS,
                "dL_dout_color must have dimensions (B, 3, H, W)");
    const int B = (int)dL_dout_color_.size(0), H = (int)dL_dout_color_.size(2), W = (int)dL_dout_color_.size(3);
    const torch::Tensor background = f32c(background_, dev), dL_dout_color = f32c(dL_dout_color_, dev);
    const torch::Tensor dL_dout_depth = dL_dout_depth_.has_value() ? f32c(*dL_dout_depth_, dev) : torch::Tensor();
    const torch::Tensor dL_dout_alpha = dL_dout_alpha_.has_value() ? f32c(*dL_dout_alpha_, dev) : torch::Tensor();
    torch::Tensor scratch = torch::empty({(int64_t)bags_backward_batch_scratch_bytes(B, (int)P)},
                                         torch::TensorOptions().dtype(torch::kUInt8).device(dev));
    check(bags_rasterize_backward_batch_blend(B, (int)P, (int)R, ptr(background), W, H, ptr<char>(geomBuffer),
                                              ptr<char>(binningBuffer), ptr<char>(imageBuffer), ptr(dL_dout_color),
                                              ptr(dL_dout_depth), ptr(dL_dout_alpha), ptr<char>(scratch),
                                              current_stream(dev)),
          "bags_rasterize_backward_batch_blend");
    return scratch;
}

// Step 2: Gaussians [first, first + count) of all views, into caller-provided tensors (so that consecutive slices fill
// the same tensors): per view dL_dmeans2D / dL_dmeans2D_densify (B,P,3), pose (38 B floats) = proj (B,16) | view (B,16) |
// campos (B,3) | shift (B,3); summed over the views dL_dmeans3D (P,3), dL_dopacity (P,1), dL_dsh (P,M,3), dL_dscales (P,3), dL_drotations
// (P,4), optional dL_dcolors (P,3), dL_dcov3D (P,6), covariance (B,P,1).  acc_mask: BAGS_ACC_* bits of the summed outputs
// that are gradient sinks (added onto instead of written).
void rasterize_gaussians_backward_batch_gaussians(
    const torch::Tensor& means3D_, const torch::Tensor& radii_, const torch::Tensor& scales_,
    const torch::Tensor& rotations_, const float scale_modifier, const torch::Tensor& cov3D_precomp_,
    const torch::Tensor& viewmatrix_, const torch::Tensor& projmatrix_, const std::vector<float>& tan_fovx,
    const std::vector<float>& tan_fovy, const int image_height, const int image_width, const torch::Tensor& sh_,
    const int degree, const torch::Tensor& campos_, const torch::Tensor& geomBuffer, const torch::Tensor& scratch,
    torch::Tensor dL_dmeans2D, torch::Tensor dL_dmeans2D_densify, torch::Tensor dL_dopacity,
    torch::Tensor dL_dmeans3D, torch::Tensor dL_dsh, torch::Tensor dL_dscales, torch::Tensor dL_drotations,
    torch::Tensor pose, const c10::optional<torch::Tensor>& dL_dcolors, const c10::optional<torch::Tensor>& dL_dcov3D,
    const c10::optional<torch::Tensor>& covariance, const int64_t acc_mask, const int64_t first,
    const int64_t count) {
    const torch::Device dev = means3D_.device();
    TORCH_CHECK(means3D_.is_cuda(), "bags_b200: means3D must be a CUDA tensor (there is no CPU path)");
    c10::cuda::CUDAGuard guard(dev);
    const int P = (int)means3D_.size(0);
    const int64_t B = (int64_t)tan_fovx.size();
    const torch::Tensor means3D = f32c(means3D_, dev), scales = f32c(scales_, dev), rotations = f32c(rotations_, dev),
                        cov3D_precomp = f32c(cov3D_precomp_, dev), viewmatrix = f32c(viewmatrix_, dev),
                        projmatrix = f32c(projmatrix_, dev), sh = f32c(sh_, dev), campos = f32c(campos_, dev);
    const torch::Tensor radii = radii_.contiguous();
    const int M = (sh.defined() && sh.numel() != 0) ? (int)sh.size(1) : 0;
    auto need = [&](const torch::Tensor& t, int64_t numel, const char* name) {
        TORCH_CHECK(t.defined() && t.is_cuda() && t.device() == dev && t.scalar_type() == torch::kFloat32 &&
                        t.is_contiguous() && t.numel() == numel,
                    "bags_b200: output ", name, " must be a contiguous fp32 CUDA tensor with ", numel, " elements");
    };
    need(dL_dmeans2D, B * P * 3, "dL_dmeans2D");
    need(dL_dmeans2D_densify, B * P * 3, "dL_dmeans2D_densify");
    need(dL_dopacity, P, "dL_dopacity");
    need(dL_dmeans3D, (int64_t)P * 3, "dL_dmeans3D");
    if (M > 0) need(dL_dsh, (int64_t)P * M * 3, "dL_dsh");
    need(dL_dscales, (int64_t)P * 3, "dL_dscales");
    need(dL_drotations, (int64_t)P * 4, "dL_drotations");
    need(pose, B * 38, "pose");
    float* col = nullptr; float* cov3 = nullptr; float* covar = nullptr;
    if (dL_dcolors.has_value() && dL_dcolors->defined()) { need(*dL_dcolors, (int64_t)P * 3, "dL_dcolors"); col = ptr(*dL_dcolors); }
    if (dL_dcov3D.has_value() && dL_dcov3D->defined()) { need(*dL_dcov3D, (int64_t)P * 6, "dL_dcov3D"); cov3 = ptr(*dL_dcov3D); }
    if (covariance.has_value() && covariance->defined()) { need(*covariance, B * P, "covariance"); covar = ptr(*covariance); }
    float* pp = ptr(pose);
    check(bags_rasterize_backward_batch_gaussians(
              (int)B, P, degree, M, image_width, image_height, ptr(means3D), ptr(sh), ptr(scales), scale_modifier,
              ptr(rotations), ptr(cov3D_precomp), ptr(viewmatrix), ptr(projmatrix), ptr(campos), tan_fovx.data(),
              tan_fovy.data(), ptr<int>(radii), ptr<char>(geomBuffer), ptr<char>(scratch), ptr(dL_dmeans2D),
              ptr(dL_dmeans2D_densify), covar, ptr(dL_dopacity), col, ptr(dL_dmeans3D), cov3, ptr(dL_dsh),
              ptr(dL_dscales), ptr(dL_drotations), pp, pp + 16 * B, pp + 32 * B, pp + 35 * B, (unsigned)acc_mask,
              (int)first, (int)count, current_stream(dev)),
          "bags_rasterize_backward_batch_gaussians");
}

// markVisible, rasterize_points.cu:278-297
torch::Tensor mark_visible(const torch::Tensor& means3D_, const torch::Tensor& viewmatrix_,
                           const torch::Tensor& projmatrix_) {
    const torch::Device dev = means3D_.device();
    TORCH_CHECK(means3D_.is_cuda(), "bags_b200: means3D must be a CUDA tensor (there is no CPU path)");
    c10::cuda::CUDAGuard guard(dev);
    const int P = (int)means3D_.size(0);
    torch::Tensor present = torch::zeros({P}, torch::TensorOptions().dtype(torch::kBool).device(dev));
    if (P != 0) {
        const torch::Tensor means3D = f32c(means3D_, dev), viewmatrix = f32c(viewmatrix_, dev),
                            projmatrix = f32c(projmatrix_, dev);
        check(bags_mark_visible(P, ptr(means3D), ptr(viewmatrix), ptr(projmatrix), ptr<uint8_t>(present),
                                current_stream(dev)),
              "bags_mark_visible");
    }
    return present;
}

// ComputeRelocationCUDA, rasterize_points.cu:299-325
std::tuple<torch::Tensor, torch::Tensor> compute_relocation(const torch::Tensor& opacity_old_,
                                                            const torch::Tensor& scale_old_, const torch::Tensor& N_,
                                                            const torch::Tensor& binoms_, const int n_max) {
    const torch::Device dev = opacity_old_.device();
    TORCH_CHECK(opacity_old_.is_cuda(), "bags_b200: opacity_old must be a CUDA tensor (there is no CPU path)");
    c10::cuda::CUDAGuard guard(dev);
    const int P = (int)opacity_old_.size(0);
    const auto f32 = torch::TensorOptions().dtype(torch::kFloat32).device(dev);
    torch::Tensor final_opacity = torch::zeros({P}, f32);
    torch::Tensor final_scale = torch::zeros({3 * P}, f32);
    if (P != 0) {
        const torch::Tensor opacity_old = f32c(opacity_old_, dev), scale_old = f32c(scale_old_, dev),
                            binoms = f32c(binoms_, dev);
        const torch::Tensor N = N_.to(torch::TensorOptions().dtype(torch::kInt32).device(dev)).contiguous();
        check(bags_compute_relocation(P, ptr(opacity_old), ptr(scale_old), ptr<int>(N), ptr(binoms), n_max,
                                      ptr(final_opacity), ptr(final_scale), current_stream(dev)),
              "bags_compute_relocation");
    }
    return std::make_tuple(final_opacity, final_scale);
}

// Test-only: the private forward state as a dict of tensors (bags_unpack_state).
py::dict unpack_state(const int P, const int64_t R, const int H, const int W, const torch::Tensor& geomBuffer,
                      const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer) {
    const torch::Device dev = geomBuffer.device();
    c10::cuda::CUDAGuard guard(dev);
    const int64_t tiles = (int64_t)((W + 15) / 16) * ((H + 15) / 16);
    const auto i32 = torch::TensorOptions().dtype(torch::kInt32).device(dev);
    const auto f32 = i32.dtype(torch::kFloat32);
    torch::Tensor tiles_touched = torch::empty({P}, i32), depth_bits = torch::empty({P}, i32),
                  xy = torch::empty({P, 2}, f32), conic_opacity = torch::empty({P, 4}, f32),
                  rgb = torch::empty({P, 3}, f32), point_list = torch::empty({R}, i32),
                  point_list_keys = torch::empty({R}, i32.dtype(torch::kInt64)), ranges = torch::empty({tiles, 2}, i32),
                  n_contrib = torch::empty({(int64_t)H * W}, i32), final_T = torch::empty({(int64_t)H * W}, f32);
    check(bags_unpack_state(P, (int)R, W, H, ptr<char>(geomBuffer), ptr<char>(binningBuffer), ptr<char>(imageBuffer),
                            ptr<uint32_t>(tiles_touched), ptr<uint32_t>(depth_bits), ptr(xy), ptr(conic_opacity),
                            ptr(rgb), ptr<uint32_t>(point_list), ptr<uint64_t>(point_list_keys),
                            ptr<uint32_t>(ranges), ptr<uint32_t>(n_contrib), ptr(final_T), current_stream(dev)),
          "bags_unpack_state");
    py::dict d;
    d["tiles_touched"] = tiles_touched;
    d["depth_bits"] = depth_bits;
    d["xy"] = xy;
    d["conic_opacity"] = conic_opacity;
    d["rgb"] = rgb;
    d["point_list"] = point_list;
    d["point_list_keys"] = point_list_keys;
    d["ranges"] = ranges;
    d["n_contrib"] = n_contrib;
    d["final_T"] = final_T;
    return d;
}

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "B200-native drop-in for the reference's diff_gaussian_rasterization._C (libbags_b200.so behind it)";
    // the library calls never touch Python objects: the GIL is released so that a blocking forward (it waits for
    // num_rendered like the reference's cudaMemcpy) does not stall other Python threads
    m.def("rasterize_gaussians", &rasterize_gaussians, py::call_guard<py::gil_scoped_release>());
    m.def("rasterize_gaussians_backward", &rasterize_gaussians_backward, py::arg("background"), py::arg("means3D"),
          py::arg("radii"), py::arg("colors"), py::arg("scales"), py::arg("rotations"), py::arg("scale_modifier"),
          py::arg("cov3D_precomp"), py::arg("viewmatrix"), py::arg("projmatrix"), py::arg("intrinsic"),
          py::arg("tan_fovx"), py::arg("tan_fovy"), py::arg("image_height"), py::arg("image_width"),
          py::arg("dL_dout_color"), py::arg("dL_dout_depth"), py::arg("dL_dout_alpha"), py::arg("sh"),
          py::arg("degree"), py::arg("campos"), py::arg("geomBuffer"), py::arg("R"), py::arg("binningBuffer"),
          py::arg("imageBuffer"), py::arg("alphas"), py::arg("debug"), py::arg("sink_means3D") = py::none(),
          py::arg("sink_sh") = py::none(), py::arg("sink_opacity") = py::none(), py::arg("sink_scales") = py::none(),
          py::arg("sink_rotations") = py::none(), py::call_guard<py::gil_scoped_release>());
    m.def("mark_visible", &mark_visible, py::call_guard<py::gil_scoped_release>());
    m.def("compute_relocation", &compute_relocation, py::call_guard<py::gil_scoped_release>());
    m.def("rasterize_gaussians_async", &rasterize_gaussians_async, py::call_guard<py::gil_scoped_release>());
    m.def("unpack_state", &unpack_state);
    m.def("rasterize_gaussians_batch", &rasterize_gaussians_batch, py::call_guard<py::gil_scoped_release>());
    m.def("rasterize_gaussians_backward_batch_blend", &rasterize_gaussians_backward_batch_blend,
          py::call_guard<py::gil_scoped_release>());
    m.def("rasterize_gaussians_backward_batch_gaussians", &rasterize_gaussians_backward_batch_gaussians,
          py::call_guard<py::gil_scoped_release>());
    m.attr("MAX_BATCH") = (int)BAGS_MAX_BATCH;
    m.attr("ACC_MEANS3D") = (int)BAGS_ACC_MEANS3D;
    m.attr("ACC_SH") = (int)BAGS_ACC_SH;
    m.attr("ACC_OPACITY") = (int)BAGS_ACC_OPACITY;
    m.attr("ACC_SCALE") = (int)BAGS_ACC_SCALE;
    m.attr("ACC_ROT") = (int)BAGS_ACC_ROT;
}
