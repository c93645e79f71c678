/*
 * bags_b200.h -- C ABI of the B200-native (sm_100a) bundle-adjusting Gaussian-splatting rasterizer.
 *
 * This is the drop-in boundary for the reference's native hot path.  Every entry point takes plain device
 * pointers, sizes and a cudaStream_t (passed as void*); there are no torch / pybind types in any signature.
 * Each function names the reference interface it replaces (paths relative to the reference checkout):
 *
 *   bags_rasterize_forward        <- CudaRasterizer::Rasterizer::forward   3dgs-pose/cuda_rasterizer/rasterizer.h:31-60
 *                                    (called from RasterizeGaussiansCUDA,  3dgs-pose/rasterize_points.cu:36-128)
 *   bags_rasterize_backward       <- CudaRasterizer::Rasterizer::backward  3dgs-pose/cuda_rasterizer/rasterizer.h:62-102
 *                                    + the four .mean(0) reductions and "covariance + 1e-4" of
 *                                    RasterizeGaussiansBackwardCUDA,        3dgs-pose/rasterize_points.cu:130-276
 *   bags_mark_visible             <- CudaRasterizer::Rasterizer::markVisible rasterizer.h:24-29, rasterize_points.cu:278-297
 *   bags_compute_relocation       <- UTILS::ComputeRelocation              3dgs-pose/cuda_rasterizer/utils.h, utils.cu:6-51
 *   bags_knn_mean_dist2           <- SimpleKNN::knn                        simple-knn/simple_knn.cu:185-221 (distCUDA2, spatial.cu:14-26)
 *   bags_distort_forward/backward <- utils/util_distortion.py:271-311 (apply_distortion: F.interpolate of the control-
 *                                    point flow + F.grid_sample of the render) and :58-77 (center_crop) -- the post-step
 *                                    that --outside_rasterizer runs on every rendered canvas
 *   bags_loss_forward/backward    <- utils/loss_utils.py:18-19 (l1_loss) and :49-76 (ssim / _ssim: five grouped 11x11
 *                                    Gaussian convolutions + the SSIM map), combined at train.py:325
 *   bags_densify_gather           <- GaussianModel.densify_and_prune's torch.cat / masked-index surgery over the parameters
 *                                    and Adam moments, scene/gaussian_model.py:315-447
 *   bags_adam_step                <- torch.optim.Adam.step() as scene/gaussian_model.py:197-206 / train.py:420 use it
 *   bags_rasterize_forward_batch / bags_rasterize_backward_batch_* <- B calls of the two functions above for B cameras
 *                                    over the same Gaussians (the per-face render() loop of utils/cubemap_utils.py:229,265)
 *   bags_allreduce_p2p            (no reference counterpart: the reference is single-GPU) the gradient exchange of
 *                                    view-parallel training as one kernel over NVLink peer memory
 *   bags_rasterize_forward_async  <- the same forward without the blocking read-back of num_rendered
 *                                    (rasterizer_impl.cu:291): caller-provided binning capacity, count kept on the device
 *   bags_unpack_state             (test-only) exposes the private binning/image state so parity tests can compare
 *                                  tiles_touched / sorted keys / point list / ranges / n_contrib / final_T bit for bit
 *                                  with the reference's GeometryState/BinningState/ImageState (rasterizer_impl.h:21-72).
 *
 * All matrices use the reference's convention: 16 floats, flat index m[4*c + r] (the transposed, row-vector form that
 * scene/cameras.py hands to GaussianRasterizationSettings).  All tensors are fp32 and contiguous unless stated.
 * Return value: >= 0 on success, a negative BAGS_E_* code on failure (bags_last_error() gives the text).
 *
 * Threading: every entry point may be called concurrently from several host threads (any devices, any streams); the
 * library keeps no per-call state outside the caller's buffers (the pinned word + event a synchronous forward needs
 * come from a locked pool).  bags_last_error() is per thread.  The tuning knobs and the profiling switch are process
 * wide and meant to be set while nothing else runs.
 */
#ifndef BAGS_B200_H_INCLUDED
#define BAGS_B200_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BAGS_E_INVALID   (-1)   /* bad argument (null pointer, negative size, unsupported mode) */
#define BAGS_E_CUDA      (-2)   /* a CUDA runtime call or kernel launch failed                   */
#define BAGS_E_ALLOC     (-3)   /* an allocation callback returned NULL                          */
#define BAGS_E_OVERFLOW  (-4)   /* tile-instance count does not fit the 30-bit sort bookkeeping  */

/* Scratch-buffer allocation callback: must return a device pointer to at least `bytes` bytes, 256-byte aligned,
 * valid until the matching backward call.  Mirrors std::function<char*(size_t)> of rasterizer.h:32-34. */
typedef char* (*bags_alloc_fn)(size_t bytes, void* user);

const char* bags_last_error(void);
int bags_version(void);

/* Sizes of the opaque state buffers (bytes).  geom/image depend only on P / W*H; binning needs num_rendered. */
size_t bags_geom_bytes(int P);
size_t bags_image_bytes(int width, int height);
size_t bags_binning_bytes(int P, int num_rendered, int width, int height);
size_t bags_backward_scratch_bytes(int P);

/* Forward: preprocess -> depth sort -> scan -> emit (tile,id) -> tile sort -> ranges -> blend.
 * Optional inputs are selected by NULL exactly like the reference (shs vs colors_precomp, scales+rotations vs
 * cov3D_precomp).  Outputs: out_color [3,H,W], out_depth [H,W], out_alpha [H,W], radii [P] (int32),
 * mean2Dx/mean2Dy [P] (= p_hom.z / shifted p_hom.w, forward.cu:650-651).  Every output element is written
 * (no pre-zeroing needed).  Returns num_rendered. */
int bags_rasterize_forward(
    bags_alloc_fn geom_alloc, void* geom_user,
    bags_alloc_fn binning_alloc, void* binning_user,
    bags_alloc_fn image_alloc, void* image_user,
    int P, int D, int M,
    const float* background,
    int width, int height,
    const float* means3D,
    const float* shs,
    const float* colors_precomp,
    const float* opacities,
    const float* scales,
    float scale_modifier,
    const float* rotations,
    const float* cov3D_precomp,
    const float* viewmatrix,
    const float* projmatrix,
    const float* intrinsic,
    const float* cam_pos,
    const float* shift_factors,
    float tan_fovx, float tan_fovy,
    int prefiltered,
    float* out_color,
    float* out_depth,
    float* out_alpha,
    int* radii,
    float* mean2Dx,
    float* mean2Dy,
    int debug,
    void* stream);

/* Asynchronous forward (extension; the reference blocks on a 4-byte device-to-host copy of num_rendered in every
 * forward, rasterizer_impl.cu:291, to size its binning buffer).  Same contract as bags_rasterize_forward except:
 *   - nothing waits on the device and no host memory is allocated: the call only enqueues work on `stream`, so it can be
 *     captured in a CUDA graph and the host can run any number of views ahead of the GPU;
 *   - `capacity` is the number of tile instances the binning buffer is sized for (binning_alloc is asked for
 *     bags_binning_bytes(P, capacity, ...) bytes; pass the same number as `R` to the backward).  The grids of the emit /
 *     tile-sort / range kernels are sized by it and the kernels read the actual count on the device;
 *   - *num_rendered_out (device or pinned host memory, may be NULL) receives the actual count, in stream order.
 *     A count above `capacity` means the tile lists were truncated: the images and every gradient of that view are
 *     WRONG (nothing is written out of bounds) and the caller must re-render with a larger capacity.  A trainer checks
 *     the count together with the loss, one step late; the count of the previous iteration plus some slack is the
 *     natural capacity.
 * Returns 0 on success. */
int bags_rasterize_forward_async(
    bags_alloc_fn geom_alloc, void* geom_user,
    bags_alloc_fn binning_alloc, void* binning_user,
    bags_alloc_fn image_alloc, void* image_user,
    int P, int D, int M,
    const float* background,
    int width, int height,
    const float* means3D,
    const float* shs,
    const float* colors_precomp,
    const float* opacities,
    const float* scales,
    float scale_modifier,
    const float* rotations,
    const float* cov3D_precomp,
    const float* viewmatrix,
    const float* projmatrix,
    const float* intrinsic,
    const float* cam_pos,
    const float* shift_factors,
    float tan_fovx, float tan_fovy,
    int prefiltered,
    float* out_color,
    float* out_depth,
    float* out_alpha,
    int* radii,
    float* mean2Dx,
    float* mean2Dy,
    int capacity,
    unsigned long long* num_rendered_out,
    void* stream);

/* Backward.  Dense per-Gaussian outputs ([P,...], every row written, zero for culled Gaussians):
 *   dL_dmean2D [P,3], dL_dmean2D_densify [P,3], dL_dcolor [P,3], dL_dopacity [P], dL_dmean3D [P,3],
 *   dL_dcov3D [P,6], dL_dsh [P,M,3] (may be NULL when M == 0), dL_dscale [P,3], dL_drot [P,4], covariance [P]
 *   (= accumulated conic trace + 1e-4, rasterize_points.cu:274).
 * Pose / intrinsic outputs, already reduced as mean over all P rows like the reference's .mean(0):
 *   dL_dcampos [3], dL_dprojmatrix [16], dL_dviewmatrix [16], dL_dshift [3].
 * `scratch` must hold bags_backward_scratch_bytes(P) bytes (contents undefined on entry).
 * dL_dpix_depth and / or dL_dalpha may be NULL: no gradient reaches the depth / weight image (treated as zeros; the
 * reference always receives materialised zero tensors from autograd). */
int bags_rasterize_backward(
    int P, int D, int M, int R,
    const float* background,
    int width, int height,
    const float* means3D,
    const float* shs,
    const float* colors_precomp,
    const float* alphas,
    const float* scales,
    float scale_modifier,
    const float* rotations,
    const float* cov3D_precomp,
    const float* viewmatrix,
    const float* projmatrix,
    const float* intrinsic,
    const float* campos,
    float tan_fovx, float tan_fovy,
    const int* radii,
    const char* geom_buffer,
    const char* binning_buffer,
    const char* image_buffer,
    const float* dL_dpix,
    const float* dL_dpix_depth,
    const float* dL_dalpha,
    char* scratch,
    float* dL_dmean2D,
    float* dL_dmean2D_densify,
    float* covariance,
    float* dL_dopacity,
    float* dL_dcolor,
    float* dL_dmean3D,
    float* dL_dcov3D,
    float* dL_dsh,
    float* dL_dscale,
    float* dL_drot,
    float* dL_dprojmatrix,
    float* dL_dviewmatrix,
    float* dL_dcampos,
    float* dL_dshift,
    int debug,
    void* stream);

/* Gradient accumulation (extension; the reference returns fresh tensors and leaves `.grad += new` to autograd, three
 * extra passes over 236 B/Gaussian per view).  Same contract as bags_rasterize_backward, except that every output
 * selected in `accumulate_mask` (BAGS_ACC_*) is ADDED to what the caller's buffer already holds: rows of visible
 * Gaussians are updated with one fp32 round-to-nearest add per element (bit-identical to `buffer += fresh gradient`),
 * rows of culled Gaussians are neither read nor written.  accumulate_mask == 0 is bags_rasterize_backward. */
#define BAGS_ACC_MEANS3D 1u
#define BAGS_ACC_SH 2u
#define BAGS_ACC_OPACITY 4u
#define BAGS_ACC_SCALE 8u
#define BAGS_ACC_ROT 16u
int bags_rasterize_backward_accum(
    int P, int D, int M, int R,
    const float* background,
    int width, int height,
    const float* means3D,
    const float* shs,
    const float* colors_precomp,
    const float* alphas,
    const float* scales,
    float scale_modifier,
    const float* rotations,
    const float* cov3D_precomp,
    const float* viewmatrix,
    const float* projmatrix,
    const float* intrinsic,
    const float* campos,
    float tan_fovx, float tan_fovy,
    const int* radii,
    const char* geom_buffer,
    const char* binning_buffer,
    const char* image_buffer,
    const float* dL_dpix,
    const float* dL_dpix_depth,
    const float* dL_dalpha,
    char* scratch,
    float* dL_dmean2D,
    float* dL_dmean2D_densify,
    float* covariance,
    float* dL_dopacity,
    float* dL_dcolor,
    float* dL_dmean3D,
    float* dL_dcov3D,
    float* dL_dsh,
    float* dL_dscale,
    float* dL_drot,
    float* dL_dprojmatrix,
    float* dL_dviewmatrix,
    float* dL_dcampos,
    float* dL_dshift,
    unsigned accumulate_mask,
    int debug,
    void* stream);

/* ---- B views of the same Gaussians as ONE launch set (extension) -------------------------
 * The reference renders several cameras over the same Gaussians per iteration -- the five cubemap faces of
 * utils/cubemap_utils.py:229,265 each go through gaussian_renderer.render() -> one full rasterizer forward, and each
 * view's backward re-reads every parameter and re-adds every gradient.  Here the B views are one batch:
 *   - every kernel of the pipeline is launched once for all views (the view is a grid dimension), so the latency-bound
 *     stages (depth sort of 1 M keys, emit, ranges) fill the GPU and the blend kernels' tails overlap;
 *   - the forward preprocess reads mean / scale / rotation / SH once per Gaussian and evaluates the 3D covariance once;
 *   - the backward preprocess reads the SH row once, sums the Gaussian-parameter gradients of the B views on chip and
 *     writes (or, with accumulate_mask, adds) them once;
 *   - nothing waits on the host (the asynchronous protocol of bags_rasterize_forward_async: one `capacity` per view).
 * Per view the results are those of the single-view entry points bit for bit (images, radii, every private state
 * array); summed gradients equal buffer + g_0 + g_1 + ... accumulated in view order.
 *
 * Layout: per-view inputs are stacked -- viewmatrix / projmatrix [B][16], cam_pos / shift_factors [B][3]; tan_fovx /
 * tan_fovy are HOST arrays [B].  Per-view outputs are stacked -- out_color [B][3][H][W], out_depth / out_alpha
 * [B][H][W], radii / mean2Dx / mean2Dy [B][P], num_rendered_out [B].  The three state buffers hold B consecutive
 * single-view states: geom_alloc is asked for B * bags_geom_bytes(P) bytes (binning: B * bags_binning_bytes(P, capacity,
 * ..), image: B * bags_image_bytes(..)) and must return 256-byte aligned memory; view v's state starts v * bytes in and
 * is a valid single-view state (bags_unpack_state / bags_rasterize_backward accept it).  1 <= B <= BAGS_MAX_BATCH, all
 * views share width, height, background, scale_modifier and the SH degree.  Returns 0 on success. */
#define BAGS_MAX_BATCH 8
int bags_rasterize_forward_batch(
    int B,
    bags_alloc_fn geom_alloc, void* geom_user,
    bags_alloc_fn binning_alloc, void* binning_user,
    bags_alloc_fn image_alloc, void* image_user,
    int P, int D, int M,
    const float* background,
    int width, int height,
    const float* means3D,
    const float* shs,
    const float* colors_precomp,
    const float* opacities,
    const float* scales,
    float scale_modifier,
    const float* rotations,
    const float* cov3D_precomp,
    const float* viewmatrix,
    const float* projmatrix,
    const float* cam_pos,
    const float* shift_factors,
    const float* tan_fovx, const float* tan_fovy,
    int prefiltered,
    float* out_color,
    float* out_depth,
    float* out_alpha,
    int* radii,
    float* mean2Dx,
    float* mean2Dy,
    int capacity,
    unsigned long long* num_rendered_out,
    void* stream);

/* Backward of a batch, in two steps so that a multi-GPU caller can overlap the gradient exchange with the computation:
 *   1. bags_rasterize_backward_batch_blend: zeroes `scratch` (bags_backward_batch_scratch_bytes(B, P) bytes) and runs the
 *      blend backward of all B views (one grid) into it.  dL_dpix [B][3][H][W]; dL_dpix_depth / dL_dalpha [B][H][W] or NULL.
 *   2. bags_rasterize_backward_batch_gaussians: the per-Gaussian chain rule for Gaussians [first, first + count) of ALL
 *      views.  Call it once with (0, P), or slice by slice in ascending order; when a slice has been enqueued, the rows
 *      [first, first + count) of the summed outputs are final -- a trainer records an event and starts the all-reduce of
 *      that slice on another stream while the next slice computes.  The call whose slice ends at P also completes the
 *      per-view pose / intrinsic gradients.
 * Outputs.  Per view: dL_dmean2D, dL_dmean2D_densify [B][P][3]; covariance [B][P] (may be NULL); dL_dprojmatrix,
 * dL_dviewmatrix [B][16]; dL_dcampos, dL_dshift [B][3].  Summed over the views: dL_dmean3D [P][3], dL_dopacity [P],
 * dL_dsh [P][M][3], dL_dscale [P][3], dL_drot [P][4], and -- may be NULL -- dL_dcolor [P][3], dL_dcov3D [P][6].
 * accumulate_mask (BAGS_ACC_*): the selected summed outputs are added onto the caller's buffer (rows no view sees are
 * neither read nor written) instead of written. */
size_t bags_backward_batch_scratch_bytes(int B, int P);
int bags_rasterize_backward_batch_blend(
    int B, int P, int R,
    const float* background,
    int width, int height,
    const char* geom_buffer,
    const char* binning_buffer,
    const char* image_buffer,
    const float* dL_dpix,
    const float* dL_dpix_depth,
    const float* dL_dalpha,
    char* scratch,
    void* stream);
int bags_rasterize_backward_batch_gaussians(
    int B, int P, int D, int M,
    int width, int height,
    const float* means3D,
    const float* shs,
    const float* scales,
    float scale_modifier,
    const float* rotations,
    const float* cov3D_precomp,
    const float* viewmatrix,
    const float* projmatrix,
    const float* campos,
    const float* tan_fovx, const float* tan_fovy,
    const int* radii,
    const char* geom_buffer,
    const char* scratch,
    float* dL_dmean2D,
    float* dL_dmean2D_densify,
    float* covariance,
    float* dL_dopacity,
    float* dL_dcolor,
    float* dL_dmean3D,
    float* dL_dcov3D,
    float* dL_dsh,
    float* dL_dscale,
    float* dL_drot,
    float* dL_dprojmatrix,
    float* dL_dviewmatrix,
    float* dL_dcampos,
    float* dL_dshift,
    unsigned accumulate_mask,
    int first, int count,
    void* stream);

/* In-place all_reduce(SUM) of an fp32 buffer over NVLink peer memory, one process per GPU of one NVSwitch node (SURVEY.md
 * section 8e: the gradient exchange of view-parallel training; the reference is single-GPU and has no counterpart).
 * buffers[p] / flags[p] (HOST arrays of `world` device pointers, p = rank): rank p's gradient buffer (n floats, 16-byte
 * aligned, n a multiple of 4) and its flag block (bags_allreduce_p2p_flag_bytes() bytes, zeroed once at allocation),
 * mapped into this process (CUDA IPC for the peers' entries).  Every rank calls it with the same world / n, its own rank,
 * on a stream ordered behind the kernels that produced its buffer.  Rank r loads the `world` contributions to the slice
 * it owns, adds them in rank order and stores the sums into all `world` buffers; two flag barriers bracket the loop.  On
 * return of the kernel every buffer holds the same sums, bit for bit.  1 <= world <= 8.  Returns 0 on success. */
size_t bags_allreduce_p2p_flag_bytes(void);
/* cudaDeviceEnablePeerAccess from the current device to `peer_device` (no error if it is enabled already). */
int bags_enable_peer_access(int peer_device);
/* Map a peer process's device allocation, given its 64-byte CUDA IPC memory handle, into this process under the current
 * device (peer access enabled by the driver); bags_ipc_close unmaps it. */
int bags_ipc_open(const void* handle64, void** out_ptr);
int bags_ipc_close(void* ptr);
int bags_allreduce_p2p(int world, int rank, float* const* buffers, uint32_t* const* flags, long long n, void* stream);

/* present[i] = (view * p_i).z > 0.2  (auxiliary.h:139-164).  `present` is one byte per point. */
int bags_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                      uint8_t* present, void* stream);

/* MCMC relocation, Eq. (9) (utils.cu:6-36).  binoms is [n_max, n_max]. */
int bags_compute_relocation(int P, const float* opacity_old, const float* scale_old, const int* N,
                            const float* binoms, int n_max, float* opacity_new, float* scale_new, void* stream);

/* Mean squared distance to the 3 nearest neighbours (simple_knn.cu:132-221).  `scratch` must hold
 * bags_knn_scratch_bytes(P) bytes. */
size_t bags_knn_scratch_bytes(int P);
int bags_knn_mean_dist2(int P, const float* points, float* mean_dist2, char* scratch, void* stream);

/* Lens-distortion post-step (utils/util_distortion.py:271-311 + 58-77), fused:
 *   flow(Hf,Wf,2) = bilinear upsample (align_corners=False) of `flow` [Hs,Ws,2]   (identity when Hs==Hf && Ws==Wf)
 *   canvas        = grid_sample(image [C,H,W], flow, bilinear, zeros, align_corners=True)          -> [C,Hf,Wf]
 *   out_image     = crop ? center_crop(canvas, Hc, Wc) (a second bilinear grid_sample) : canvas    -> [C,Hc,Wc]
 *   out_mask[Hc,Wc] = mask_mode 0: !(out[0]==0 && out[1]==0)   mask_mode 1: !(out[0]<1e-5 && out[1]<1e-5)  (may be NULL)
 *   out_flow [Hf,Wf,2] = the upsampled flow (may be NULL; apply_distortion returns it for reuse).
 * 1 <= C <= 4.  With crop == 0, Hc/Wc must equal Hf/Wf.  Neither the canvas nor the upsampled flow is materialised.
 * Backward: dL_dout [C,Hc,Wc] -> dL_dimage [C,H,W] and dL_dflow [Hs,Ws,2] (either may be NULL); both are zero-filled
 * by the call and then accumulated with atomics like ATen's grid_sampler_2d_backward / upsample_bilinear2d_backward. */
int bags_distort_forward(const float* image, int C, int H, int W, const float* flow, int Hs, int Ws, int Hf, int Wf,
                         int Hc, int Wc, int crop, int mask_mode, float* out_image, float* out_mask, float* out_flow,
                         void* stream);
int bags_distort_backward(const float* image, int C, int H, int W, const float* flow, int Hs, int Ws, int Hf, int Wf,
                          int Hc, int Wc, int crop, const float* dL_dout, float* dL_dimage, float* dL_dflow,
                          void* stream);

/* center_crop alone (utils/util_distortion.py:58-77), one image [C,H,W] -> [C,Hc,Wc]; backward writes every element
 * of dL_dimage [C,H,W] (gather form, deterministic). */
int bags_center_crop_forward(const float* image, int C, int H, int W, int Hc, int Wc, float* out, void* stream);
int bags_center_crop_backward(const float* dL_dout, int C, int H, int W, int Hc, int Wc, float* dL_dimage,
                              void* stream);

/* Photometric loss (utils/loss_utils.py:18-19, 49-76), fused.  img1/img2: [planes][H][W] (planes = N*C).
 * Forward: sums[plane] = {sum |img1-img2|, sum ssim_map} as doubles (fixed-order reduction; the caller divides by the
 * element counts to get the reference's means).  want_ssim == 0 evaluates the L1 sum only.  d_mu1/d_e11/d_e12
 * ([planes][H][W] each; all three or none) receive the partial derivatives of the SSIM map that the backward needs.
 * `scratch` must hold bags_loss_scratch_bytes(planes, H, W) bytes.
 * Backward: dL_dimg1 = coef[plane][0] * d(sum|.|)/dimg1 + coef[plane][1] * d(sum ssim_map)/dimg1, with coef a DEVICE
 * array [planes][2] (the upstream gradients never visit the host).  Pass NULL maps for an L1-only backward. */
size_t bags_loss_scratch_bytes(int planes, int H, int W);
/* The 11 fp32 taps of the 1-D Gaussian window the loss kernels use (host memory; no device work).  They must equal
 * utils/loss_utils.py:36-38 `gauss / gauss.sum()` bit for bit: a one-ulp difference in the normaliser biases the mean
 * SSIM of smooth images by 1e-5 (tests/test_host_logic.py pins them against torch). */
int bags_loss_window(float* taps11);
int bags_loss_forward(const float* img1, const float* img2, int planes, int H, int W, int want_ssim, float* d_mu1,
                      float* d_e11, float* d_e12, char* scratch, double* sums, void* stream);
int bags_loss_backward(const float* img1, const float* img2, int planes, int H, int W, const float* d_mu1,
                       const float* d_e11, const float* d_e12, const float* coef, float* dL_dimg1, void* stream);

/* l1_loss against the ground truth as the data set stores it (uint8, scaled by 1/255 like PILtoTorch's division):
 * loss[0] = mean |img - gt/255| (device fp32 scalar, fixed-order reduction) and dL_dimg = sign(img - gt/255) / n in
 * one pass over n contiguous elements.  `scratch` must hold bags_l1_u8_scratch_bytes(n) bytes. */
size_t bags_l1_u8_scratch_bytes(long long n);
int bags_l1_u8_forward(const float* img, const uint8_t* gt, long long n, float* dL_dimg, float* loss, char* scratch,
                       void* stream);

/* Fused multi-tensor Adam step.  Replaces `torch.optim.Adam.step()` as the reference
 * calls it (scene/gaussian_model.py:197-206 creates `torch.optim.Adam(l, lr=0.0, eps=1e-15)` over six parameter
 * groups; train.py:420 steps it once per iteration): no weight decay, no amsgrad, not maximize.  One launch per 8
 * tensors reads param / grad / exp_avg / exp_avg_sq once and writes param / exp_avg / exp_avg_sq once.
 * `tensors` is a HOST array; the pointers inside are device pointers to contiguous fp32 arrays of n elements.
 * step_size = -lr / (1 - beta1^t) and bias_correction2_sqrt = sqrt(1 - beta2^t) are computed by the caller in double
 * for the tensor's own step count t (already incremented), exactly like torch's Python code. */
typedef struct bags_adam_tensor {
    float* param;
    const float* grad;
    float* exp_avg;
    float* exp_avg_sq;
    long long n;
    double step_size;
    double bias_correction2_sqrt;
} bags_adam_tensor;
int bags_adam_step(int ntensors, const bags_adam_tensor* tensors, double beta1, double beta2, double eps, void* stream);

/* Fused densification compaction.  Replaces the tensor surgery of
 * GaussianModel.densify_and_prune (scene/gaussian_model.py:334-447: densify_and_clone + densify_and_split +
 * prune_points, each a torch.cat / boolean indexing of every parameter tensor and of both Adam moments) by ONE gather.
 * The caller decides, per ORIGINAL row, with the reference's own mask expressions: keep (not a split parent, not
 * pruned), clone (copy survives), child (split parent whose children survive); `flags` [P] holds them as bits 0/1/2
 * (bit 3 = split parent, needed for the sample rank), `plan` [P][4] int32 the exclusive scans of the four bits, and
 * kept / clones / children / split_parents their totals.  Output rows: [kept | clones | children r=0 | ... | r=N-1],
 * every segment in ascending source row = the order the reference's cat + masked indexing produces.
 * Per tensor (host array of descriptors, device pointers inside): rows of `width` floats; role 0 = plain copy,
 * role 1 = xyz (child: R(rotation/|rotation|) . samples[r * split_parents + rank] + xyz), role 2 = log-scaling (child:
 * log(exp(s) / (0.8 N))).  src_m/src_v (+dst) = Adam exp_avg / exp_avg_sq or NULL: copied for kept rows, zero for new
 * rows.  `rotation` [P,4] raw quaternions, `samples` [N * split_parents, 3] the caller's torch.normal draws. */
typedef struct bags_densify_tensor {
    const float* src;
    float* dst;
    const float* src_m;
    float* dst_m;
    const float* src_v;
    float* dst_v;
    int width;
    int role;
} bags_densify_tensor;
int bags_densify_gather(long long P, int N, const int32_t* plan, const uint8_t* flags, int kept, int clones,
                        int children, int split_parents, const float* rotation, const float* samples, int ntensors,
                        const bags_densify_tensor* tensors, void* stream);

/* Test-only view of the private state.  Any output pointer may be NULL.  All outputs are device pointers:
 *   tiles_touched [P] u32, depth_bits [P] u32 (float bits of the view-space depth, undefined for culled),
 *   xy [P,2] f32, conic_opacity [P,4] f32, rgb [P,3] f32,
 *   point_list [R] u32, point_list_keys [R] u64 (tile << 32 | depth bits, reconstructed; R = the slots of the binning
 *   buffer: after an asynchronous forward that is the capacity, and the slots past the actual count read 0),
 *   ranges [tiles,2] u32,
 *   n_contrib [H*W] u32, final_T [H*W] f32. */
int bags_unpack_state(int P, int R, int width, int height,
                      const char* geom_buffer, const char* binning_buffer, const char* image_buffer,
                      uint32_t* tiles_touched, uint32_t* depth_bits, float* xy, float* conic_opacity, float* rgb,
                      uint32_t* point_list, uint64_t* point_list_keys, uint32_t* ranges,
                      uint32_t* n_contrib, float* final_T, void* stream);

/* Stand-alone access to the hand-written onesweep radix sort (u32 keys, u32 values, bits [begin_bit, end_bit)),
 * used by tests and micro-benchmarks.  `scratch` must hold bags_sort_scratch_bytes(n) bytes.  keys_in/vals_in are
 * clobbered (ping-pong); the sorted result is always left in keys_out/vals_out. */
size_t bags_sort_scratch_bytes(int n);
int bags_sort_pairs_u32(int n, uint32_t* keys_in, uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                        int begin_bit, int end_bit, char* scratch, void* stream);

/* ---- measurement hooks (no reference counterpart; used by bench.py) --------------------------------------------
 * bags_launch_count: number of kernels this library has launched in this process (all devices, all streams).
 * bags_profile_enable(1): from now on every pipeline stage is bracketed by CUDA events on the launching stream.
 * bags_profile_read: synchronises the recorded events, ADDS per-stage elapsed milliseconds into ms[BAGS_STAGE_COUNT]
 * and per-stage launch counts into calls[BAGS_STAGE_COUNT] (either may be NULL), then clears the recordings. */
enum {
    BAGS_STAGE_PREPROCESS = 0, BAGS_STAGE_DEPTH_SORT = 1, BAGS_STAGE_EMIT = 2, BAGS_STAGE_TILE_SORT = 3,
    BAGS_STAGE_RANGES = 4, BAGS_STAGE_BLEND_FWD = 5, BAGS_STAGE_BWD_CLEAR = 6, BAGS_STAGE_BLEND_BWD = 7,
    BAGS_STAGE_PREPROCESS_BWD = 8, BAGS_STAGE_COUNT = 9
};
unsigned long long bags_launch_count(void);
int bags_profile_enable(int on);
int bags_profile_read(double* ms, unsigned long long* calls);

/* Tuning knobs (no reference counterpart).  BLEND_*_PPT: pixels per thread of the forward / backward blend kernel
 * (1, 2 or 4; a warp owns that many 8x4 patches of the tile).  Results do not depend on them beyond the fp32
 * summation order of the gradients.  BLEND_CULL: 1 (default) skips, per 8x4 patch, the splats whose alpha >= 1/255
 * bounding box misses the patch; 0 evaluates every splat of the tile list for every patch (must give the same forward
 * bits -- used by the tests).  SORT_LOOKBACK: predecessor status words the onesweep passes fetch per look-back
 * round (4 or 8; default 4; no effect on results).  SORT_SMALL_TILES: 1 (default) sorts 2048 instead of 4096 keys
 * per CTA when a pass would not fill the GPU twice (latency-bound sizes, e.g. the 1 M-key depth sort); SORT_HWMATCH: 1
 * ranks those small tiles with match.any instead of one vote per digit bit (default 0).  PBWD_MIN_BLOCKS: resident CTAs
 * per SM the backward preprocess kernel is compiled for (5 = 96 registers; 6 = 80 registers, default).  BLEND_STAGING: how
 * the blend kernels bring a batch of splat records into shared memory -- 0 (default): ldg -> registers -> sts, one
 * buffer; 1: cp.async (LDGSTS) into two alternating buffers, one batch ahead of the one being blended (same results;
 * measured in profiles/).  BLEND_BWD_REDUCE: how the blend backward sums a splat's 12 gradient slots over the warp -- 1
 * (default): through shared memory (45 instead of 68 instructions per splat and warp; measured 0.787 -> 0.741 ms per
 * view); 0: transposed shuffle butterfly (same sums up to fp32 summation order).
 * Returns the previous value (other values only
 * query), BAGS_E_INVALID for an unknown knob. */
enum { BAGS_TUNE_BLEND_FWD_PPT = 0, BAGS_TUNE_BLEND_BWD_PPT = 1, BAGS_TUNE_BLEND_CULL = 2,
       BAGS_TUNE_SORT_LOOKBACK = 3, BAGS_TUNE_SORT_SMALL_TILES = 4, BAGS_TUNE_SORT_HWMATCH = 5,
       BAGS_TUNE_PBWD_MIN_BLOCKS = 6, BAGS_TUNE_BLEND_STAGING = 7, BAGS_TUNE_BLEND_BWD_REDUCE = 8 };
int bags_set_tuning(int knob, int value);

#ifdef __cplusplus
}
#endif
#endif /* BAGS_B200_H_INCLUDED */
