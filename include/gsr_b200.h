/*
 * gsr_b200.h — C ABI of the B200-native differentiable Gaussian-splat rasterizer core
 * (libgsr_b200.so).
 *
 * This is the drop-in boundary of the hot path: the five operations below replace, one for
 * one, the C++ statics of the reference's CUDA core that its torch binding calls
 *
 *   gsr_light_forward   <- CudaRasterizer::Rasterizer::forward   (light)
 *                          diff-gaussian-rasterization-light/cuda_rasterizer/rasterizer.h:31-63
 *   gsr_light_backward  <- CudaRasterizer::Rasterizer::backward  (light)  rasterizer.h:65-104
 *   gsr_full_forward    <- CudaRasterizer::Rasterizer::forward   (full)
 *                          diff-gaussian-rasterization-full/cuda_rasterizer/rasterizer.h:31-59
 *   gsr_full_backward   <- CudaRasterizer::Rasterizer::backward  (full)   rasterizer.h:61-102
 *   gsr_mark_visible    <- CudaRasterizer::Rasterizer::markVisible        rasterizer.h:24-29
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to fp32 / int32 data unless stated otherwise; "absent"
 *    optional inputs (shs / colors_precomp, scales+rotations / cov3D_precomp) are NULL, exactly
 *    as the reference's kernels test them (forward.cu:205,241);
 *  - 4x4 matrices are 16 floats read column-major (m[0],m[4],m[8],m[12] = row 0), i.e. callers
 *    pass the transposed world-to-camera / full-projection matrices like Inria 3DGS does
 *    (auxiliary.h:58-77);
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); all work is
 *    enqueued on it.  The forward calls block once on that stream to read back the duplicate
 *    count that sizes the binning buffer (the reference blocks in the same place,
 *    rasterizer_impl.cu:287);
 *  - the three opaque buffers (geometry / binning / image state) are obtained through C
 *    allocator callbacks `char* (*)(void* ctx, size_t bytes)` which must return device memory
 *    aligned to at least 256 bytes that stays valid until the matching backward has run
 *    (the reference uses std::function<char*(size_t)> resize lambdas for the same purpose,
 *    rasterize_points.cu:27-33).  Their internal layout is private to this library;
 *  - rotations, dL_dconic, dL_drot (and shs / dL_dsh for the bulk-copy path, which otherwise falls
 *    back) are accessed with 128-bit loads / stores and must be 16-byte aligned; a misaligned pointer
 *    is rejected with GSR_E_INVALID (fresh torch allocations always are aligned; the torch shim
 *    re-packs offset views);
 *  - outputs need not be pre-zeroed: every output element is written by the call;
 *  - every entry point returns 0 on success and a negative GSR_E_* code on failure;
 *    gsr_last_error() gives a thread-local human-readable message.
 */
#ifndef GSR_B200_H
#define GSR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSR_B200_ABI_VERSION 5

#if defined(__GNUC__)
#define GSR_API __attribute__((visibility("default")))
#else
#define GSR_API
#endif

#define GSR_OK 0
#define GSR_E_INVALID (-1)   /* bad argument (shape / null pointer / unsupported channel count) */
#define GSR_E_CUDA (-2)      /* a CUDA runtime call or kernel failed */
#define GSR_E_ALLOC (-3)     /* an allocator callback returned NULL */

typedef char* (*gsr_alloc_fn)(void* ctx, size_t bytes);

/* ABI version of the loaded library (== GSR_B200_ABI_VERSION it was built with). */
GSR_API int gsr_abi_version(void);

/* Thread-local message describing the last failure in this thread ("" if none). */
GSR_API const char* gsr_last_error(void);

/* Number of floats of caller-provided, uninitialised device scratch a backward call needs. */
GSR_API size_t gsr_backward_scratch_floats(int P);

/* Tuning switches (not part of the reference surface).  gsr_set_option changes the PROCESS-WIDE
 * DEFAULT; every entry point copies the defaults into a thread-local snapshot when it starts and
 * uses only that copy, so a call never sees a switch change half way through and concurrent calls
 * from several threads / streams / devices do not share mutable option state.  None of the switches
 * changes a result beyond the order of floating-point additions.  All other per-call state of the
 * library (the speculative-binning size estimate, pinned read-back slots, kernel attributes) is kept
 * per (device, image size, Gaussian count) context, see "async_binning".
 *   "exact_ng"      1: gsr_full_forward returns the exact number of valid (pixel, Gaussian) pairs
 *                   like the reference (costs a second host sync per forward); 0 (default): skip
 *                   the count and the sync and return 0 — the reference's Python only hands the
 *                   value back to the backward to size scratch lists that do not exist here.
 *   "stage_timing"  1: record per-stage CUDA events (see gsr_stage_times); 0 (default): off.
 *   "tight_tiles"   1 (default): drop (tile, Gaussian) duplicates that provably cannot reach the
 *                   alpha >= 15/255 threshold inside the tile (every output is unchanged bit for
 *                   bit; num_rendered is smaller than the reference's); 0: the reference's
 *                   3-sigma rectangle rule, num_rendered identical to the reference.
 *   "tile_sort"     1 (default): tile-local binning — entries are scattered into their tile's
 *                   segment and each tile is sorted on chip (same order as the reference's
 *                   device-wide (tile | depth) radix sort; tiles with more than 8192 entries make
 *                   the frame fall back to the radix path); 0: always the radix path.
 *   "bwd_packed"    backward blend kernel: 3 = two pixels per lane, packed fp32x2 arithmetic, every
 *                   quarter warp (a 4x4 pixel block) walks its own entry list; 1 = the same arithmetic
 *                   with one list per warp (8x8 block); 0 = one pixel per lane, 8x4 blocks;
 *                   2 (default) = 3 (same results up to summation order).
 *   "fwd_packed"    forward blend kernel: 1 = two pixels per lane, packed fp32x2 arithmetic, quarter-warp
 *                   entry lists; 0 = one pixel per lane, half-warp lists; 2 (default) = 1 for -full, 0 for
 *                   -light (measured).  Bit-identical images either way.
 *   "bwd_occ"       CTAs per SM the quarter-list kernel is built for: 8 (64 registers) or 7 (72
 *                   registers, default: measured 5-9 % faster at C3 / C4).
 *   "async_binning" 1 (default): the forward sizes the binning buffer from the previous frame's
 *                   duplicate count (+25 %) and enqueues the scatter and the per-tile sort before
 *                   the host has read this frame's count, so the GPU does not idle during the
 *                   read-back (guarded kernels; the binning is redone when the estimate was too
 *                   small); 0: wait for the count first, exact buffer size.  The estimate is kept per
 *                   (device, width, height, P) context, so rasterizers of different shapes in one
 *                   process (640x480 tracking + 1080p mapping, one thread per GPU) do not disturb each
 *                   other; after a miss a context speculates again only once a frame's counts would
 *                   have fitted the previous frame's estimate.
 *   "bulk_sh"       1 (default): the per-Gaussian kernels move SH rows (M = 16 or 4, 16-byte aligned)
 *                   between global and shared memory with cp.async.bulk (TMA), one row per thread and
 *                   only for Gaussians that survive culling; 0: block-wide coalesced staging.
 *   "cnt_stride"    spacing (in 32-bit words, 1..32, default 8 = one per 32-byte sector) of the per-tile
 *                   entry counters / scatter cursors: neighbouring tiles' atomics no longer serialise on
 *                   one cache line (measured: preprocess_fwd 0.078 -> 0.069 ms, scatter 0.059 -> 0.042 ms).
 *   "track_headroom_pct" head room (percent, default 50) of the tracker's binning buffer over the
 *                   counts of its probing forward; negative values force the overflow / retry path
 *                   (test hook).
 * Returns the previous value, or GSR_E_INVALID for an unknown key. */
GSR_API int gsr_set_option(const char* key, int value);
GSR_API int gsr_get_option(const char* key);

/* Stage timing (bench / profiling aid, not part of the reference surface).
 * With gsr_set_option("stage_timing", 1) every entry point brackets each pipeline stage with
 * CUDA events recorded on the caller's stream.  gsr_stage_times() waits for the recorded events
 * and returns, per stage, the accumulated device milliseconds, the number of timed scopes and the
 * number of kernels this library launched inside them (CUB's internal kernels are counted as one
 * launch per call); `reset` != 0 clears the accumulators afterwards.  Arrays hold
 * GSR_STAGE_COUNT entries; any may be NULL. */
#define GSR_STAGE_COUNT 10
GSR_API int gsr_stage_times(double* ms, long long* scopes, long long* launches, int reset);
/* Kernels (and memset nodes) this library has launched in this process since the last reset, counted
 * on the host at launch time, independent of "stage_timing" (CUB calls count as one). */
GSR_API long long gsr_launch_count(int reset);
GSR_API const char* gsr_stage_name(int stage);

/* Optional extras of the backward calls (not part of the reference surface; pass NULL for none).
 * They serve view-level data parallelism: dL/dSH of a view is the rank-1 product
 * basis_k(view direction) x dL_dcolor_masked[c], so ranks can exchange 3 floats per Gaussian and view
 * instead of 3*M and rebuild the summed dL/dSH locally with gsr_sh_grad_from_views(). */
typedef struct gsr_backward_extras {
  float* dL_dcolor_masked; /* out [P,3]: dL/dcolor with the channels clamped at 0 by the forward zeroed
                              (0 for culled Gaussians); NULL = not wanted; needs SH colours */
  int skip_sh_grad;        /* 1: do not write dL_dsh */
  /* Densification statistics of Inria-3DGS-style mapping loops, accumulated by
   * the per-Gaussian backward kernel for every visible Gaussian (radii > 0) — what
   * add_densification_stats does with five torch kernels over [P] after every backward:
   *   densify_grad_accum[i] += |dL_dmean2D[i].xy|,  densify_denom[i] += 1,
   *   max_radii2D[i] = max(max_radii2D[i], radii[i]).
   * in/out [P] each; NULL = not wanted (grad_accum and denom only as a pair); ignored when the
   * call computes no map gradient (-light, map_off). */
  float* densify_grad_accum;
  float* densify_denom;
  float* max_radii2D;
  /* View-parallel exchange (dp.py): when non-NULL (a cudaEvent_t) and dL_dcolor_masked is wanted, the masked
   * colour gradient is written by a small kernel of its own right after the blend backward — before the
   * per-Gaussian backward kernel runs — and this event is recorded on `stream` at that point, so that the
   * caller can start gathering it from the peers underneath the per-Gaussian kernel.  NULL = written by the
   * per-Gaussian kernel as before. */
  void* masked_color_ready_event;
} gsr_backward_extras;

/* ---- light variant ---------------------------------------------------------------------- */

/* replaces L/cuda_rasterizer/rasterizer.h:31-63  (Rasterizer::forward).
 * out_color[3,H,W], out_depth/out_median_depth/out_alpha/out_depth_var[H,W], radii[P],
 * gau_uncertainty[P], gau_related_pixels[P].  *num_rendered receives the number of
 * (Gaussian, tile) duplicates (the int the reference returns). */
GSR_API int gsr_light_forward(
    gsr_alloc_fn geom_alloc, void* geom_ctx,
    gsr_alloc_fn binning_alloc, void* binning_ctx,
    gsr_alloc_fn img_alloc, void* img_ctx,
    int P, int D, int M,
    const float* background, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp,
    const float* opacities, const float* scales, float scale_modifier,
    const float* rotations, const float* cov3D_precomp,
    const float* viewmatrix, const float* projmatrix, const float* cam_pos,
    float tan_fovx, float tan_fovy, int prefiltered,
    float* out_color, float* out_depth, float* out_median_depth, float* out_alpha,
    const float* gt_depth, float* out_depth_var,
    float* gau_uncertainty, int* gau_related_pixels, int* radii,
    int debug, void* stream, int* num_rendered);

/* replaces L/cuda_rasterizer/rasterizer.h:65-104 (Rasterizer::backward).
 * dL_dmean2D[P,3], dL_dconic[P,4], dL_dopacity[P], dL_dcolor[P,3], dL_ddepth[P],
 * dL_dmean3D[P,3], dL_dcov3D[P,6], dL_dsh[P,M,3], dL_dscale[P,3], dL_drot[P,4],
 * dL_dview[16] (already summed over pixels; the reference returns [H*W,16] and sums in Python).
 * `scratch` = gsr_backward_scratch_floats(P) floats. */
GSR_API int gsr_light_backward(
    int P, int D, int M, int R,
    const float* background, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp,
    const float* alphas, const float* scales, float scale_modifier,
    const float* rotations, const float* cov3D_precomp,
    const float* viewmatrix, const float* projmatrix, const float* cam_pos,
    float tan_fovx, float tan_fovy, const int* radii,
    char* geom_buffer, char* binning_buffer, char* img_buffer,
    const float* dL_dpix, const float* dL_dpix_depth,
    const float* dL_dpix_median_depth, const float* dL_dpix_depth_var,
    float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
    float* dL_ddepth, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
    float* dL_dscale, float* dL_drot,
    int debug, const float* perspec_matrix, float* dL_dview,
    const float* gt_depth, int track_off, int map_off,
    float* scratch, void* stream, const gsr_backward_extras* extras);

/* ---- full variant ----------------------------------------------------------------------- */

/* replaces F/cuda_rasterizer/rasterizer.h:31-59.  out_uncertainty is the accumulated
 * alpha*T ("opacity_map").  *num_related receives the number of valid (pixel, Gaussian)
 * pairs (the second int of the reference's tuple). */
GSR_API int gsr_full_forward(
    gsr_alloc_fn geom_alloc, void* geom_ctx,
    gsr_alloc_fn binning_alloc, void* binning_ctx,
    gsr_alloc_fn img_alloc, void* img_ctx,
    int P, int D, int M,
    const float* background, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp,
    const float* opacities, const float* scales, float scale_modifier,
    const float* rotations, const float* cov3D_precomp,
    const float* viewmatrix, const float* projmatrix, const float* cam_pos,
    float tan_fovx, float tan_fovy, int prefiltered,
    float* out_color, float* out_depth, float* out_uncertainty, int* radii,
    const float* gt_depth, void* stream, int* num_rendered, int* num_related);

/* replaces F/cuda_rasterizer/rasterizer.h:61-102.  Same gradient outputs as the light
 * variant; dL_dview[16].  The reference's per-pair scratch lists (dpixel_dgc, gau_id_list,
 * pix_id_list, dpixel_dndcs, dpixel_dinvcovs, ddepth_dndcs, ddepth_dinvcovs) and per-Gaussian
 * Jacobian tables do not exist here: the pose gradient is reduced per Gaussian on chip. */
GSR_API int gsr_full_backward(
    int P, int D, int M, int R,
    const float* background, int width, int height,
    const float* means3D, const float* shs, const float* colors_precomp,
    const float* scales, float scale_modifier,
    const float* rotations, const float* cov3D_precomp,
    const float* viewmatrix, const float* projmatrix, const float* cam_pos,
    float tan_fovx, float tan_fovy, const int* radii,
    char* geom_buffer, char* binning_buffer, char* img_buffer,
    const float* dL_dpix, const float* dL_dpix_depth, const float* dL_dpix_uncertainty,
    float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
    float* dL_ddepth, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
    float* dL_dscale, float* dL_drot,
    const float* perspec_matrix, float* dL_dview,
    const float* gt_depth, float* scratch, void* stream, const gsr_backward_extras* extras);

/* Sum over `nviews` camera views of dL/dSH, rebuilt from each view's masked colour gradient:
 *   dL_dsh[g][k][c] = sum_v basis_k(normalize(means3D[g] - campos_v)) * dR_v[g][c],  k < (D+1)^2,
 * zero for k >= (D+1)^2.  dR_v = dR_all + v * view_stride (floats), campos_v = campos_all +
 * v * campos_stride (floats).  Same value as adding the per-view dL_dsh outputs of the backward
 * calls (up to summation order).  Not part of the reference surface. */
GSR_API int gsr_sh_grad_from_views(int P, int D, int M, const float* means3D, int nviews,
                                   const float* dR_all, size_t view_stride,
                                   const float* campos_all, size_t campos_stride,
                                   float* dL_dsh, void* stream);

/* The same sum with the views' gradients read IN PLACE (at most 16 views): dR_ptrs[v] -> [P,3] masked colour
 * gradient of view v, campos_ptrs[v] -> its camera position [3]; the pointers (HOST arrays of DEVICE
 * pointers) may point into peer GPUs' memory (symmetric memory / P2P mappings): the all-gather of
 * view-level data parallelism then happens as the P2P loads of this kernel. */
GSR_API int gsr_sh_grad_from_view_ptrs(int P, int D, int M, const float* means3D, int nviews,
                                       const float* const* dR_ptrs, const float* const* campos_ptrs,
                                       float* dL_dsh, void* stream);

/* In-switch (NVLS) all-reduce of this rank's 1/world slice of count_floats floats starting
 * offset_floats into a symmetric buffer, given the buffer's MULTICAST alias: multimem.ld_reduce sums
 * the replicas in the NVSwitch, multimem.st writes the sums back to every replica.  All ranks call it
 * (between two barriers of their own) and the whole range is all-reduced.  offset / count: multiples of 4.
 * max_blocks > 0 confines the kernel to that many CTAs (the reduction is a chain of switch round trips,
 * not bandwidth bound, so a few CTAs are enough and the rest of the GPU stays free for a concurrent kernel). */
GSR_API int gsr_nvls_allreduce_slice(float* multicast_ptr, size_t offset_floats, size_t count_floats,
                                     int rank, int world, int max_blocks, void* stream);

/* All-gather by P2P loads: copies count_floats floats from each of the nviews source pointers (HOST array of
 * DEVICE pointers, typically the peers' replicas of a symmetric buffer) to dst + v * dst_stride_floats.  A pure
 * copy bound by the NVLink ports; max_blocks > 0 confines it to that many CTAs so that it can run underneath
 * another kernel.  count / stride: multiples of 4 floats; pointers 16-byte aligned; at most 16 views. */
GSR_API int gsr_p2p_gather(const float* const* src_ptrs, int nviews, size_t count_floats, float* dst,
                           size_t dst_stride_floats, int max_blocks, void* stream);

/* The same slice all-reduce over plain P2P loads and stores (no multicast needed): replica_ptrs is a HOST array
 * of `world` DEVICE pointers to the replicas of the symmetric buffer (entry `rank` is this rank's own).  The
 * rank reads its 1/world slice from every replica, adds the replicas in rank order (all ranks obtain
 * bit-identical sums) and stores the result into every replica.  All ranks call it between two barriers of
 * their own.  offset / count: multiples of 4 floats; pointers 16-byte aligned; at most 16 ranks. */
GSR_API int gsr_p2p_allreduce_slice(float* const* replica_ptrs, size_t offset_floats, size_t count_floats,
                                    int rank, int world, int max_blocks, void* stream);

/* ---- pose tracker (not part of the reference surface) ------------
 * K iterations of CG-SLAM's tracking loop — render(-light, map_off) -> masked L1 colour + depth loss
 * -> backward -> dL/dviewmatrix -> chain rule to (quaternion, translation) -> Adam step — run on the
 * device with no host interaction: one iteration is 8 kernels + 1 memset captured in a CUDA graph
 * (static-capacity binning, loss and cotangents fused into the forward blend, pose update on chip).
 * It replaces, for this loop, the reference's per-iteration sequence
 *   _C.rasterize_gaussians (L/rasterize_points.cu:36-128, blocking count read-back
 *   L/cuda_rasterizer/rasterizer_impl.cu:287) -> torch loss -> _C.rasterize_gaussians_backward
 *   (L/rasterize_points.cu:130-235) -> torch.sum(dL_dview) (L/diff_gaussian_rasterization/__init__.py:160)
 *   -> autograd through the pose parametrisation -> torch.optim.Adam.
 * Pose: W2C = [R(q/|q|) t; 0 1], q = (w,x,y,z); viewmatrix = W2C^T in the reference's layout.
 * Loss: L = sum_pix m (w_color |C - C_gt|_1 + w_depth |D - D_gt|),
 *       m = (use_depth_mask == 0 || D_gt > 0) && (alpha > alpha_thresh), mask treated as constant. */
typedef struct gsr_tracker gsr_tracker;

typedef struct gsr_track_params {
  float w_color, w_depth;   /* loss weights */
  float alpha_thresh;       /* silhouette mask threshold on the rendered alpha; < 0: off */
  int use_depth_mask;       /* 1: only pixels with gt_depth > 0 */
  float lr_rot, lr_trans;   /* Adam learning rates of the quaternion / the translation */
  float beta1, beta2, eps;  /* Adam constants (torch.optim.Adam semantics) */
} gsr_track_params;

typedef struct gsr_track_result {
  float q[4], t[3];            /* pose after the last iteration */
  float last_dL_dview[16];     /* dL/dviewmatrix of the last iteration (reference layout, before the step) */
  float last_grad[7];          /* dL/dq[4], dL/dt[3] of the last iteration */
  float last_twist_grad[6];    /* the same gradient projected on the SE(3) tangent space: dL/d(omega, v)
                                  for the left perturbation W2C' = exp(xi^) W2C */
  int iterations;              /* iterations run */
  int num_rendered;            /* (Gaussian, tile) duplicates of the probing forward at the start pose */
  int retries;                 /* reruns after a binning-buffer overflow */
  int kernels_per_iteration;   /* graph nodes per iteration */
} gsr_track_result;

/* perspec_matrix_host: 16 floats on the HOST, the reference's `perspec_matrix` argument (P^T row-major).
 * max_iterations bounds `iterations` of gsr_tracker_run (size of the device-side loss history).
 * Returns NULL on failure (see gsr_last_error). */
GSR_API gsr_tracker* gsr_tracker_create(int P, int D, int M, int width, int height, float tan_fovx,
                                        float tan_fovy, const float* perspec_matrix_host,
                                        int max_iterations);
GSR_API void gsr_tracker_destroy(gsr_tracker* t);
/* Device pointers, borrowed (must stay valid and unchanged in address while the tracker uses them);
 * same meaning and optional-NULL rules as gsr_light_forward. */
GSR_API int gsr_tracker_set_scene(gsr_tracker* t, const float* means3D, const float* shs,
                                  const float* colors_precomp, const float* opacities,
                                  const float* scales, float scale_modifier, const float* rotations,
                                  const float* cov3D_precomp, const float* background);
/* gt_color[3,H,W], gt_depth[H,W]: device pointers, borrowed. */
GSR_API int gsr_tracker_set_frame(gsr_tracker* t, const float* gt_color, const float* gt_depth);
/* Host pointers; resets the Adam state. */
GSR_API int gsr_tracker_set_pose(gsr_tracker* t, const float* quat_wxyz, const float* trans);
/* Runs `iterations` tracking iterations from the current pose (Adam state carries over between
 * calls) and blocks until they are done.  loss_history (host, [iterations]) and result may be NULL.
 * Stream contract: the tracker works on a private non-blocking stream.  Before it reads the borrowed
 * scene / frame tensors it waits (on the device) for everything enqueued so far on `caller_stream`
 * (a cudaStream_t as void*; NULL = the legacy default stream) — the stream on which the caller
 * produced or updated those tensors (activations, optimiser steps, the in-place copy of a new frame).
 * Because the call blocks until the iterations are done, work enqueued afterwards on any stream
 * sees the tracker's reads completed.  Tensors written by OTHER streams must be ordered into
 * caller_stream by the caller. */
GSR_API int gsr_tracker_run(gsr_tracker* t, const gsr_track_params* params, int iterations,
                            float* loss_history, gsr_track_result* result, void* caller_stream);

/* ---- RGB-D L1 loss + cotangents (helper next to the hot path; not part of the reference surface) ----
 * One pass over the rendered images of a frame:
 *   L = w_color sum |C - C_gt| + w_depth sum_m |D - D_gt|
 *       + (light) w_aux0 sum_m |D_median - D_gt| + w_aux1 sum depth_var
 *       + (full)  w_aux0 sum (1 - O)                     (O = the "uncertainty" / accumulated-opacity output)
 * m = 1, or (depth_mask != 0) the pixels with D_gt > 0.  Writes L to loss[0] (device) and the cotangent
 * images dL/dC [3,H,W], dL/dD, dL/daux0, dL/daux1 (light only) [H,W] that gsr_*_backward consume as dL_dpix,
 * dL_dpix_depth, dL_dpix_median_depth / dL_dpix_uncertainty, dL_dpix_depth_var.  The ground truth comes as
 * fp32 ([3,H,W] colour, [H,W] depth) or in dataset formats: uint8 colour scaled by color_scale (1/255) and
 * int16 depth scaled by depth_scale (1e-3 for millimetres).  The sum is formed in a fixed order
 * (deterministic).  scratch: gsr_rgbd_l1_scratch_floats(width, height) floats of device memory. */
typedef struct gsr_rgbd_l1 {
  float w_color, w_depth, w_aux0, w_aux1;
  float color_scale, depth_scale;
  int depth_mask;
} gsr_rgbd_l1;
GSR_API size_t gsr_rgbd_l1_scratch_floats(int width, int height);
GSR_API int gsr_rgbd_l1_loss(int variant /* 0 light, 1 full */, int width, int height,
                             const float* color, const float* depth, const float* aux0, const float* aux1,
                             const void* gt_color, int gt_color_is_u8, const void* gt_depth, int gt_depth_is_i16,
                             const gsr_rgbd_l1* prm, float* dL_dcolor, float* dL_ddepth, float* dL_daux0,
                             float* dL_daux1, float* loss, float* scratch, void* stream);

/* replaces Rasterizer::markVisible (rasterizer.h:24-29): present[i] = view-space z > 0.2.
 * `present` is one byte per Gaussian (0/1). */
GSR_API int gsr_mark_visible(int P, const float* means3D, const float* viewmatrix,
                     const float* projmatrix, unsigned char* present, void* stream);

/* Debug / test aid: decode the private geometry buffer into caller arrays (any may be NULL):
 * depths[P], means2D[P,2], conic_opacity[P,4], rgb[P,3], cov3D[P,6], tiles_touched[P],
 * clamped[P,3] (bytes). */
GSR_API int gsr_decode_geometry(const char* geom_buffer, int P, float* depths, float* means2D,
                        float* conic_opacity, float* rgb, float* cov3D,
                        uint32_t* tiles_touched, unsigned char* clamped, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GSR_B200_H */
