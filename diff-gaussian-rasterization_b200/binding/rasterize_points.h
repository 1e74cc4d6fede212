// rasterize_points.h — torch-facing entry points of the `_C` extension module.
//
// Same three exported operations, argument order and return tuples as the reference binding
// (diff-gaussian-rasterization-light/rasterize_points.h:18-76 with -DGSR_VARIANT_LIGHT,
//  diff-gaussian-rasterization-full/rasterize_points.h:18-72 with -DGSR_VARIANT_FULL); the
// bodies only allocate tensors and hand raw pointers to libgsr_b200.so (include/gsr_b200.h).
#pragma once
#include <torch/extension.h>

#include <tuple>

#if defined(GSR_VARIANT_LIGHT)

std::tuple<int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor,
           torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansCUDA(const torch::Tensor& background, const torch::Tensor& means3D,
                       const torch::Tensor& colors, const torch::Tensor& opacity,
                       const torch::Tensor& scales, const torch::Tensor& rotations,
                       const float scale_modifier, const torch::Tensor& cov3D_precomp,
                       const torch::Tensor& viewmatrix, const torch::Tensor& gt_depth,
                       const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy,
                       const int image_height, const int image_width, const torch::Tensor& sh,
                       const int degree, const torch::Tensor& campos, const bool prefiltered,
                       const bool debug);

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor,
           torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansBackwardCUDA(
    const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& radii,
    const torch::Tensor& colors, const torch::Tensor& scales, const torch::Tensor& rotations,
    const float scale_modifier, const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix,
    const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy,
    const torch::Tensor& dL_dout_color, const torch::Tensor& dL_dout_depth,
    const torch::Tensor& dL_dout_median_depth, const torch::Tensor& dL_dout_depth_var,
    const torch::Tensor& gt_depth, const torch::Tensor& sh, const int degree,
    const torch::Tensor& campos, const torch::Tensor& geomBuffer, const int R,
    const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer,
    const torch::Tensor& alphas, const bool debug, const torch::Tensor& perspec_matrix,
    const bool track_off, const bool map_off);

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor,
           torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansBackwardSelect(
    const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& radii,
    const torch::Tensor& colors, const torch::Tensor& scales, const torch::Tensor& rotations,
    const float scale_modifier, const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix,
    const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy,
    const torch::Tensor& dL_dout_color, const torch::Tensor& dL_dout_depth,
    const torch::Tensor& dL_dout_median_depth, const torch::Tensor& dL_dout_depth_var,
    const torch::Tensor& gt_depth, const torch::Tensor& sh, const int degree,
    const torch::Tensor& campos, const torch::Tensor& geomBuffer, const int R,
    const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer,
    const torch::Tensor& alphas, const bool debug, const torch::Tensor& perspec_matrix,
    const bool track_off, const bool map_off, const bool want_colors_grad, const bool want_cov3D_grad);

#elif defined(GSR_VARIANT_FULL)

std::tuple<int, int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor,
           torch::Tensor, torch::Tensor>
RasterizeGaussiansCUDA(const torch::Tensor& background, const torch::Tensor& means3D,
                       const torch::Tensor& colors, const torch::Tensor& opacity,
                       const torch::Tensor& scales, const torch::Tensor& rotations,
                       const float scale_modifier, const torch::Tensor& cov3D_precomp,
                       const torch::Tensor& viewmatrix, const torch::Tensor& gt_depth,
                       const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy,
                       const int image_height, const int image_width, const torch::Tensor& sh,
                       const int degree, const torch::Tensor& campos, const bool prefiltered);

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor,
           torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansBackwardCUDA(
    const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& radii,
    const torch::Tensor& colors, const torch::Tensor& scales, const torch::Tensor& rotations,
    const float scale_modifier, const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix,
    const torch::Tensor& gt_depth, const torch::Tensor& projmatrix, const float tan_fovx,
    const float tan_fovy, const torch::Tensor& dL_dout_color, const torch::Tensor& dL_dout_depth,
    const torch::Tensor& dL_dout_uncertainty, const torch::Tensor& sh, const int degree,
    const torch::Tensor& campos, const torch::Tensor& geomBuffer, const int R,
    const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer, const int NG,
    const torch::Tensor& perspec_matrix);

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor,
           torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansBackwardSelect(
    const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& radii,
    const torch::Tensor& colors, const torch::Tensor& scales, const torch::Tensor& rotations,
    const float scale_modifier, const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix,
    const torch::Tensor& gt_depth, const torch::Tensor& projmatrix, const float tan_fovx,
    const float tan_fovy, const torch::Tensor& dL_dout_color, const torch::Tensor& dL_dout_depth,
    const torch::Tensor& dL_dout_uncertainty, const torch::Tensor& sh, const int degree,
    const torch::Tensor& campos, const torch::Tensor& geomBuffer, const int R,
    const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer, const int NG,
    const torch::Tensor& perspec_matrix, const bool want_colors_grad, const bool want_cov3D_grad);

#else
#error "define GSR_VARIANT_LIGHT or GSR_VARIANT_FULL"
#endif

// Registers (or clears, with an undefined / empty tensor) a flat fp32 CUDA buffer of
// P*(3 + 3M + 1 + 3 + 4) floats laid out [means3D | sh | opacity | scales | rotations].  While it is
// registered and its size matches, RasterizeGaussiansBackwardCUDA writes those five gradients
// straight into it (the returned tensors are views), so a data-parallel caller can all-reduce one
// buffer without packing.  Not part of the reference surface.
//
// factorized_sh = true selects the layout used by the SH-factorized exchange
//   [ dL_dcolor_masked (3P) | campos (3) + pad (1) | means3D (3P) | opacity (P) | scales (3P) | rotations (4P) ]
// (14P + 4 floats): the backward then writes the masked colour gradient and the camera position
// instead of dL_dsh (whose returned tensor is undefined / None); the summed dL_dsh of all views is
// rebuilt with shGradFromViews after an all-gather of the first 3P + 4 floats.
// The arena is ONE-SHOT: registering arms it, the next backward that fits writes into it and disarms
// it, armGradArena() re-arms it (dp.SceneGradReducer does so after every exchange).  Backwards that
// run while it is disarmed return fresh tensors, which autograd accumulates into .grad as usual — so
// several backwards per exchange (several views per rank, a tracking pass on the same scene) add up
// instead of clobbering each other.
void setGradArena(const torch::Tensor& arena, bool factorized_sh, bool early_masked_color);
bool waitMaskedColor(int64_t stream);
bool armGradArena();
torch::Tensor shGradFromViewPtrs(const torch::Tensor& means3D, const std::vector<int64_t>& dR_ptrs,
                                 const std::vector<int64_t>& campos_ptrs, const int degree, const int M);
void p2pGather(const std::vector<int64_t>& src_ptrs, int64_t count_floats, torch::Tensor& dst, int64_t max_blocks);
void p2pAllreduceSlice(const std::vector<int64_t>& replica_ptrs, int64_t offset_floats, int64_t count_floats,
                       int64_t rank, int64_t max_blocks);
void nvlsAllreduceSlice(int64_t multicast_ptr, int64_t offset_floats, int64_t count_floats, int64_t rank,
                        int64_t world, int64_t max_blocks);
void setDensifyStats(const torch::Tensor& grad_accum, const torch::Tensor& denom,
                     const torch::Tensor& max_radii2D);

// gathered: [nviews, 3P + 4] (rows = the first 3P + 4 floats of each rank's factorized arena).
// Returns sum over views of dL_dsh, [P, M, 3].
torch::Tensor shGradFromViews(const torch::Tensor& means3D, const torch::Tensor& gathered, const int degree,
                              const int M);

torch::Tensor markVisible(torch::Tensor& means3D, torch::Tensor& viewmatrix,
                          torch::Tensor& projmatrix);

// RGB-D L1 loss and its cotangent images in one pass (gsr_rgbd_l1_loss; not part of the reference
// surface).  color [3,H,W], depth / aux0 / aux1 [1,H,W] or [H,W] fp32 = the rasterizer's outputs
// (-light: aux0 = median depth, aux1 = depth_var; -full: aux0 = the opacity map, aux1 ignored);
// gt_color: uint8 or fp32 [3,H,W]; gt_depth: int16 (scaled by depth_scale) or fp32 [H,W] / [1,H,W].
// Returns (loss [1], dL_dcolor, dL_ddepth, dL_daux0, dL_daux1) shaped like the inputs.
std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
rgbdL1Loss(const torch::Tensor& color, const torch::Tensor& depth, const torch::Tensor& aux0,
           const torch::Tensor& aux1, const torch::Tensor& gt_color, const torch::Tensor& gt_depth,
           double w_color, double w_depth, double w_aux0, double w_aux1, double depth_scale,
           bool depth_mask);
