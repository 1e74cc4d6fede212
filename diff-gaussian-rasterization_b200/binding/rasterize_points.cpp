// rasterize_points.cpp — thin torch shim over the C ABI of libgsr_b200.so.
//
// Mirrors the reference binding (light: rasterize_points.cu:35-257, full: :35-260): shape
// check on means3D, allocation of outputs / gradients / the three opaque state tensors, pointer
// hand-off.  Differences, all invisible to the Python wrapper:
//   * kernels run on the CURRENT torch stream of means3D's device (the reference launches on
//     the legacy default stream);
//   * outputs are torch::empty — the core writes every element (no zero-fill passes);
//   * light backward returns dL_dview as [1,4,4] (already reduced over pixels) instead of
//     [H*W,4,4]; the wrapper's torch.sum(dim=0) still yields the same [4,4];
//   * empty ("absent") optional tensors are passed as NULL, like the reference's null data_ptr.
#include "rasterize_points.h"

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "../../include/gsr_b200.h"

namespace {

char* resize_cb(void* ctx, size_t bytes) {
  auto* t = static_cast<torch::Tensor*>(ctx);
  // growing a buffer that already holds something (the binning buffer after a speculation miss): resize_ would
  // copy the old contents into the new storage; a fresh tensor does not (the old storage is released in stream order)
  if (t->numel() != 0 && static_cast<size_t>(t->numel()) < bytes)
    *t = torch::empty({static_cast<long long>(bytes)}, t->options());
  else
    t->resize_({static_cast<long long>(bytes)});
  return reinterpret_cast<char*>(t->data_ptr());
}

// contiguous fp32 CUDA view of an input, or an undefined tensor for "absent" (empty) inputs
torch::Tensor prep(const torch::Tensor& t, const torch::Device& dev, const char* name) {
  if (!t.defined() || t.numel() == 0) return torch::Tensor();
  TORCH_CHECK(t.scalar_type() == torch::kFloat32, name, " must be float32");
  TORCH_CHECK(t.device() == dev, name, " must live on ", dev, " (got ", t.device(), ")");
  torch::Tensor c = t.contiguous();
  // the core reads rotations / SH rows with 128-bit loads: a contiguous view at a storage offset
  // that is not a multiple of 16 bytes is re-packed (fresh allocations are always aligned)
  if ((reinterpret_cast<uintptr_t>(c.data_ptr()) & 15) != 0) c = c.clone();
  return c;
}

const float* fptr(const torch::Tensor& t) {
  return t.defined() ? t.data_ptr<float>() : nullptr;
}
float* fptr_mut(const torch::Tensor& t) {
  return t.defined() ? t.data_ptr<float>() : nullptr;
}

void check_rc(int rc, const char* who) {
  TORCH_CHECK(rc == GSR_OK, who, " failed (", rc, "): ", gsr_last_error());
}

torch::Tensor scratch_for(int P, const torch::TensorOptions& fopts) {
  return torch::empty({static_cast<long long>(gsr_backward_scratch_floats(P))}, fopts);
}

torch::Tensor g_grad_arena;  // see setGradArena
bool g_arena_armed = false;  // the arena takes the NEXT backward only (one-shot), see armGradArena
torch::Tensor g_densify_accum, g_densify_denom, g_max_radii;  // see setDensifyStats
bool g_arena_factorized = false;
// factorized arena: masked colour gradient written right after the blend backward + event (gsr_backward_extras)
bool g_arena_early = false;
cudaEvent_t g_masked_ready = nullptr;   // recorded by the backward that took the arena
bool g_masked_recorded = false;

struct SceneGrads {
  torch::Tensor means3D, sh, opacity, scales, rotations;
  torch::Tensor masked_color;  // factorized arena only
  bool factorized = false;
};

// densification accumulators registered with setDensifyStats, when they fit this scene
void fill_densify(gsr_backward_extras& ex, int P, const torch::Tensor& like) {
  if (g_densify_accum.defined() && g_densify_accum.numel() == P && g_densify_accum.device() == like.device()) {
    ex.densify_grad_accum = g_densify_accum.data_ptr<float>();
    ex.densify_denom = g_densify_denom.data_ptr<float>();
  }
  if (g_max_radii.defined() && g_max_radii.numel() == P && g_max_radii.device() == like.device())
    ex.max_radii2D = g_max_radii.data_ptr<float>();
}

// the five scene-parameter gradients: views of the registered arena when it fits, fresh tensors otherwise
SceneGrads alloc_scene_grads(int P, int M, const torch::Tensor& like, const torch::TensorOptions& fopts) {
  SceneGrads g;
  // One-shot: the arena is OVERWRITTEN, not accumulated into, so it may only take one backward per
  // exchange.  A further backward before the arena is re-armed (several views per rank, a tracking
  // pass on the same scene) gets fresh tensors and autograd adds them to .grad as usual.
  const bool armed = g_arena_armed;
  const bool arena_ok = armed && g_grad_arena.defined() && g_grad_arena.device() == like.device() &&
                        g_grad_arena.scalar_type() == torch::kFloat32 && g_grad_arena.is_contiguous() &&
                        P > 0 && P % 4 == 0;
  if (arena_ok && g_arena_factorized && M > 0 && g_grad_arena.numel() == (int64_t)P * 14 + 4) {
    int64_t off = 0;
    auto take = [&](int64_t n, std::vector<int64_t> shape) {
      torch::Tensor t = g_grad_arena.narrow(0, off, n).view(shape);
      off += n;
      return t;
    };
    g.factorized = true;
    g.masked_color = take((int64_t)P * 3, {P, 3});
    off += 4;  // campos + pad, filled by the caller
    g.means3D = take((int64_t)P * 3, {P, 3});
    g.opacity = take((int64_t)P, {P, 1});
    g.scales = take((int64_t)P * 3, {P, 3});
    g.rotations = take((int64_t)P * 4, {P, 4});
    g_arena_armed = false;
    return g;  // g.sh stays undefined: dL_dsh is not written in this mode
  }
  const int64_t need = (int64_t)P * (3 + 3 * (int64_t)M + 1 + 3 + 4);
  if (armed && !g_arena_factorized && g_grad_arena.defined() && g_grad_arena.numel() == need && g_grad_arena.device() == like.device() &&
      g_grad_arena.scalar_type() == torch::kFloat32 && g_grad_arena.is_contiguous() && P > 0 &&
      P % 4 == 0 /* keeps every slice 16-byte aligned for the kernels' 128-bit stores */) {
    int64_t off = 0;
    auto take = [&](int64_t n, std::vector<int64_t> shape) {
      torch::Tensor t = g_grad_arena.narrow(0, off, n).view(shape);
      off += n;
      return t;
    };
    g.means3D = take((int64_t)P * 3, {P, 3});
    g.sh = take((int64_t)P * M * 3, {P, M, 3});
    g.opacity = take((int64_t)P, {P, 1});
    g.scales = take((int64_t)P * 3, {P, 3});
    g.rotations = take((int64_t)P * 4, {P, 4});
    g_arena_armed = false;
  } else {
    g.means3D = torch::empty({P, 3}, fopts);
    g.sh = torch::empty({P, M, 3}, fopts);
    g.opacity = torch::empty({P, 1}, fopts);
    g.scales = torch::empty({P, 3}, fopts);
    g.rotations = torch::empty({P, 4}, fopts);
  }
  return g;
}

}  // namespace

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
                       const bool debug) {
  if (means3D.ndimension() != 2 || means3D.size(1) != 3) {
    AT_ERROR("means3D must have dimensions (num_points, 3)");
  }
  TORCH_CHECK(means3D.is_cuda(), "means3D must be a CUDA tensor");
  const c10::cuda::CUDAGuard guard(means3D.device());
  const auto dev = means3D.device();
  const int P = means3D.size(0);
  const int H = image_height, W = image_width;
  auto fopts = means3D.options().dtype(torch::kFloat32);
  auto iopts = means3D.options().dtype(torch::kInt32);
  auto bopts = means3D.options().dtype(torch::kByte);

  torch::Tensor out_color = torch::empty({3, H, W}, fopts);
  torch::Tensor out_depth = torch::empty({1, H, W}, fopts);
  torch::Tensor out_median_depth = torch::empty({1, H, W}, fopts);
  torch::Tensor out_depth_var = torch::empty({1, H, W}, fopts);
  torch::Tensor out_alpha = torch::empty({1, H, W}, fopts);
  torch::Tensor radii = torch::empty({P}, iopts);
  torch::Tensor gau_uncertainty = torch::empty({P, 1}, fopts);
  torch::Tensor gau_related_pixels = torch::empty({P, 1}, iopts);
  torch::Tensor geomBuffer = torch::empty({0}, bopts);
  torch::Tensor binningBuffer = torch::empty({0}, bopts);
  torch::Tensor imgBuffer = torch::empty({0}, bopts);

  int M = 0;
  if (sh.size(0) != 0) M = sh.size(1);

  const auto bg = prep(background, dev, "bg"), mu = prep(means3D, dev, "means3D"),
             col = prep(colors, dev, "colors_precomp"), op = prep(opacity, dev, "opacities"),
             sc = prep(scales, dev, "scales"), rot = prep(rotations, dev, "rotations"),
             cov = prep(cov3D_precomp, dev, "cov3D_precomp"), vm = prep(viewmatrix, dev, "viewmatrix"),
             gtd = prep(gt_depth, dev, "gt_depth"), pm = prep(projmatrix, dev, "projmatrix"),
             shc = prep(sh, dev, "shs"), cp = prep(campos, dev, "campos");
  if (P != 0) TORCH_CHECK(gtd.defined() && gtd.numel() == (int64_t)H * W, "gt_depth must hold H*W values");

  int rendered = 0;
  const int rc = gsr_light_forward(
      resize_cb, &geomBuffer, resize_cb, &binningBuffer, resize_cb, &imgBuffer, P, degree, M,
      fptr(bg), W, H, fptr(mu), fptr(shc), fptr(col), fptr(op), fptr(sc), scale_modifier, fptr(rot),
      fptr(cov), fptr(vm), fptr(pm), fptr(cp), tan_fovx, tan_fovy, prefiltered ? 1 : 0,
      out_color.data_ptr<float>(), out_depth.data_ptr<float>(), out_median_depth.data_ptr<float>(),
      out_alpha.data_ptr<float>(), fptr(gtd), out_depth_var.data_ptr<float>(),
      P ? gau_uncertainty.data_ptr<float>() : nullptr, P ? gau_related_pixels.data_ptr<int>() : nullptr,
      P ? radii.data_ptr<int>() : nullptr, debug ? 1 : 0,
      at::cuda::getCurrentCUDAStream().stream(), &rendered);
  check_rc(rc, "gsr_light_forward");
  return std::make_tuple(rendered, out_color, out_depth, out_median_depth, out_depth_var, out_alpha,
                         radii, geomBuffer, binningBuffer, imgBuffer, gau_uncertainty,
                         gau_related_pixels);
}

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
    const bool track_off, const bool map_off) {
  return RasterizeGaussiansBackwardSelect(background, means3D, radii, colors, scales, rotations, scale_modifier,
                                          cov3D_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color,
                                          dL_dout_depth, dL_dout_median_depth, dL_dout_depth_var, gt_depth, sh, degree,
                                          campos, geomBuffer, R, binningBuffer, imageBuffer, alphas, debug,
                                          perspec_matrix, track_off, map_off, true, true);
}

// The same call with the two gradients autograd throws away when SH colours / scale + rotation are used
// (dL/dcolors_precomp, dL/dcov3D_precomp: 36 bytes per Gaussian) made optional: not wanted -> not written, an
// empty tensor is returned in their place.  The Python package passes ctx.needs_input_grad.
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
    const bool track_off, const bool map_off, const bool want_colors_grad, const bool want_cov3D_grad) {
  const c10::cuda::CUDAGuard guard(means3D.device());
  const auto dev = means3D.device();
  const int P = means3D.size(0);
  const int H = dL_dout_color.size(1);
  const int W = dL_dout_color.size(2);
  int M = 0;
  if (sh.size(0) != 0) M = sh.size(1);
  auto fopts = means3D.options().dtype(torch::kFloat32);

  SceneGrads sg = alloc_scene_grads(P, M, means3D, fopts);
  torch::Tensor dL_dmeans3D = sg.means3D;
  torch::Tensor dL_dmeans2D = torch::empty({P, 3}, fopts);
  torch::Tensor dL_dcolors = torch::empty({want_colors_grad ? P : 0, 3}, fopts);
  // dL/dconic and dL/ddepth are intermediates of the reference (written by its blend backward, read by its
  // per-Gaussian backward, never returned): here they live in registers of the fused per-Gaussian kernel
  torch::Tensor dL_dopacity = sg.opacity;
  torch::Tensor dL_dcov3D = torch::empty({want_cov3D_grad ? P : 0, 6}, fopts);
  torch::Tensor dL_dsh = sg.sh;  // undefined in the factorized arena mode
  gsr_backward_extras extras{nullptr, 0, nullptr, nullptr, nullptr, nullptr};
  fill_densify(extras, P, means3D);
  if (sg.factorized) {
    extras.dL_dcolor_masked = sg.masked_color.data_ptr<float>();
    extras.skip_sh_grad = 1;
    g_grad_arena.narrow(0, (int64_t)P * 3, 3).copy_(campos.reshape({-1}).narrow(0, 0, 3), /*non_blocking=*/true);
    if (g_arena_early && sh.numel() != 0 && !map_off) {
      if (g_masked_ready == nullptr)
        TORCH_CHECK(cudaEventCreateWithFlags(&g_masked_ready, cudaEventDisableTiming) == cudaSuccess, "cudaEventCreate failed");
      extras.masked_color_ready_event = g_masked_ready;
      g_masked_recorded = true;
    }
  }
  torch::Tensor dL_dscales = sg.scales;
  torch::Tensor dL_drotations = sg.rotations;
  torch::Tensor dL_dview = torch::empty({1, 4, 4}, fopts);
  torch::Tensor scratch = scratch_for(P, fopts);

  const auto bg = prep(background, dev, "bg"), mu = prep(means3D, dev, "means3D"),
             col = prep(colors, dev, "colors_precomp"), sc = prep(scales, dev, "scales"),
             rot = prep(rotations, dev, "rotations"), cov = prep(cov3D_precomp, dev, "cov3D_precomp"),
             vm = prep(viewmatrix, dev, "viewmatrix"), pm = prep(projmatrix, dev, "projmatrix"),
             gc = prep(dL_dout_color, dev, "dL_dout_color"), gd = prep(dL_dout_depth, dev, "dL_dout_depth"),
             gm = prep(dL_dout_median_depth, dev, "dL_dout_median_depth"),
             gv = prep(dL_dout_depth_var, dev, "dL_dout_depth_var"), gtd = prep(gt_depth, dev, "gt_depth"),
             shc = prep(sh, dev, "shs"), cp = prep(campos, dev, "campos"), al = prep(alphas, dev, "alphas"),
             per = prep(perspec_matrix, dev, "perspec_matrix");
  const torch::Tensor rad = radii.contiguous();
  const torch::Tensor gb = geomBuffer.contiguous(), bb = binningBuffer.contiguous(),
                      ib = imageBuffer.contiguous();

  const int rc = gsr_light_backward(
      P, degree, M, R, fptr(bg), W, H, fptr(mu), fptr(shc), fptr(col), fptr(al), fptr(sc),
      scale_modifier, fptr(rot), fptr(cov), fptr(vm), fptr(pm), fptr(cp), tan_fovx, tan_fovy,
      P ? rad.data_ptr<int>() : nullptr, reinterpret_cast<char*>(gb.data_ptr()),
      reinterpret_cast<char*>(bb.data_ptr()), reinterpret_cast<char*>(ib.data_ptr()), fptr(gc),
      fptr(gd), fptr(gm), fptr(gv), dL_dmeans2D.data_ptr<float>(), nullptr,
      dL_dopacity.data_ptr<float>(), (want_colors_grad ? dL_dcolors.data_ptr<float>() : nullptr), nullptr,
      dL_dmeans3D.data_ptr<float>(), (want_cov3D_grad ? dL_dcov3D.data_ptr<float>() : nullptr), fptr_mut(dL_dsh),
      dL_dscales.data_ptr<float>(), dL_drotations.data_ptr<float>(), debug ? 1 : 0, fptr(per),
      dL_dview.data_ptr<float>(), fptr(gtd), track_off ? 1 : 0, map_off ? 1 : 0,
      scratch.data_ptr<float>(), at::cuda::getCurrentCUDAStream().stream(), &extras);
  check_rc(rc, "gsr_light_backward");
  return std::make_tuple(dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh,
                         dL_dscales, dL_drotations, dL_dview);
}

#else  // GSR_VARIANT_FULL

std::tuple<int, int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor,
           torch::Tensor, torch::Tensor>
RasterizeGaussiansCUDA(const torch::Tensor& background, const torch::Tensor& means3D,
                       const torch::Tensor& colors, const torch::Tensor& opacity,
                       const torch::Tensor& scales, const torch::Tensor& rotations,
                       const float scale_modifier, const torch::Tensor& cov3D_precomp,
                       const torch::Tensor& viewmatrix, const torch::Tensor& gt_depth,
                       const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy,
                       const int image_height, const int image_width, const torch::Tensor& sh,
                       const int degree, const torch::Tensor& campos, const bool prefiltered) {
  if (means3D.ndimension() != 2 || means3D.size(1) != 3) {
    AT_ERROR("means3D must have dimensions (num_points, 3)");
  }
  TORCH_CHECK(means3D.is_cuda(), "means3D must be a CUDA tensor");
  const c10::cuda::CUDAGuard guard(means3D.device());
  const auto dev = means3D.device();
  const int P = means3D.size(0);
  const int H = image_height, W = image_width;
  auto fopts = means3D.options().dtype(torch::kFloat32);
  auto iopts = means3D.options().dtype(torch::kInt32);
  auto bopts = means3D.options().dtype(torch::kByte);

  torch::Tensor out_color = torch::empty({3, H, W}, fopts);
  torch::Tensor out_depth = torch::empty({1, H, W}, fopts);
  torch::Tensor out_uncertainty = torch::empty({1, H, W}, fopts);
  torch::Tensor radii = torch::empty({P}, iopts);
  torch::Tensor geomBuffer = torch::empty({0}, bopts);
  torch::Tensor binningBuffer = torch::empty({0}, bopts);
  torch::Tensor imgBuffer = torch::empty({0}, bopts);

  int M = 0;
  if (sh.size(0) != 0) M = sh.size(1);

  const auto bg = prep(background, dev, "bg"), mu = prep(means3D, dev, "means3D"),
             col = prep(colors, dev, "colors_precomp"), op = prep(opacity, dev, "opacities"),
             sc = prep(scales, dev, "scales"), rot = prep(rotations, dev, "rotations"),
             cov = prep(cov3D_precomp, dev, "cov3D_precomp"), vm = prep(viewmatrix, dev, "viewmatrix"),
             gtd = prep(gt_depth, dev, "gt_depth"), pm = prep(projmatrix, dev, "projmatrix"),
             shc = prep(sh, dev, "shs"), cp = prep(campos, dev, "campos");

  int rendered = 0, related = 0;
  const int rc = gsr_full_forward(
      resize_cb, &geomBuffer, resize_cb, &binningBuffer, resize_cb, &imgBuffer, P, degree, M,
      fptr(bg), W, H, fptr(mu), fptr(shc), fptr(col), fptr(op), fptr(sc), scale_modifier, fptr(rot),
      fptr(cov), fptr(vm), fptr(pm), fptr(cp), tan_fovx, tan_fovy, prefiltered ? 1 : 0,
      out_color.data_ptr<float>(), out_depth.data_ptr<float>(), out_uncertainty.data_ptr<float>(),
      P ? radii.data_ptr<int>() : nullptr, fptr(gtd), at::cuda::getCurrentCUDAStream().stream(),
      &rendered, &related);
  check_rc(rc, "gsr_full_forward");
  return std::make_tuple(rendered, related, out_color, out_depth, out_uncertainty, radii, geomBuffer,
                         binningBuffer, imgBuffer);
}

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
    const torch::Tensor& perspec_matrix) {
  return RasterizeGaussiansBackwardSelect(background, means3D, radii, colors, scales, rotations, scale_modifier,
                                          cov3D_precomp, viewmatrix, gt_depth, projmatrix, tan_fovx, tan_fovy,
                                          dL_dout_color, dL_dout_depth, dL_dout_uncertainty, sh, degree, campos,
                                          geomBuffer, R, binningBuffer, imageBuffer, NG, perspec_matrix, true, true);
}

// (see the -light variant: dL/dcolors_precomp and dL/dcov3D_precomp made optional)
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
    const torch::Tensor& perspec_matrix, const bool want_colors_grad, const bool want_cov3D_grad) {
  (void)NG;  // sized the reference's per-pair scratch lists; there are none here
  const c10::cuda::CUDAGuard guard(means3D.device());
  const auto dev = means3D.device();
  const int P = means3D.size(0);
  const int H = dL_dout_color.size(1);
  const int W = dL_dout_color.size(2);
  int M = 0;
  if (sh.size(0) != 0) M = sh.size(1);
  auto fopts = means3D.options().dtype(torch::kFloat32);

  SceneGrads sg = alloc_scene_grads(P, M, means3D, fopts);
  torch::Tensor dL_dmeans3D = sg.means3D;
  torch::Tensor dL_dmeans2D = torch::empty({P, 3}, fopts);
  torch::Tensor dL_dcolors = torch::empty({want_colors_grad ? P : 0, 3}, fopts);
  torch::Tensor dL_dopacity = sg.opacity;
  torch::Tensor dL_dcov3D = torch::empty({want_cov3D_grad ? P : 0, 6}, fopts);
  torch::Tensor dL_dsh = sg.sh;  // undefined in the factorized arena mode
  gsr_backward_extras extras{nullptr, 0, nullptr, nullptr, nullptr, nullptr};
  fill_densify(extras, P, means3D);
  if (sg.factorized) {
    extras.dL_dcolor_masked = sg.masked_color.data_ptr<float>();
    extras.skip_sh_grad = 1;
    g_grad_arena.narrow(0, (int64_t)P * 3, 3).copy_(campos.reshape({-1}).narrow(0, 0, 3), /*non_blocking=*/true);
    if (g_arena_early && sh.numel() != 0) {
      if (g_masked_ready == nullptr)
        TORCH_CHECK(cudaEventCreateWithFlags(&g_masked_ready, cudaEventDisableTiming) == cudaSuccess, "cudaEventCreate failed");
      extras.masked_color_ready_event = g_masked_ready;
      g_masked_recorded = true;
    }
  }
  torch::Tensor dL_dscales = sg.scales;
  torch::Tensor dL_drotations = sg.rotations;
  torch::Tensor dL_dview = torch::empty({4, 4}, fopts);
  torch::Tensor scratch = scratch_for(P, fopts);

  const auto bg = prep(background, dev, "bg"), mu = prep(means3D, dev, "means3D"),
             col = prep(colors, dev, "colors_precomp"), sc = prep(scales, dev, "scales"),
             rot = prep(rotations, dev, "rotations"), cov = prep(cov3D_precomp, dev, "cov3D_precomp"),
             vm = prep(viewmatrix, dev, "viewmatrix"), pm = prep(projmatrix, dev, "projmatrix"),
             gc = prep(dL_dout_color, dev, "dL_dout_color"), gd = prep(dL_dout_depth, dev, "dL_dout_depth"),
             gu = prep(dL_dout_uncertainty, dev, "dL_dout_uncertainty"),
             gtd = prep(gt_depth, dev, "gt_depth"), shc = prep(sh, dev, "shs"),
             cp = prep(campos, dev, "campos"), per = prep(perspec_matrix, dev, "perspec_matrix");
  const torch::Tensor rad = radii.contiguous();
  const torch::Tensor gb = geomBuffer.contiguous(), bb = binningBuffer.contiguous(),
                      ib = imageBuffer.contiguous();

  const int rc = gsr_full_backward(
      P, degree, M, R, fptr(bg), W, H, fptr(mu), fptr(shc), fptr(col), fptr(sc), scale_modifier,
      fptr(rot), fptr(cov), fptr(vm), fptr(pm), fptr(cp), tan_fovx, tan_fovy,
      P ? rad.data_ptr<int>() : nullptr, reinterpret_cast<char*>(gb.data_ptr()),
      reinterpret_cast<char*>(bb.data_ptr()), reinterpret_cast<char*>(ib.data_ptr()), fptr(gc),
      fptr(gd), fptr(gu), dL_dmeans2D.data_ptr<float>(), nullptr,
      dL_dopacity.data_ptr<float>(), (want_colors_grad ? dL_dcolors.data_ptr<float>() : nullptr), nullptr,
      dL_dmeans3D.data_ptr<float>(), (want_cov3D_grad ? dL_dcov3D.data_ptr<float>() : nullptr), fptr_mut(dL_dsh),
      dL_dscales.data_ptr<float>(), dL_drotations.data_ptr<float>(), fptr(per),
      dL_dview.data_ptr<float>(), fptr(gtd), scratch.data_ptr<float>(),
      at::cuda::getCurrentCUDAStream().stream(), &extras);
  check_rc(rc, "gsr_full_backward");
  return std::make_tuple(dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh,
                         dL_dscales, dL_drotations, dL_dview);
}

#endif

// Registers (or, with undefined / empty tensors, clears) the densification accumulators that the
// following backward calls update in place (gsr_backward_extras in include/gsr_b200.h).
void setDensifyStats(const torch::Tensor& grad_accum, const torch::Tensor& denom,
                     const torch::Tensor& max_radii2D) {
  auto ok = [](const torch::Tensor& t) {
    return t.defined() && t.numel() > 0;
  };
  auto check = [](const torch::Tensor& t, const char* name) {
    TORCH_CHECK(t.is_cuda() && t.scalar_type() == torch::kFloat32 && t.is_contiguous(), name,
                " must be a contiguous float32 CUDA tensor");
  };
  g_densify_accum = torch::Tensor(); g_densify_denom = torch::Tensor(); g_max_radii = torch::Tensor();
  if (ok(grad_accum) || ok(denom)) {
    TORCH_CHECK(ok(grad_accum) && ok(denom) && grad_accum.numel() == denom.numel(),
                "grad_accum and denom must be given together and have the same number of elements");
    check(grad_accum, "grad_accum"); check(denom, "denom");
    g_densify_accum = grad_accum; g_densify_denom = denom;
  }
  if (ok(max_radii2D)) {
    check(max_radii2D, "max_radii2D");
    g_max_radii = max_radii2D;
  }
}

void setGradArena(const torch::Tensor& arena, bool factorized_sh, bool early_masked_color) {
  g_arena_factorized = factorized_sh;
  g_arena_early = factorized_sh && early_masked_color;
  g_masked_recorded = false;
  if (!arena.defined() || arena.numel() == 0) {
    g_grad_arena = torch::Tensor();
    g_arena_factorized = false;
    g_arena_armed = false;
    return;
  }
  TORCH_CHECK(arena.is_cuda() && arena.scalar_type() == torch::kFloat32 && arena.is_contiguous() &&
                  arena.dim() == 1,
              "grad arena must be a contiguous 1-D float32 CUDA tensor");
  TORCH_CHECK((reinterpret_cast<uintptr_t>(arena.data_ptr()) & 15) == 0,
              "grad arena must be 16-byte aligned (the kernels write it with 128-bit stores)");
  g_grad_arena = arena;
  g_arena_armed = true;
}

// Makes `stream` (a cudaStream_t as an integer) wait until the masked colour gradient of the backward that took
// the arena is complete (set_grad_arena(..., early_masked_color=True)); returns false — and does nothing — when no
// such backward has run since the arena was armed, in which case the caller orders after the whole backward.
bool waitMaskedColor(int64_t stream) {
  if (!g_masked_recorded || g_masked_ready == nullptr) return false;
  TORCH_CHECK(cudaStreamWaitEvent(reinterpret_cast<cudaStream_t>(stream), g_masked_ready, 0) == cudaSuccess,
              "cudaStreamWaitEvent failed");
  g_masked_recorded = false;
  return true;
}

bool armGradArena() {
  g_arena_armed = g_grad_arena.defined();
  return g_arena_armed;
}

torch::Tensor shGradFromViews(const torch::Tensor& means3D, const torch::Tensor& gathered, const int degree,
                              const int M) {
  TORCH_CHECK(means3D.is_cuda() && gathered.is_cuda(), "sh_grad_from_views needs CUDA tensors");
  const c10::cuda::CUDAGuard guard(means3D.device());
  const auto dev = means3D.device();
  const int P = means3D.size(0);
  const auto mu = prep(means3D, dev, "means3D"), ga = prep(gathered, dev, "gathered");
  TORCH_CHECK(ga.dim() == 2 && ga.size(1) == (int64_t)P * 3 + 4, "gathered must be [nviews, 3P + 4]");
  torch::Tensor out = torch::empty({P, M, 3}, means3D.options().dtype(torch::kFloat32));
  const int rc = gsr_sh_grad_from_views(P, degree, M, fptr(mu), (int)ga.size(0), fptr(ga), (size_t)ga.size(1),
                                        fptr(ga) + (size_t)P * 3, (size_t)ga.size(1), out.data_ptr<float>(),
                                        at::cuda::getCurrentCUDAStream().stream());
  check_rc(rc, "gsr_sh_grad_from_views");
  return out;
}

torch::Tensor shGradFromViewPtrs(const torch::Tensor& means3D, const std::vector<int64_t>& dR_ptrs,
                                 const std::vector<int64_t>& campos_ptrs, const int degree, const int M) {
  TORCH_CHECK(means3D.is_cuda(), "sh_grad_from_view_ptrs needs a CUDA tensor");
  TORCH_CHECK(dR_ptrs.size() == campos_ptrs.size(), "one camera position per view");
  const c10::cuda::CUDAGuard guard(means3D.device());
  const auto dev = means3D.device();
  const int P = means3D.size(0);
  const auto mu = prep(means3D, dev, "means3D");
  std::vector<const float*> dr, cp;
  for (size_t v = 0; v < dR_ptrs.size(); ++v) {
    dr.push_back(reinterpret_cast<const float*>(dR_ptrs[v]));
    cp.push_back(reinterpret_cast<const float*>(campos_ptrs[v]));
  }
  torch::Tensor out = torch::empty({P, M, 3}, means3D.options().dtype(torch::kFloat32));
  const int rc = gsr_sh_grad_from_view_ptrs(P, degree, M, fptr(mu), (int)dr.size(), dr.data(), cp.data(),
                                            out.data_ptr<float>(), at::cuda::getCurrentCUDAStream().stream());
  check_rc(rc, "gsr_sh_grad_from_view_ptrs");
  return out;
}

void nvlsAllreduceSlice(int64_t multicast_ptr, int64_t offset_floats, int64_t count_floats, int64_t rank,
                        int64_t world, int64_t max_blocks) {
  const int rc = gsr_nvls_allreduce_slice(reinterpret_cast<float*>(multicast_ptr), (size_t)offset_floats,
                                          (size_t)count_floats, (int)rank, (int)world, (int)max_blocks,
                                          at::cuda::getCurrentCUDAStream().stream());
  check_rc(rc, "gsr_nvls_allreduce_slice");
}

void p2pGather(const std::vector<int64_t>& src_ptrs, int64_t count_floats, torch::Tensor& dst, int64_t max_blocks) {
  TORCH_CHECK(dst.is_cuda() && dst.scalar_type() == torch::kFloat32 && dst.is_contiguous() && dst.dim() == 2 &&
                  dst.size(0) == (int64_t)src_ptrs.size() && dst.size(1) >= count_floats,
              "p2p_gather: dst must be a contiguous float32 CUDA tensor [nviews, >= count]");
  const c10::cuda::CUDAGuard guard(dst.device());
  std::vector<const float*> sp;
  for (int64_t p : src_ptrs) sp.push_back(reinterpret_cast<const float*>(p));
  const int rc = gsr_p2p_gather(sp.data(), (int)sp.size(), (size_t)count_floats, dst.data_ptr<float>(),
                                (size_t)dst.size(1), (int)max_blocks, at::cuda::getCurrentCUDAStream().stream());
  check_rc(rc, "gsr_p2p_gather");
}

void p2pAllreduceSlice(const std::vector<int64_t>& replica_ptrs, int64_t offset_floats, int64_t count_floats,
                       int64_t rank, int64_t max_blocks) {
  std::vector<float*> rp;
  for (int64_t p : replica_ptrs) rp.push_back(reinterpret_cast<float*>(p));
  const int rc = gsr_p2p_allreduce_slice(rp.data(), (size_t)offset_floats, (size_t)count_floats, (int)rank,
                                         (int)rp.size(), (int)max_blocks,
                                         at::cuda::getCurrentCUDAStream().stream());
  check_rc(rc, "gsr_p2p_allreduce_slice");
}

torch::Tensor markVisible(torch::Tensor& means3D, torch::Tensor& viewmatrix,
                          torch::Tensor& projmatrix) {
  TORCH_CHECK(means3D.is_cuda(), "means3D must be a CUDA tensor");
  const c10::cuda::CUDAGuard guard(means3D.device());
  const auto dev = means3D.device();
  const int P = means3D.size(0);
  torch::Tensor present = torch::full({P}, false, means3D.options().dtype(at::kBool));
  if (P != 0) {
    const auto mu = prep(means3D, dev, "means3D"), vm = prep(viewmatrix, dev, "viewmatrix"),
               pm = prep(projmatrix, dev, "projmatrix");
    const int rc = gsr_mark_visible(P, fptr(mu), fptr(vm), fptr(pm),
                                    reinterpret_cast<unsigned char*>(present.data_ptr<bool>()),
                                    at::cuda::getCurrentCUDAStream().stream());
    check_rc(rc, "gsr_mark_visible");
  }
  return present;
}

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
rgbdL1Loss(const torch::Tensor& color, const torch::Tensor& depth, const torch::Tensor& aux0,
           const torch::Tensor& aux1, const torch::Tensor& gt_color, const torch::Tensor& gt_depth,
           double w_color, double w_depth, double w_aux0, double w_aux1, double depth_scale,
           bool depth_mask) {
  TORCH_CHECK(color.is_cuda() && color.dim() == 3 && color.size(0) == 3, "color must be a [3,H,W] CUDA tensor");
  const c10::cuda::CUDAGuard guard(color.device());
  const auto dev = color.device();
  const int H = color.size(1), W = color.size(2);
  const int64_t HW = (int64_t)H * W;
#if defined(GSR_VARIANT_LIGHT)
  const int variant = 0;
#else
  const int variant = 1;
#endif
  const auto c = prep(color, dev, "color"), d = prep(depth, dev, "depth"), a0 = prep(aux0, dev, "aux0");
  const auto a1 = variant == 0 ? prep(aux1, dev, "aux1") : torch::Tensor();
  TORCH_CHECK(d.defined() && d.numel() == HW && a0.defined() && a0.numel() == HW &&
                  (variant == 1 || (a1.defined() && a1.numel() == HW)),
              "depth / aux images must hold H*W values");
  TORCH_CHECK(gt_color.is_cuda() && gt_color.device() == dev && gt_color.numel() == 3 * HW &&
                  (gt_color.scalar_type() == torch::kUInt8 || gt_color.scalar_type() == torch::kFloat32),
              "gt_color must be a uint8 or float32 [3,H,W] tensor on the same device");
  TORCH_CHECK(gt_depth.is_cuda() && gt_depth.device() == dev && gt_depth.numel() == HW &&
                  (gt_depth.scalar_type() == torch::kInt16 || gt_depth.scalar_type() == torch::kFloat32),
              "gt_depth must be an int16 or float32 tensor with H*W values on the same device");
  const torch::Tensor gc = gt_color.contiguous(), gd = gt_depth.contiguous();
  auto fopts = color.options().dtype(torch::kFloat32);
  torch::Tensor loss = torch::empty({1}, fopts);
  torch::Tensor g_color = torch::empty_like(c), g_depth = torch::empty_like(d), g_a0 = torch::empty_like(a0);
  torch::Tensor g_a1 = variant == 0 ? torch::empty_like(a1) : torch::Tensor();
  torch::Tensor scratch = torch::empty({static_cast<long long>(gsr_rgbd_l1_scratch_floats(W, H))}, fopts);
  gsr_rgbd_l1 prm{(float)w_color, (float)w_depth, (float)w_aux0, (float)w_aux1, 1.0f / 255.0f, (float)depth_scale,
                  depth_mask ? 1 : 0};
  const int rc = gsr_rgbd_l1_loss(variant, W, H, fptr(c), fptr(d), fptr(a0), fptr(a1), gc.data_ptr(),
                                  gc.scalar_type() == torch::kUInt8 ? 1 : 0, gd.data_ptr(),
                                  gd.scalar_type() == torch::kInt16 ? 1 : 0, &prm, g_color.data_ptr<float>(),
                                  g_depth.data_ptr<float>(), g_a0.data_ptr<float>(), fptr_mut(g_a1),
                                  loss.data_ptr<float>(), scratch.data_ptr<float>(),
                                  at::cuda::getCurrentCUDAStream().stream());
  check_rc(rc, "gsr_rgbd_l1_loss");
  return std::make_tuple(loss, g_color, g_depth, g_a0, g_a1);
}
