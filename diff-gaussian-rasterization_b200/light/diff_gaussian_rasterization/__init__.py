"""diff_gaussian_rasterization (-light surface) backed by the B200-native core.

Public names, argument orders, return tuples and gradient tuples are those of the reference
package (diff-gaussian-rasterization-light/diff_gaussian_rasterization/__init__.py):
  GaussianRasterizationSettings  (:180-195, 15 fields incl. debug / perspec_matrix / track_off / map_off)
  GaussianRasterizer.forward     (:213-248) -> (color, radii, depth, depth_median, depth_var,
                                               opacity_map, gau_uncertainty, gau_related_pixels)
  GaussianRasterizer.markVisible (:202-211)
  rasterize_gaussians            (:21-46)
The compiled `_C` module next to this file is a thin shim over libgsr_b200.so; importing this
package without it fails (there is no Python / CPU fallback).
"""
from typing import NamedTuple

import torch
import torch.nn as nn

from . import _C  # noqa: F401  (hard requirement: the CUDA extension)

__all__ = ["GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians",
           "set_densify_stats", "rgbd_l1_loss"]


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool
    perspec_matrix: torch.Tensor
    track_off: bool
    map_off: bool


def _snapshot(args):
    return tuple(a.detach().cpu().clone() if isinstance(a, torch.Tensor) else a for a in args)


def _guarded(fn, args, debug, dump_name, what):
    """Run fn(*args); in debug mode keep a CPU copy of the arguments and dump it on failure
    (the reference writes snapshot_fw.dump / snapshot_bw.dump the same way, :90-97,:149-156)."""
    if not debug:
        return fn(*args)
    saved = _snapshot(args)
    try:
        return fn(*args)
    except Exception:
        torch.save(saved, dump_name)
        print("\nAn error occured in %s. Arguments were written to %s for debugging.\n" % (what, dump_name))
        raise


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                cov3Ds_precomp, viewmatrix, gt_depth, raster_settings):
        rs = raster_settings
        fwd_args = (rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier,
                    cov3Ds_precomp, viewmatrix, gt_depth, rs.projmatrix, rs.tanfovx, rs.tanfovy,
                    rs.image_height, rs.image_width, sh, rs.sh_degree, rs.campos, rs.prefiltered,
                    rs.debug)
        (num_rendered, color, depth, depth_median, depth_var, opacity_map, radii, geom, binning, img,
         gau_uncertainty, gau_related_pixels) = _guarded(
            _C.rasterize_gaussians, fwd_args, rs.debug, "snapshot_fw.dump", "forward")
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, viewmatrix,
                              radii, sh, geom, binning, img, opacity_map, gt_depth)
        return (color, radii, depth, depth_median, depth_var, opacity_map, gau_uncertainty,
                gau_related_pixels)

    @staticmethod
    def backward(ctx, grad_color, grad_radii, grad_depth, grad_depth_median, grad_depth_var,
                 grad_alpha, grad_gau_uncertainty, grad_gau_related_pixels):
        # grad_alpha / grad_gau_* are not propagated (the reference drops them too, :116-146)
        rs = ctx.raster_settings
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, viewmatrix, radii, sh, geom,
         binning, img, opacity_map, gt_depth) = ctx.saved_tensors
        bwd_args = (rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier,
                    cov3Ds_precomp, viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, grad_color,
                    grad_depth, grad_depth_median, grad_depth_var, gt_depth, sh, rs.sh_degree,
                    rs.campos, geom, ctx.num_rendered, binning, img, opacity_map, rs.debug,
                    rs.perspec_matrix, rs.track_off, rs.map_off)
        # dL/dcolors_precomp and dL/dcov3Ds_precomp are only written when autograd wants them (with SH colours
        # and scale + rotation it throws them away: 36 bytes per Gaussian of the per-Gaussian kernel's stores)
        want_colors, want_cov = ctx.needs_input_grad[3], ctx.needs_input_grad[7]
        (g_means2D, g_colors, g_opac, g_means3D, g_cov3D, g_sh, g_scales, g_rot, g_view) = _guarded(
            _C.rasterize_gaussians_backward_select, bwd_args + (want_colors, want_cov), rs.debug,
            "snapshot_bw.dump", "backward")
        if not want_colors:
            g_colors = None
        if not want_cov:
            g_cov3D = None
        # [1,4,4] here, already summed over pixels on the device ([H*W,4,4] + torch.sum upstream,
        # L/__init__.py:160): a view is enough, no reduction kernel
        g_view = g_view[0] if g_view.shape[0] == 1 else torch.sum(g_view, dim=0)
        return (g_means3D, g_means2D, g_sh, g_colors, g_opac, g_scales, g_rot, g_cov3D, g_view,
                None, None)


def set_densify_stats(grad_accum=None, denom=None, max_radii2D=None):
    """Extension (not in the reference): register the mapping loop's densification accumulators
    (Inria 3DGS add_densification_stats: xyz_gradient_accum [P,1], denom [P,1], max_radii2D [P],
    fp32 CUDA) so that every following backward updates them inside its per-Gaussian kernel for the
    visible Gaussians (radii > 0).  Call with no arguments to stop."""
    e = torch.empty(0)
    _C.set_densify_stats(e if grad_accum is None else grad_accum, e if denom is None else denom,
                         e if max_radii2D is None else max_radii2D)


def rgbd_l1_loss(outputs, gt_color, gt_depth, w_color=1.0, w_depth=1.0, w_median=1.0, w_var=1.0,
                 depth_scale=1e-3, depth_mask=False):
    """Extension (not in the reference): the RGB-D L1 mapping loss of one frame and the cotangents of
    the rasterizer's differentiable outputs, in ONE device pass instead of ~30 torch kernels:
        L = w_color sum|C - C_gt| + w_depth sum_m|D - D_gt| + w_median sum_m|D_med - D_gt| + w_var sum depth_var
    `outputs` is the tuple GaussianRasterizer returned; gt_color is uint8 (scaled by 1/255) or fp32
    [3,H,W]; gt_depth int16 (scaled by depth_scale, e.g. millimetres) or fp32 [H,W] / [1,H,W];
    m = pixels with gt depth > 0 when depth_mask.  Returns (loss [1], tensors, cotangents): run the
    backward with  torch.autograd.backward(tensors, cotangents)."""
    color, _radii, depth, median, var = outputs[0], outputs[1], outputs[2], outputs[3], outputs[4]
    loss, g_c, g_d, g_m, g_v = _C.rgbd_l1_loss(color.detach(), depth.detach(), median.detach(), var.detach(),
                                               gt_color, gt_depth, w_color, w_depth, w_median, w_var,
                                               depth_scale, depth_mask)
    return loss, [color, depth, median, var], [g_c, g_d, g_m, g_v]


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                        cov3Ds_precomp, viewmatrix, gt_depth, raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales,
                                     rotations, cov3Ds_precomp, viewmatrix, gt_depth,
                                     raster_settings)


_ABSENT = torch.Tensor([])


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        """Boolean mask of points in front of the near plane of raster_settings.viewmatrix."""
        with torch.no_grad():
            rs = self.raster_settings
            return _C.mark_visible(positions, rs.viewmatrix, rs.projmatrix)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None,
                rotations=None, cov3D_precomp=None, viewmatrix=None, gt_depth=None):
        if (shs is None) == (colors_precomp is None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        pair_missing = scales is None or rotations is None
        pair_given = scales is not None or rotations is not None
        if (pair_missing and cov3D_precomp is None) or (pair_given and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        absent = _ABSENT  # empty CPU tensor = "not provided", as upstream passes torch.Tensor([])
        shs = absent if shs is None else shs
        colors_precomp = absent if colors_precomp is None else colors_precomp
        scales = absent if scales is None else scales
        rotations = absent if rotations is None else rotations
        cov3D_precomp = absent if cov3D_precomp is None else cov3D_precomp
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales,
                                   rotations, cov3D_precomp, viewmatrix, gt_depth,
                                   self.raster_settings)
