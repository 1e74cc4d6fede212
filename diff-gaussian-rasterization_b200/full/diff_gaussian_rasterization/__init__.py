"""diff_gaussian_rasterization (-full surface) backed by the B200-native core.

Public names, argument orders, return tuples and gradient tuples are those of the reference
package (diff-gaussian-rasterization-full/diff_gaussian_rasterization/__init__.py):
  GaussianRasterizationSettings  (:153-165, 12 fields incl. perspec_matrix)
  GaussianRasterizer.forward     (:183-218) -> (color, radii, depth, uncertainty)
  GaussianRasterizer.markVisible (:172-181)
  rasterize_gaussians            (:17-42)
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
    perspec_matrix: torch.Tensor


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                cov3Ds_precomp, viewmatrix, gt_depth, raster_settings):
        rs = raster_settings
        (num_rendered, num_related_gaussians, color, depth, uncertainty, radii, geom, binning,
         img) = _C.rasterize_gaussians(
            rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier,
            cov3Ds_precomp, viewmatrix, gt_depth, rs.projmatrix, rs.tanfovx, rs.tanfovy,
            rs.image_height, rs.image_width, sh, rs.sh_degree, rs.campos, rs.prefiltered)
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.num_related_gaussians = num_related_gaussians
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, viewmatrix,
                              radii, sh, geom, binning, img, gt_depth)
        return color, radii, depth, uncertainty

    @staticmethod
    def backward(ctx, grad_out_color, grad_radii, grad_out_depth, grad_out_uncertainty):
        rs = ctx.raster_settings
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, viewmatrix, radii, sh, geom,
         binning, img, gt_depth) = ctx.saved_tensors
        # dL/dcolors_precomp and dL/dcov3Ds_precomp are only written when autograd wants them (with SH colours
        # and scale + rotation it throws them away: 36 bytes per Gaussian of the per-Gaussian kernel's stores)
        want_colors, want_cov = ctx.needs_input_grad[3], ctx.needs_input_grad[7]
        (g_means2D, g_colors, g_opac, g_means3D, g_cov3D, g_sh, g_scales, g_rot,
         g_view) = _C.rasterize_gaussians_backward_select(
            rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier,
            cov3Ds_precomp, viewmatrix, gt_depth, rs.projmatrix, rs.tanfovx, rs.tanfovy,
            grad_out_color, grad_out_depth, grad_out_uncertainty, sh, rs.sh_degree, rs.campos, geom,
            ctx.num_rendered, binning, img, ctx.num_related_gaussians, rs.perspec_matrix, want_colors, want_cov)
        return (g_means3D, g_means2D, g_sh, g_colors if want_colors else None, g_opac, g_scales, g_rot,
                g_cov3D if want_cov else None, g_view, None, None)


def set_densify_stats(grad_accum=None, denom=None, max_radii2D=None):
    """Extension (not in the reference): register the mapping loop's densification accumulators
    (Inria 3DGS add_densification_stats: xyz_gradient_accum [P,1], denom [P,1], max_radii2D [P],
    fp32 CUDA) so that every following backward updates them inside its per-Gaussian kernel for the
    visible Gaussians (radii > 0).  Call with no arguments to stop."""
    e = torch.empty(0)
    _C.set_densify_stats(e if grad_accum is None else grad_accum, e if denom is None else denom,
                         e if max_radii2D is None else max_radii2D)


def rgbd_l1_loss(outputs, gt_color, gt_depth, w_color=1.0, w_depth=1.0, w_silhouette=1.0,
                 depth_scale=1e-3, depth_mask=False):
    """Extension (not in the reference): the RGB-D L1 mapping loss of one frame and the cotangents of
    the rasterizer's differentiable outputs, in ONE device pass instead of ~30 torch kernels:
        L = w_color sum|C - C_gt| + w_depth sum_m|D - D_gt| + w_silhouette sum(1 - O)
    `outputs` is the tuple GaussianRasterizer returned (color, radii, depth, opacity map O); gt_color
    is uint8 (scaled by 1/255) or fp32 [3,H,W]; gt_depth int16 (scaled by depth_scale, e.g.
    millimetres) or fp32 [H,W] / [1,H,W]; m = pixels with gt depth > 0 when depth_mask.
    Returns (loss [1], tensors, cotangents): run the backward with
    torch.autograd.backward(tensors, cotangents)."""
    color, _radii, depth, omap = outputs[0], outputs[1], outputs[2], outputs[3]
    loss, g_c, g_d, g_o, _ = _C.rgbd_l1_loss(color.detach(), depth.detach(), omap.detach(), omap.detach(),
                                             gt_color, gt_depth, w_color, w_depth, w_silhouette, 0.0,
                                             depth_scale, depth_mask)
    return loss, [color, depth, omap], [g_c, g_d, g_o]


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
