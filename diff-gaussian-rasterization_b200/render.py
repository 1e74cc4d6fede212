"""Layer-4 `render()` helper with CG-SLAM's two call signatures (reference README.md:33,71) and
output-dict keys (README.md:46-51 full, :86-95 light), plus the camera utilities that turn a
world-to-camera matrix into the tensors the rasterizer expects.

The reference repository does not contain this function (it lives in CG-SLAM); it is provided so
that the drop-in claim can be exercised end to end.  Conventions follow Inria 3DGS, which CG-SLAM
builds on:
  gaussians      : object with get_xyz [P,3], get_opacity [P,1], get_scaling [P,3],
                   get_rotation [P,4], get_features [P,M,3], active_sh_degree
  viewpoint_cam  : object with projection_matrix [4,4] (= P^T, OpenGL-style perspective, transposed)
                   and optionally znear / zfar (used when projection_matrix is absent)
  pipe           : object with optional .debug / .compute_cov3D_python / .convert_SHs_python
  viewmatrix     : w2c^T  ([4,4], requires_grad for tracking)
  fov            : (tan(fov_x/2), tan(fov_y/2));  HW: (H, W)
"""
import math

import torch


def perspective_matrix(tanfovx, tanfovy, znear=0.01, zfar=100.0, device="cpu"):
    """Inria 3DGS getProjectionMatrix (z in [0,1], w = z_cam); returned NOT transposed."""
    P = torch.zeros(4, 4, dtype=torch.float32, device=device)
    P[0, 0] = 1.0 / tanfovx
    P[1, 1] = 1.0 / tanfovy
    P[3, 2] = 1.0
    P[2, 2] = zfar / (zfar - znear)
    P[2, 3] = -(zfar * znear) / (zfar - znear)
    return P


def camera_tensors(w2cT, perspecT):
    """(projmatrix, campos) for a transposed world-to-camera matrix and transposed perspective.
    projmatrix = (P @ w2c)^T = w2c^T @ P^T ; campos = -R^T t = last row of inverse(w2c^T)."""
    with torch.no_grad():
        proj = w2cT @ perspecT
        campos = torch.linalg.inv(w2cT)[3, :3].contiguous()
    return proj.contiguous(), campos


def _pick(obj, name, default=None):
    return getattr(obj, name, default) if obj is not None else default


def render(viewpoint_cam, gaussians, pipe, bg_color, viewmatrix=None, fov=None, HW=None,
           gt_depth=None, track_off=None, map_off=None, scaling_modifier=1.0, override_color=None,
           variant=None, rasterizer_module=None):
    """Render one view.  The -light surface is used when track_off / map_off are given (or
    variant == 'light'), the -full surface otherwise — matching which of the two reference packages
    accepts those arguments.  `rasterizer_module` lets the caller pass an already imported
    `diff_gaussian_rasterization` package (ours or the reference's)."""
    if variant is None:
        variant = "light" if (track_off is not None or map_off is not None) else "full"
    if rasterizer_module is None:
        import diff_gaussian_rasterization as rasterizer_module  # whichever variant is installed
    mod = rasterizer_module
    H, W = int(HW[0]), int(HW[1])
    tanfovx, tanfovy = float(fov[0]), float(fov[1])
    dev = viewmatrix.device
    perspecT = _pick(viewpoint_cam, "projection_matrix")
    if perspecT is None:
        perspecT = perspective_matrix(tanfovx, tanfovy, _pick(viewpoint_cam, "znear", 0.01),
                                      _pick(viewpoint_cam, "zfar", 100.0), dev).t().contiguous()
    perspecT = perspecT.to(dev)
    projmatrix, campos = camera_tensors(viewmatrix.detach(), perspecT)

    means3D = gaussians.get_xyz
    screenspace_points = torch.zeros_like(means3D, requires_grad=True)
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    kw = dict(image_height=H, image_width=W, tanfovx=tanfovx, tanfovy=tanfovy, bg=bg_color,
              scale_modifier=scaling_modifier, viewmatrix=viewmatrix.detach(), projmatrix=projmatrix,
              sh_degree=int(_pick(gaussians, "active_sh_degree", 0)), campos=campos, prefiltered=False,
              perspec_matrix=perspecT)
    if variant == "light":
        kw.update(debug=bool(_pick(pipe, "debug", False)), track_off=bool(track_off), map_off=bool(map_off))
    rasterizer = mod.GaussianRasterizer(mod.GaussianRasterizationSettings(**kw))

    shs, colors = (None, override_color) if override_color is not None else (gaussians.get_features, None)
    res = rasterizer(means3D=means3D, means2D=screenspace_points, opacities=gaussians.get_opacity,
                     shs=shs, colors_precomp=colors, scales=gaussians.get_scaling,
                     rotations=gaussians.get_rotation, cov3D_precomp=None, viewmatrix=viewmatrix,
                     gt_depth=gt_depth)
    if variant == "light":
        color, radii, depth, depth_median, depth_var, opacity_map, gau_unc, gau_px = res
        out = {"render": color, "depth": depth, "depth_median": depth_median, "opacity_map": opacity_map,
               "depth_var": depth_var, "gau_uncertainty": gau_unc, "num_related_pixels": gau_px}
    else:
        color, radii, depth, uncertainty = res
        out = {"render": color, "depth": depth, "opacity_map": uncertainty}
    # Inria-style extras callers of render() conventionally read
    out.update({"viewspace_points": screenspace_points, "visibility_filter": radii > 0, "radii": radii})
    return out


def fov_from_focal(focal, pixels):
    return 2.0 * math.atan(pixels / (2.0 * focal))
