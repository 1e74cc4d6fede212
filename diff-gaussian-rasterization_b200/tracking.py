"""Camera-pose tracking on the device.

`PoseTracker` drives the C-ABI tracker of libgsr_b200.so (include/gsr_b200.h, csrc/tracker.cu):
K iterations of CG-SLAM's tracking loop — render(-light, map_off) -> masked L1 colour + depth loss
-> backward -> dL/dviewmatrix -> quaternion / translation gradient -> Adam — replayed from one CUDA
graph with no host interaction.

`torch_tracking_loop` is the same loop written the way a CG-SLAM-style caller writes it today:
`GaussianRasterizer` of a `diff_gaussian_rasterization` package (ours or the reference build), the
loss in torch, autograd through the pose parametrisation and `torch.optim.Adam`.  It is the
tracker's parity reference (tests/test_tracking_gpu.py) and its baseline (tools/bench_tracking.py).

Pose: world-to-camera W2C = [R(q/|q|) t; 0 1], q = (w, x, y, z); the rasterizer's `viewmatrix` is
W2C^T, `projmatrix` = viewmatrix @ perspec_matrix, `campos` = -R^T t.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class TrackParams(ctypes.Structure):
    _fields_ = [("w_color", ctypes.c_float), ("w_depth", ctypes.c_float),
                ("alpha_thresh", ctypes.c_float), ("use_depth_mask", ctypes.c_int),
                ("lr_rot", ctypes.c_float), ("lr_trans", ctypes.c_float),
                ("beta1", ctypes.c_float), ("beta2", ctypes.c_float), ("eps", ctypes.c_float)]


class TrackResult(ctypes.Structure):
    _fields_ = [("q", ctypes.c_float * 4), ("t", ctypes.c_float * 3),
                ("last_dL_dview", ctypes.c_float * 16), ("last_grad", ctypes.c_float * 7),
                ("last_twist_grad", ctypes.c_float * 6),
                ("iterations", ctypes.c_int), ("num_rendered", ctypes.c_int),
                ("retries", ctypes.c_int), ("kernels_per_iteration", ctypes.c_int)]


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "lib", "libgsr_b200.so")
        if not os.path.exists(path):
            raise ImportError("libgsr_b200.so has not been built (python __graft_entry__.py); "
                              "there is no fallback")
        lib = ctypes.CDLL(path)
        vp, fp, ci, cf = ctypes.c_void_p, ctypes.POINTER(ctypes.c_float), ctypes.c_int, ctypes.c_float
        lib.gsr_tracker_create.restype = vp
        lib.gsr_tracker_create.argtypes = [ci, ci, ci, ci, ci, cf, cf, fp, ci]
        lib.gsr_tracker_destroy.restype = None
        lib.gsr_tracker_destroy.argtypes = [vp]
        lib.gsr_tracker_set_scene.argtypes = [vp, vp, vp, vp, vp, vp, cf, vp, vp, vp]
        lib.gsr_tracker_set_frame.argtypes = [vp, vp, vp]
        lib.gsr_tracker_set_pose.argtypes = [vp, fp, fp]
        lib.gsr_tracker_run.argtypes = [vp, ctypes.POINTER(TrackParams), ci, fp,
                                        ctypes.POINTER(TrackResult), vp]
        lib.gsr_last_error.restype = ctypes.c_char_p
        _LIB = lib
    return _LIB


def _check(rc):
    if rc != 0:
        raise RuntimeError("libgsr_b200: " + _lib().gsr_last_error().decode())


def _dev_ptr(t, name, shape=None):
    if t is None:
        return None
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise ValueError("%s must be a contiguous fp32 CUDA tensor" % name)
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError("%s has shape %s, expected %s" % (name, tuple(t.shape), tuple(shape)))
    return ctypes.c_void_p(t.data_ptr())


def default_params(w_color=0.5, w_depth=1.0, alpha_thresh=0.99, use_depth_mask=True,
                   lr_rot=4e-4, lr_trans=2e-3, beta1=0.9, beta2=0.999, eps=1e-8):
    return dict(w_color=w_color, w_depth=w_depth, alpha_thresh=alpha_thresh,
                use_depth_mask=bool(use_depth_mask), lr_rot=lr_rot, lr_trans=lr_trans,
                beta1=beta1, beta2=beta2, eps=eps)


class PoseTracker:
    """Device-resident pose tracker.  Tensors handed to set_scene / set_frame are borrowed: keep
    them alive and in place (copy new frame data INTO the same tensors to reuse the captured graph)."""

    def __init__(self, num_gaussians, sh_degree, num_sh_coeffs, height, width, tanfovx, tanfovy,
                 perspec_matrix, max_iterations=256):
        persp = perspec_matrix.detach().to("cpu", torch.float32).contiguous().view(-1)
        arr = (ctypes.c_float * 16)(*persp.tolist())
        self._h = _lib().gsr_tracker_create(int(num_gaussians), int(sh_degree), int(num_sh_coeffs),
                                            int(width), int(height), float(tanfovx), float(tanfovy),
                                            arr, int(max_iterations))
        if not self._h:
            raise RuntimeError("libgsr_b200: " + _lib().gsr_last_error().decode())
        self.P, self.H, self.W = int(num_gaussians), int(height), int(width)
        self.M = int(num_sh_coeffs)
        self.max_iterations = int(max_iterations)
        self._keep = {}
        self._device = torch.device("cuda", torch.cuda.current_device())

    def close(self):
        if getattr(self, "_h", None):
            _lib().gsr_tracker_destroy(self._h)
            self._h = None

    __del__ = close

    def set_scene(self, means3D, opacities, shs=None, colors_precomp=None, scales=None,
                  rotations=None, cov3D_precomp=None, bg=None, scale_modifier=1.0):
        P = self.P
        self._device = means3D.device
        self._keep["scene"] = (means3D, opacities, shs, colors_precomp, scales, rotations,
                               cov3D_precomp, bg)
        _check(_lib().gsr_tracker_set_scene(
            self._h, _dev_ptr(means3D, "means3D", (P, 3)),
            _dev_ptr(shs, "shs", (P, self.M, 3)) if shs is not None else None,
            _dev_ptr(colors_precomp, "colors_precomp", (P, 3)) if colors_precomp is not None else None,
            _dev_ptr(opacities, "opacities", (P, 1)), _dev_ptr(scales, "scales"),
            float(scale_modifier), _dev_ptr(rotations, "rotations"),
            _dev_ptr(cov3D_precomp, "cov3D_precomp"), _dev_ptr(bg, "bg", (3,))))

    def set_frame(self, gt_color, gt_depth):
        self._keep["frame"] = (gt_color, gt_depth)
        if gt_depth.dim() == 3:
            gt_depth = gt_depth[0]
        _check(_lib().gsr_tracker_set_frame(self._h, _dev_ptr(gt_color, "gt_color", (3, self.H, self.W)),
                                            _dev_ptr(gt_depth, "gt_depth", (self.H, self.W))))

    def set_pose(self, quat_wxyz, trans):
        q = (ctypes.c_float * 4)(*[float(v) for v in quat_wxyz])
        t = (ctypes.c_float * 3)(*[float(v) for v in trans])
        _check(_lib().gsr_tracker_set_pose(self._h, q, t))

    def run(self, iterations, **params):
        """-> dict(q, t, loss (list per iteration), last_dL_dview, last_grad, num_rendered, retries)."""
        prm = default_params(**params)
        cp = TrackParams(prm["w_color"], prm["w_depth"], prm["alpha_thresh"], int(prm["use_depth_mask"]),
                         prm["lr_rot"], prm["lr_trans"], prm["beta1"], prm["beta2"], prm["eps"])
        hist = (ctypes.c_float * int(iterations))()
        res = TrackResult()
        # the tracker's private stream is ordered after the CURRENT torch stream: the scene / frame
        # tensors are usually the result of asynchronous torch work queued there (gsr_b200.h)
        cur = torch.cuda.current_stream(self._device).cuda_stream
        _check(_lib().gsr_tracker_run(self._h, ctypes.byref(cp), int(iterations), hist, ctypes.byref(res),
                                      ctypes.c_void_p(cur)))
        return dict(q=list(res.q), t=list(res.t), loss=list(hist), last_dL_dview=list(res.last_dL_dview),
                    last_grad=list(res.last_grad), last_twist_grad=list(res.last_twist_grad),
                    num_rendered=res.num_rendered, retries=res.retries,
                    kernels_per_iteration=res.kernels_per_iteration)


# ---- the same loop through the reference surface (parity reference + baseline) -----------------

def quat_to_rotation(q):
    """R(q/|q|) for q = (w, x, y, z), differentiable."""
    q = q / q.norm()
    w, x, y, z = q[0], q[1], q[2], q[3]
    return torch.stack([
        torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)]),
        torch.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)]),
        torch.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)])])


def rotation_to_quat(R):
    """(w, x, y, z) of a rotation matrix (numerically safe branch on the largest diagonal term)."""
    R = R.detach().to(torch.float64).cpu()
    tr = R[0, 0] + R[1, 1] + R[2, 2]
    if tr > 0:
        s = torch.sqrt(tr + 1.0) * 2
        q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = torch.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = [(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s]
    elif R[1, 1] > R[2, 2]:
        s = torch.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = [(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s]
    else:
        s = torch.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = [(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s]
    return torch.tensor([float(v) for v in q], dtype=torch.float32)


def torch_tracking_loop(mod, scene, gt_color, gt_depth, H, W, tanfovx, tanfovy, perspec_matrix,
                        quat_wxyz, trans, iterations, sh_degree=3, use_sh=True, **params):
    """CG-SLAM-style tracking loop through `mod` (a -light `diff_gaussian_rasterization` package).
    scene: dict of CUDA tensors (means3D, opacities, scales, rotations, shs | colors, bg).
    Returns dict(q, t, loss, grads (per iteration: dL/dq, dL/dt), dL_dview (per iteration))."""
    prm = default_params(**params)
    dev = gt_color.device
    q = torch.tensor([float(v) for v in quat_wxyz], device=dev, requires_grad=True)
    t = torch.tensor([float(v) for v in trans], device=dev, requires_grad=True)
    opt = torch.optim.Adam([{"params": [q], "lr": prm["lr_rot"]}, {"params": [t], "lr": prm["lr_trans"]}],
                           betas=(prm["beta1"], prm["beta2"]), eps=prm["eps"])
    perspT = perspec_matrix.to(dev)
    gt_d = gt_depth if gt_depth.dim() == 3 else gt_depth[None]
    means3D = scene["means3D"]
    means2D = torch.zeros_like(means3D)
    losses, grads, dviews = [], [], []
    bottom = torch.tensor([[0.0, 0.0, 0.0, 1.0]], device=dev)
    for _ in range(iterations):
        opt.zero_grad(set_to_none=True)
        R = quat_to_rotation(q)
        w2c = torch.cat([torch.cat([R, t[:, None]], dim=1), bottom], dim=0)
        viewmatrix = w2c.t().contiguous()
        viewmatrix.retain_grad()
        with torch.no_grad():
            vm = viewmatrix.detach()
            projmatrix = (vm @ perspT).contiguous()
            campos = (-(vm[:3, :3] @ vm[3, :3])).contiguous()
        rs = mod.GaussianRasterizationSettings(
            image_height=H, image_width=W, tanfovx=tanfovx, tanfovy=tanfovy, bg=scene["bg"],
            scale_modifier=1.0, viewmatrix=vm, projmatrix=projmatrix, sh_degree=sh_degree,
            campos=campos, prefiltered=False, debug=False, perspec_matrix=perspT, track_off=False,
            map_off=True)
        rast = mod.GaussianRasterizer(rs)
        color, radii, depth, dmed, dvar, alpha, gunc, gpx = rast(
            means3D=means3D, means2D=means2D, opacities=scene["opacities"],
            shs=scene["shs"] if use_sh else None, colors_precomp=None if use_sh else scene["colors"],
            scales=scene["scales"], rotations=scene["rotations"], cov3D_precomp=None,
            viewmatrix=viewmatrix, gt_depth=gt_d)
        with torch.no_grad():
            mask = alpha > prm["alpha_thresh"]
            if prm["use_depth_mask"]:
                mask = mask & (gt_d > 0)
            mask = mask.to(color.dtype)
        loss = (prm["w_color"] * (mask * (color - gt_color).abs()).sum() +
                prm["w_depth"] * (mask * (depth - gt_d).abs()).sum())
        loss.backward()
        losses.append(loss.detach())  # no per-iteration host sync; read back once at the end
        grads.append(torch.cat([q.grad, t.grad]).detach())
        dviews.append(viewmatrix.grad.detach().reshape(-1).clone())
        opt.step()
    return dict(q=q.detach().cpu().tolist(), t=t.detach().cpu().tolist(),
                loss=torch.stack(losses).cpu().tolist(), grads=[g.cpu() for g in grads],
                dL_dview=[v.cpu() for v in dviews])
