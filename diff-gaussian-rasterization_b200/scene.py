"""Synthetic scene / camera generator shared by the tests, the bench and the oracle driver.

Everything is generated on the CPU in fp32 from a seeded torch.Generator so that the CUDA path,
the reference CUDA build and the CPU oracle see identical bits:
  camera  : fov_x = 60 deg, tanfovy = tanfovx*H/W, znear 0.01, zfar 100, W2C = rotation of
            `rot_deg` about (1,2,3)/|.| + translation (0.05,-0.03,0.1);
            viewmatrix = W2C^T, projmatrix = (P W2C)^T, perspec_matrix = P^T, campos = -R^T t
  points  : camera-space z ~ U[0.5,10], x,y = z*tanfov*U[-1.1,1.1]; 2 % get z ~ U[-1,0.2]
            (near cull); mapped to world space
  scales  : projected sigma_px ~ logU[a,b] pixels, per-axis anisotropy U[0.3,1]
  rot     : random unit quaternions; opacity U[0.05,0.95]; SH dc U[-1.5,1.5], rest N(0,0.1)
  `backdrop`: optionally adds a layer of large opaque splats far away so that every pixel has at
            least one contributor (needed to compare dL/dview with the reference -full build,
            whose ComputePG reads uninitialised shared memory in tiles that contain a pixel with
            no contributor or lie partly outside the image — see DESIGN.md).
"""
import math
from typing import NamedTuple, Optional

import torch

CONFIGS = {
    # name: (P, W, H, sigma_px range)
    "C1": (10_000, 320, 240, (2.0, 12.0)),
    "C2": (100_000, 640, 480, (1.0, 8.0)),
    "C3": (1_000_000, 1920, 1080, (1.0, 8.0)),
    "C4": (5_000_000, 1920, 1080, (0.5, 4.0)),
}


class Camera(NamedTuple):
    W: int
    H: int
    tanfovx: float
    tanfovy: float
    viewmatrix: torch.Tensor      # [4,4] = W2C^T
    projmatrix: torch.Tensor      # [4,4] = (P @ W2C)^T
    perspec_matrix: torch.Tensor  # [4,4] = P^T
    campos: torch.Tensor          # [3]
    w2c: torch.Tensor             # [4,4]


class Scene(NamedTuple):
    means3D: torch.Tensor   # [P,3]
    scales: torch.Tensor    # [P,3]
    rotations: torch.Tensor # [P,4]
    opacities: torch.Tensor # [P,1]
    shs: torch.Tensor       # [P,16,3]
    colors: torch.Tensor    # [P,3] (for the colors_precomp path)
    bg: torch.Tensor        # [3]
    gt_depth: torch.Tensor  # [1,H,W]


def _axis_angle(axis, deg):
    a = torch.tensor(axis, dtype=torch.float64)
    a = a / a.norm()
    th = math.radians(deg)
    K = torch.tensor([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]], dtype=torch.float64)
    return torch.eye(3, dtype=torch.float64) + math.sin(th) * K + (1 - math.cos(th)) * (K @ K)


def make_camera(W, H, rot_deg=5.0, trans=(0.05, -0.03, 0.1), fov_x_deg=60.0, znear=0.01,
                zfar=100.0, seed: Optional[int] = None) -> Camera:
    """seed=None -> the canonical camera; seed=k -> the k-th view of the multi-view config (C5):
    rotation angle / translation jittered deterministically."""
    if seed is not None and seed != 0:
        g = torch.Generator(device="cpu").manual_seed(1000 + seed)
        j = torch.rand(4, generator=g, dtype=torch.float64)
        rot_deg = rot_deg + float(j[0]) * 6.0 - 3.0
        trans = (trans[0] + float(j[1]) * 0.2 - 0.1, trans[1] + float(j[2]) * 0.2 - 0.1,
                 trans[2] + float(j[3]) * 0.2 - 0.1)
    tanfovx = math.tan(math.radians(fov_x_deg) * 0.5)
    tanfovy = tanfovx * H / W
    R = _axis_angle((1.0, 2.0, 3.0), rot_deg)
    w2c = torch.eye(4, dtype=torch.float64)
    w2c[:3, :3] = R
    w2c[:3, 3] = torch.tensor(trans, dtype=torch.float64)
    top, right = tanfovy * znear, tanfovx * znear
    Pm = torch.zeros(4, 4, dtype=torch.float64)
    Pm[0, 0] = 2.0 * znear / (2 * right)
    Pm[1, 1] = 2.0 * znear / (2 * top)
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    campos = -(R.T @ w2c[:3, 3])
    f32 = lambda t: t.to(torch.float32).contiguous()
    return Camera(W, H, float(tanfovx), float(tanfovy), f32(w2c.T), f32((Pm @ w2c).T), f32(Pm.T),
                  f32(campos), f32(w2c))


def make_scene(P, cam: Camera, sigma_px=(1.0, 8.0), seed=0, backdrop=False,
               near_frac=0.02) -> Scene:
    g = torch.Generator(device="cpu").manual_seed(seed)
    U = lambda *s: torch.rand(*s, generator=g, dtype=torch.float32)
    W, H = cam.W, cam.H
    focal_x = W / (2.0 * cam.tanfovx)
    nb = 0
    if backdrop:
        # a grid of big opaque splats at z = 12 covering the frustum
        nbx, nby = max(2, W // 24), max(2, H // 24)
        nb = nbx * nby
    Pm = P - nb
    z = 0.5 + 9.5 * U(Pm)
    near = U(Pm) < near_frac
    z = torch.where(near, -1.0 + 1.2 * U(Pm), z)
    x = z.abs() * cam.tanfovx * (2.2 * U(Pm) - 1.1)
    y = z.abs() * cam.tanfovy * (2.2 * U(Pm) - 1.1)
    sig = math.log(sigma_px[0]) + (math.log(sigma_px[1]) - math.log(sigma_px[0])) * U(Pm)
    sig = torch.exp(sig)
    aniso = 0.3 + 0.7 * U(Pm, 3)
    scales = (sig * z.abs().clamp(min=0.2) / focal_x).unsqueeze(1) * aniso
    opac = 0.05 + 0.9 * U(Pm, 1)
    if nb:
        gx = (torch.arange(nbx, dtype=torch.float32) + 0.5) / nbx * 2 - 1
        gy = (torch.arange(nby, dtype=torch.float32) + 0.5) / nby * 2 - 1
        yy, xx = torch.meshgrid(gy, gx, indexing="ij")
        zb = torch.full((nb,), 12.0) + 0.5 * U(nb)
        xb = zb * cam.tanfovx * xx.reshape(-1)
        yb = zb * cam.tanfovy * yy.reshape(-1)
        sb = (30.0 * zb / focal_x).unsqueeze(1) * torch.ones(nb, 3)
        z, x, y = torch.cat([z, zb]), torch.cat([x, xb]), torch.cat([y, yb])
        scales = torch.cat([scales, sb])
        opac = torch.cat([opac, torch.full((nb, 1), 0.95)])
    p_cam = torch.stack([x, y, z], dim=1)
    R = cam.w2c[:3, :3]
    t = cam.w2c[:3, 3]
    # R^T (p - t), row-vector form.  Written with elementwise ops only (no GEMM, no library
    # reduction): a threaded BLAS may pick a different summation order from run to run, and a 1-ulp
    # difference in one mean is enough to move a splat across a tile boundary (num_rendered +-1).
    dlt = (p_cam - t).double()
    Rd = R.double()
    means = (dlt[:, 0:1] * Rd[0] + dlt[:, 1:2] * Rd[1] + dlt[:, 2:3] * Rd[2]).float()
    q = torch.randn(P, 4, generator=g, dtype=torch.float32)
    qd = q.double()
    q = (qd / torch.sqrt(qd[:, 0:1] ** 2 + qd[:, 1:2] ** 2 + qd[:, 2:3] ** 2 + qd[:, 3:4] ** 2)).float()
    shs = torch.empty(P, 16, 3, dtype=torch.float32)
    shs[:, 0, :] = 3.0 * U(P, 3) - 1.5
    shs[:, 1:, :] = 0.1 * torch.randn(P, 15, 3, generator=g, dtype=torch.float32)
    colors = U(P, 3)
    gt_depth = 0.5 + 9.5 * U(1, H, W)
    bg = torch.tensor([0.1, 0.2, 0.3], dtype=torch.float32)
    return Scene(means.contiguous(), scales.contiguous(), q.contiguous(), opac.contiguous(),
                 shs.contiguous(), colors.contiguous(), bg, gt_depth.contiguous())


def make_cotangents(cam: Camera, n_aux, seed=1):
    """Seeded N(0,1) cotangents: colour [3,H,W] and n_aux single-channel maps [1,H,W]."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    col = torch.randn(3, cam.H, cam.W, generator=g, dtype=torch.float32)
    aux = [torch.randn(1, cam.H, cam.W, generator=g, dtype=torch.float32) for _ in range(n_aux)]
    return col, aux


def config(name, backdrop=False, seed=0):
    P, W, H, sig = CONFIGS[name]
    cam = make_camera(W, H)
    return cam, make_scene(P, cam, sig, seed=seed, backdrop=backdrop)
