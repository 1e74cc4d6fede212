"""Generates the golden vectors under tests/golden/ by running the UNMODIFIED reference CUDA build
(baseline/_ref, built by baseline/build_ref.sh from /root/reference) on a B200.

The reference ships no tests or fixtures, so these files are what pins the oracle
and the CUDA path to the reference's actual behaviour.  Each case_<name>.npz holds the inputs
(bit-exact fp32), every forward output, every gradient the Python wrapper returns and the
per-Gaussian forward state decoded from the reference's geomBuffer
(cuda_rasterizer/rasterizer_impl.cu:156-171 layout).

Run on the GPU box:   python tests/golden/make_golden.py --out gpurun_out/golden
then copy gpurun_out/golden/*.npz into tests/golden/ and commit.
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import parity_util as pu  # noqa: E402

ge = pu.ge

# name: (variant, P, W, H, sigma range, seed, backdrop, use_sh, sh_degree, cov_precomp, modes)
# -full cases use a backdrop and 16-aligned sizes: its ComputePG kernel reads uninitialised shared
# memory in tiles with an empty or out-of-image pixel, so dL_dview of the reference
# is only well defined when every pixel of every tile has a contributor.
CASES = {
    "light_sh3":     ("light", 600, 96, 64, (2.0, 10.0), 3, False, True, 3, False,
                      [(False, False), (True, False), (False, True)]),
    "light_ragged":  ("light", 500, 100, 70, (2.0, 10.0), 4, False, True, 3, False, [(False, False)]),
    "light_sh1":     ("light", 400, 64, 48, (2.0, 8.0), 5, False, True, 1, False, [(False, False)]),
    "light_rgb_cov": ("light", 400, 64, 48, (2.0, 8.0), 6, False, False, 0, True, [(False, False)]),
    "full_sh3":      ("full", 600, 96, 64, (2.0, 10.0), 3, True, True, 3, False, [(False, False)]),
    "full_sh2":      ("full", 400, 64, 48, (2.0, 8.0), 7, True, True, 2, False, [(False, False)]),
    "full_rgb_cov":  ("full", 400, 64, 48, (2.0, 8.0), 8, True, False, 0, True, [(False, False)]),
}


def cov3d_of(scene):
    """fp32 3D covariances [P,6] from scale + rotation (any consistent rounding will do: the
    result is stored and fed to every implementation as the `cov3D_precomp` input)."""
    q = scene.rotations.double()
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                     2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1).reshape(-1, 3, 3)
    S = torch.diag_embed(scene.scales.double())
    Sig = R @ S @ S @ R.transpose(1, 2)
    c = torch.stack([Sig[:, 0, 0], Sig[:, 0, 1], Sig[:, 0, 2], Sig[:, 1, 1], Sig[:, 1, 2], Sig[:, 2, 2]], 1)
    return c.float().contiguous()


def build_case(name):
    variant, P, W, H, sig, seed, backdrop, use_sh, deg, covp, modes = CASES[name]
    sc = ge.load_scene_module()
    cam = sc.make_camera(W, H)
    scene = sc.make_scene(P, cam, sig, seed=seed, backdrop=backdrop)
    cot = sc.make_cotangents(cam, 3 if variant == "light" else 2, seed=seed + 100)
    cov = cov3d_of(scene) if covp else None
    return variant, cam, scene, cot, use_sh, deg, cov, modes


def decode_ref_geom(geom, P):
    base = geom.data_ptr()
    raw = geom.cpu().numpy()
    off = 0
    out = {}

    def take(nbytes):
        nonlocal off
        addr = (base + off + 127) // 128 * 128
        off = addr - base
        a = raw[off:off + nbytes]
        off += nbytes
        return a
    out["depth"] = take(4 * P).view(np.float32).copy()
    out["clamped"] = take(3 * P).view(np.uint8).reshape(P, 3).copy()
    take(4 * P)  # internal radii
    out["means2D"] = take(8 * P).view(np.float32).reshape(P, 2).copy()
    out["cov3D"] = take(24 * P).view(np.float32).reshape(P, 6).copy()
    out["conic_opacity"] = take(16 * P).view(np.float32).reshape(P, 4).copy()
    out["rgb"] = take(12 * P).view(np.float32).reshape(P, 3).copy()
    out["tiles_touched"] = take(4 * P).view(np.uint32).copy()
    return out


def ref_geometry(ref, variant, cam, scene, use_sh, deg, cov):
    dev = "cuda:0"
    E = torch.Tensor([])
    d = lambda t: t.to(dev)
    args = (d(scene.bg), d(scene.means3D), E if use_sh else d(scene.colors), d(scene.opacities),
            E if cov is not None else d(scene.scales), E if cov is not None else d(scene.rotations), 1.0,
            d(cov) if cov is not None else E, d(cam.viewmatrix), d(scene.gt_depth), d(cam.projmatrix),
            cam.tanfovx, cam.tanfovy, cam.H, cam.W, d(scene.shs) if use_sh else E, deg, d(cam.campos), False)
    if variant == "light":
        r = ref._C.rasterize_gaussians(*args, False)
        geom, nr, ng = r[7], r[0], 0
    else:
        r = ref._C.rasterize_gaussians(*args)
        geom, nr, ng = r[6], r[0], r[1]
    torch.cuda.synchronize()
    g = decode_ref_geom(geom, scene.means3D.shape[0])
    return g, int(nr), int(ng)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join("gpurun_out", "golden"))
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    for name in CASES:
        variant, cam, scene, cot, use_sh, deg, cov, modes = build_case(name)
        ref = ge.load_reference(variant)
        assert ref is not None, "baseline/_ref is missing: run baseline/build_ref.sh first"
        blob = dict(
            in_means3D=scene.means3D.numpy(), in_scales=scene.scales.numpy(),
            in_rotations=scene.rotations.numpy(), in_opacities=scene.opacities.numpy(),
            in_shs=scene.shs.numpy(), in_colors=scene.colors.numpy(), in_bg=scene.bg.numpy(),
            in_gt_depth=scene.gt_depth.numpy(), in_viewmatrix=cam.viewmatrix.numpy(),
            in_projmatrix=cam.projmatrix.numpy(), in_perspec=cam.perspec_matrix.numpy(),
            in_campos=cam.campos.numpy(), in_w2c=cam.w2c.numpy(),
            in_cot_color=cot[0].numpy(), in_cot_aux=np.stack([c.numpy() for c in cot[1]]),
            meta=np.array([cam.W, cam.H, deg, int(use_sh), int(cov is not None)], dtype=np.int64),
            tanfov=np.array([cam.tanfovx, cam.tanfovy], dtype=np.float64))
        if cov is not None:
            blob["in_cov3D"] = cov.numpy()
        geom, nr, ng = ref_geometry(ref, variant, cam, scene, use_sh, deg, cov)
        for k, v in geom.items():
            blob["geom_" + k] = v
        blob["num_rendered"] = np.array([nr, ng], dtype=np.int64)
        for (track_off, map_off) in modes:
            tag = "m%d%d_" % (int(track_off), int(map_off))
            outs, grads = pu.run_variant(ref, variant, cam, scene, cot, use_sh=use_sh, sh_degree=deg,
                                         track_off=track_off, map_off=map_off, cov_precomp=cov)
            for k, v in outs.items():
                blob[tag + "out_" + k] = v
            for k, v in grads.items():
                if v is not None:
                    blob[tag + "grad_" + k] = v
        path = os.path.join(a.out, "case_%s.npz" % name)
        np.savez_compressed(path, **blob)
        print("wrote %s (%d KB) num_rendered=%d num_related=%d" % (path, os.path.getsize(path) // 1024, nr, ng))


if __name__ == "__main__":
    main()
