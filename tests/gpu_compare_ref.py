"""GPU-box script (not a pytest file): compare the B200-native path with the reference CUDA build
(baseline/_ref) on identical inputs, decode per-Gaussian intermediates for a bit-equality census,
and time both.  Usage (under gpurun):
    python tests/gpu_compare_ref.py [--configs small,C1,C2] [--time C3] [--out gpurun_out/cmp.txt]
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import parity_util as pu  # noqa: E402
ge = pu.ge


def scene_for(name, backdrop):
    sc = ge.load_scene_module()
    if name == "small":
        cam = sc.make_camera(128, 96)
        return cam, sc.make_scene(3000, cam, (2.0, 10.0), seed=5, backdrop=backdrop)
    if name == "mid":
        cam = sc.make_camera(320, 240)
        return cam, sc.make_scene(20000, cam, (2.0, 12.0), seed=6, backdrop=backdrop)
    return sc.config(name, backdrop=backdrop)


def geom_census(variant, cam, scene, log):
    """Bit-equality census of per-Gaussian forward state: ours (gsr_decode_geometry) against the
    reference's geomBuffer decoded with the layout of rasterizer_impl.cu:156-171."""
    ref = ge.load_reference(variant)
    mine = ge.load_variant(variant)
    if ref is None:
        return
    dev = "cuda:0"
    P = scene.means3D.shape[0]
    args = lambda: (scene.bg.to(dev), scene.means3D.to(dev), torch.Tensor([]), scene.opacities.to(dev),
                    scene.scales.to(dev), scene.rotations.to(dev), 1.0, torch.Tensor([]),
                    cam.viewmatrix.to(dev), scene.gt_depth.to(dev), cam.projmatrix.to(dev), cam.tanfovx,
                    cam.tanfovy, cam.H, cam.W, scene.shs.to(dev), 3, cam.campos.to(dev), False)
    if variant == "light":
        r = ref._C.rasterize_gaussians(*args(), False)
        m = mine._C.rasterize_gaussians(*args(), False)
        r_radii, r_geom, m_radii, m_geom = r[6], r[7], m[6], m[7]
        log("num_rendered ref %d ours %d" % (r[0], m[0]))
    else:
        r = ref._C.rasterize_gaussians(*args())
        m = mine._C.rasterize_gaussians(*args())
        r_radii, r_geom, m_radii, m_geom = r[5], r[6], m[5], m[6]
        log("num_rendered ref %d ours %d ; num_related ref %d ours %d" % (r[0], m[0], r[1], m[1]))
    torch.cuda.synchronize()
    # decode reference geomBuffer: arrays start at 128-byte boundaries relative to the base ptr
    base = r_geom.data_ptr()
    raw = r_geom.cpu().numpy()
    off = 0

    def take(nbytes):
        nonlocal off
        addr = (base + off + 127) // 128 * 128
        off = addr - base
        out = raw[off:off + nbytes]
        off += nbytes
        return out
    depths = take(4 * P).view(np.float32)
    clamped = take(3 * P).view(np.uint8).reshape(P, 3)
    iradii = take(4 * P).view(np.int32)
    means2D = take(8 * P).view(np.float32).reshape(P, 2)
    cov3D = take(24 * P).view(np.float32).reshape(P, 6)
    conic_o = take(16 * P).view(np.float32).reshape(P, 4)
    rgb = take(12 * P).view(np.float32).reshape(P, 3)
    tiles = take(4 * P).view(np.uint32)
    # ours
    lib = ctypes.CDLL(ge.core_library_path())
    f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
    o_depth, o_m2d, o_co, o_rgb, o_cov = f(P), f(P, 2), f(P, 4), f(P, 3), f(P, 6)
    o_tiles = torch.empty(P, dtype=torch.int32, device=dev)
    o_cl = torch.empty(P, 3, dtype=torch.uint8, device=dev)
    vp = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = lib.gsr_decode_geometry(vp(m_geom), P, vp(o_depth), vp(o_m2d), vp(o_co), vp(o_rgb), vp(o_cov),
                                 vp(o_tiles), vp(o_cl), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    rr, mr = r_radii.cpu().numpy(), m_radii.cpu().numpy()
    vis = rr > 0
    log("radii mismatches %d / %d (visible %d)" % ((rr != mr).sum(), P, vis.sum()))
    both = vis & (mr > 0)

    def bits(name, a, b):
        a = np.ascontiguousarray(a[both]).view(np.uint32)
        b = np.ascontiguousarray(b[both]).view(np.uint32)
        ne = (a != b)
        d = np.abs(a.astype(np.int64) - b.astype(np.int64))
        log("  %-14s bit-different %8d / %d  max ulp %d" % (name, ne.sum(), a.size, d.max() if d.size else 0))
    bits("depth", depths, o_depth.cpu().numpy())
    bits("means2D", means2D, o_m2d.cpu().numpy())
    bits("cov3D", cov3D, o_cov.cpu().numpy())
    bits("conic_opacity", conic_o, o_co.cpu().numpy())
    bits("rgb", rgb, o_rgb.cpu().numpy())
    log("  tiles_touched mismatches %d ; clamped mismatches %d" % (
        (tiles[both] != o_tiles.cpu().numpy().view(np.uint32)[both]).sum(),
        (clamped[both].astype(bool) != o_cl.cpu().numpy()[both].astype(bool)).sum()))


def timeit(fn, warm=3, it=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(it):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def time_variant(mod, variant, cam, scene, cot, log, tag):
    dev = "cuda:0"
    d = lambda t, rg=False: t.to(dev).clone().requires_grad_(rg)
    means3D, opac, scales, rots, shs, view = (d(scene.means3D, True), d(scene.opacities, True),
                                              d(scene.scales, True), d(scene.rotations, True),
                                              d(scene.shs, True), d(cam.viewmatrix, True))
    means2D = torch.zeros_like(means3D, requires_grad=True)
    gt = scene.gt_depth.to(dev)
    rs = pu.settings_for(mod, variant, cam, scene, dev)
    rast = mod.GaussianRasterizer(rs)
    ccol = cot[0].to(dev)
    caux = [c.to(dev) for c in cot[1]]
    state = {}

    def fwd():
        state["res"] = rast(means3D=means3D, means2D=means2D, opacities=opac, shs=shs, scales=scales,
                            rotations=rots, viewmatrix=view, gt_depth=gt)

    def fwdbwd():
        fwd()
        res = state["res"]
        if variant == "light":
            outs, cots = [res[0], res[2], res[3], res[4]], [ccol] + caux[:3]
        else:
            outs, cots = [res[0], res[2], res[3]], [ccol] + caux[:2]
        torch.autograd.backward(outs, cots)
        for t in (means3D, means2D, opac, scales, rots, shs, view):
            t.grad = None
    with torch.no_grad():
        f_med, f_min = timeit(fwd)
    fb_med, fb_min = timeit(fwdbwd)
    log("%s %s: forward %.3f ms (min %.3f)  fwd+bwd %.3f ms (min %.3f)  -> %.1f frames/s" % (
        tag, variant, f_med, f_min, fb_med, fb_min, 1000.0 / fb_med))
    return dict(fwd_ms=f_med, fwdbwd_ms=fb_med)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="small,mid")
    ap.add_argument("--time", default="")
    ap.add_argument("--variants", default="light,full")
    ap.add_argument("--out", default="gpurun_out/compare_ref.txt")
    ap.add_argument("--no-backdrop", action="store_true")
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    fh = open(a.out, "a")

    def log(s):
        print(s)
        fh.write(s + "\n")
        fh.flush()
    log("== gpu_compare_ref %s  (%s) ==" % (time.strftime("%F %T"), torch.cuda.get_device_name(0)))
    sc = ge.load_scene_module()
    results = {}
    for name in [c for c in a.configs.split(",") if c]:
        for backdrop in ([False] if a.no_backdrop else [True, False]):
            cam, scene = scene_for(name, backdrop)
            for variant in a.variants.split(","):
                log("--- config %s backdrop=%s variant=%s  P=%d %dx%d" % (
                    name, backdrop, variant, scene.means3D.shape[0], cam.W, cam.H))
                cot = sc.make_cotangents(cam, 3 if variant == "light" else 2)
                ref = ge.load_reference(variant)
                mine = ge.load_variant(variant)
                try:
                    geom_census(variant, cam, scene, log)
                except Exception as e:  # keep going: census is informational
                    log("geom census failed: %r" % (e,))
                modes = [(False, False)]
                if variant == "light" and name == "small":
                    modes += [(True, False), (False, True)]
                for track_off, map_off in modes:
                    o_m, g_m = pu.run_variant(mine, variant, cam, scene, cot, track_off=track_off, map_off=map_off)
                    if ref is not None:
                        o_r, g_r = pu.run_variant(ref, variant, cam, scene, cot, track_off=track_off, map_off=map_off)
                        ok, lines = pu.compare_runs(o_m, g_m, o_r, g_r)
                        log("vs reference (track_off=%s map_off=%s): %s" % (track_off, map_off, "OK" if ok else "MISMATCH"))
                        for ln in lines:
                            log("   " + ln)
                        log("   dL_dview ours:\n%s\n   dL_dview ref:\n%s" % (g_m["viewmatrix"], g_r["viewmatrix"]))
    for name in [c for c in a.time.split(",") if c]:
        cam, scene = scene_for(name, False)
        for variant in a.variants.split(","):
            cot = sc.make_cotangents(cam, 3 if variant == "light" else 2)
            mine = ge.load_variant(variant)
            results["%s_%s_ours" % (name, variant)] = time_variant(mine, variant, cam, scene, cot, log, "ours " + name)
            ref = ge.load_reference(variant)
            if ref is not None:
                try:
                    results["%s_%s_ref" % (name, variant)] = time_variant(ref, variant, cam, scene, cot, log, "ref  " + name)
                except Exception as e:
                    log("reference timing failed: %r" % (e,))
    log(json.dumps(results))


if __name__ == "__main__":
    main()
