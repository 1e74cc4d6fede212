#!/usr/bin/env python
"""Randomised parity sweep against the reference CUDA build (baseline/_ref), live on the GPU box: N random
cases — variant, SH degree, SH / precomputed colours, scale + rotation / precomputed covariance, Gaussian count,
image size (ragged sizes included), splat size range, backdrop, -light's track_off / map_off modes — each run
through both packages on identical tensors and compared with the strict north-star gate
(tests/parity_util.compare_runs(strict=True): 0 image elements over 1e-4, integers equal, every gradient within
1e-3 of its tensor's maximum).  Usage: python tools/parity_fuzz.py [--cases 120] [--seed 1] [--out file]"""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import parity_util as pu
ge = pu.ge


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=120); ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--out", default=None)
    ap.add_argument("--opt", action="append", default=[], help="library option key=value (gsr_set_option)")
    a = ap.parse_args()
    for kv in a.opt:
        k, v = kv.split("=")
        assert pu.set_option(k, int(v)) >= 0, "unknown option " + k
    rng = np.random.RandomState(a.seed)
    sc = ge.load_scene_module()
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(pu.GOLDEN_DIR, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec); spec.loader.exec_module(mg)   # cov3d_of (3D covariance of a scene)
    mods = {v: (ge.load_variant(v), ge.load_reference(v)) for v in ("light", "full")}
    if mods["light"][1] is None:
        print("baseline/_ref not present: nothing to compare with"); return 0
    log = open(a.out, "w") if a.out else None

    def say(s):
        print(s, flush=True)
        if log:
            log.write(s + "\n"); log.flush()
    say("# tools/parity_fuzz.py --cases %d --seed %d %s  (%s, %s)" % (a.cases, a.seed, " ".join("--opt " + o for o in a.opt), torch.cuda.get_device_name(0), time.strftime("%Y-%m-%d %H:%M:%S")))
    worst = dict(pixels_over=0, int_mismatches=0, grad_max_rel=0.0, fwd_max_abs=0.0)
    failed = 0
    for case in range(a.cases):
        variant = "light" if rng.rand() < 0.5 else "full"
        W, H = int(rng.randint(17, 420)), int(rng.randint(17, 300))
        if rng.rand() < 0.3:
            W, H = 16 * (W // 16 + 1), 16 * (H // 16 + 1)
        P = int(rng.choice([1, 7, 60, 500, 3000, 12000, 40000, 150000]))
        lo = float(rng.choice([0.3, 1.0, 2.0])); hi = lo * float(rng.choice([2.0, 6.0, 15.0]))
        backdrop = bool(rng.rand() < 0.4) and P > 2 * max(2, W // 24) * max(2, H // 24)   # the backdrop grid is part of P
        deg = int(rng.randint(0, 4))
        use_sh = bool(rng.rand() < 0.8)
        cov_pre = bool(rng.rand() < 0.15)
        track_off = map_off = False
        if variant == "light":
            m = rng.randint(0, 3)
            track_off, map_off = (m == 1), (m == 2)
        cam = sc.make_camera(W, H, seed=int(rng.randint(0, 1000)))
        scene = sc.make_scene(P, cam, (lo, hi), seed=int(rng.randint(0, 10 ** 6)), backdrop=backdrop)
        cot = sc.make_cotangents(cam, 3 if variant == "light" else 2, seed=int(rng.randint(0, 10 ** 6)))
        kw = dict(use_sh=use_sh, sh_degree=deg, track_off=track_off, map_off=map_off)
        if cov_pre:
            kw["cov_precomp"] = mg.cov3d_of(scene)
        mine, ref = mods[variant]
        o_m, g_m = pu.run_variant(mine, variant, cam, scene, cot, **kw)
        o_r, g_r = pu.run_variant(ref, variant, cam, scene, cot, **kw)
        # -full's dL/dviewmatrix is only comparable on fully covered, 16-aligned images (upstream reads
        # uninitialised shared memory elsewhere)
        if variant == "full" and not (backdrop and W % 16 == 0 and H % 16 == 0):
            g_m.pop("viewmatrix", None); g_r.pop("viewmatrix", None)
        stats = {}
        ok, lines = pu.compare_runs(o_m, g_m, o_r, g_r, grad_budget=1e-3, strict=True, stats=stats)
        for k in worst:
            worst[k] = max(worst[k], stats.get(k, 0))
        tag = "ok  " if ok else "FAIL"
        failed += 0 if ok else 1
        say("%s case %3d: -%s P=%d %dx%d sigma=(%.1f,%.1f) deg=%d sh=%d covpre=%d backdrop=%d track_off=%d map_off=%d  "
            "pixels_over=%d int_mismatches=%d grad_max_rel=%.2e" % (
                tag, case, variant, P, W, H, lo, hi, deg, use_sh, cov_pre, backdrop, track_off, map_off,
                stats.get("pixels_over", -1), stats.get("int_mismatches", -1), stats.get("grad_max_rel", -1)))
        if not ok:
            for l in lines:
                if "FAIL" in l:
                    say("      " + l)
        torch.cuda.empty_cache()
    say("# %d cases, %d failed; worst: %s" % (a.cases, failed, worst))
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
