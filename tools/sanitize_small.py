"""Small fwd+bwd of both variants for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity_util as pu
ge = pu.ge
for kv in os.environ.get("GSR_TEST_OPTS", "").split(","):
    if kv:
        k, v = kv.split("=")
        pu.set_option(k, int(v))
SIZES = ((3000, 200, 120), (20000, 330, 250)) if not os.environ.get("GSR_SANITIZE_SMALL") else ((1500, 120, 72),)
sc = ge.load_scene_module()
for variant in ("light", "full"):
    for (P, W, H) in SIZES:
        cam = sc.make_camera(W, H)
        scene = sc.make_scene(P, cam, (1.0, 10.0), seed=7)
        cot = sc.make_cotangents(cam, 3 if variant == "light" else 2)
        mod = ge.load_variant(variant)
        for packed in (0, 1):
            pu.set_option("bwd_packed", packed)
            pu.run_variant(mod, variant, cam, scene, cot)
        pu.set_option("bwd_packed", 2)
        pu.set_option("tile_sort", 0)
        pu.run_variant(mod, variant, cam, scene, cot)
        pu.set_option("tile_sort", 1)
print("rasterizer done")
# the device-side tracker (CUDA graph of 8 kernels + memset, fused loss, on-chip pose update)
import test_tracking_gpu as tt
s = tt._setup(P=3000, W=160, H=96)
t = tt._tracker(s, max_iterations=8)
r = t.run(4)
print("tracker done", r["loss"])
