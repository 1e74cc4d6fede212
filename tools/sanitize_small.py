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
        for packed in (0, 1, 4):      # scalar, packed 8x8, octet; 2 (quarter lists) is the default below
            pu.set_option("bwd_packed", packed)
            pu.run_variant(mod, variant, cam, scene, cot)
        pu.set_option("bwd_packed", 2)
        for fp in (0, 1):             # scalar / packed forward blend
            pu.set_option("fwd_packed", fp)
            pu.run_variant(mod, variant, cam, scene, cot)
        pu.set_option("fwd_packed", 2)
        pu.set_option("tile_sort", 0)
        pu.run_variant(mod, variant, cam, scene, cot)
        pu.set_option("tile_sort", 1)
# speculative binning + blend that overflows and is redone (small splats, then large ones, same context)
cam = sc.make_camera(200, 120)
small, big = sc.make_scene(6000, cam, (0.3, 0.6), seed=46), sc.make_scene(6000, cam, (2.0, 14.0), seed=47)
for variant in ("light", "full"):
    mod = ge.load_variant(variant)
    cot = sc.make_cotangents(cam, 3 if variant == "light" else 2)
    for scn in (small, small, big, big):
        pu.run_variant(mod, variant, cam, scn, cot)
# exchange helpers on one GPU: factorized arena with the early masked colour kernel, P2P gather / slice all-reduce
# of the rank's own replica, SH rebuild
import torch
mod = ge.load_variant("full")
dp = ge.load_dp_module()
P = 4000
cam = sc.make_camera(200, 120)
scene = sc.make_scene(P, cam, (1.0, 10.0), seed=9)
cot = sc.make_cotangents(cam, 2)
arena = torch.zeros(14 * P + 4, device="cuda")
mod._C.set_grad_arena(arena, True, True)
pu.run_variant(mod, "full", cam, scene, cot)
assert mod._C.wait_masked_color(torch.cuda.current_stream().cuda_stream)
gathered = torch.empty(1, 3 * P + 4, device="cuda")
mod._C.p2p_gather([arena.data_ptr()], 3 * P + 4, gathered, 4)
mod._C.p2p_allreduce_slice([arena.data_ptr()], 3 * P + 4, 11 * P, 0, 4)
out = mod._C.sh_grad_from_views(scene.means3D.cuda(), gathered, 3, 16)
out2 = mod._C.sh_grad_from_view_ptrs(scene.means3D.cuda(), [arena.data_ptr()], [arena.data_ptr() + 4 * 3 * P], 3, 16)
torch.cuda.synchronize()
assert torch.equal(gathered[0], arena[:3 * P + 4]) and torch.allclose(out, out2, rtol=1e-5, atol=1e-8)
mod._C.set_grad_arena(torch.Tensor(), False, False)
# the RGB-D loss helper
for variant in ("light", "full"):
    mod = ge.load_variant(variant)
    import bench
    f = bench.Frame(mod, variant, cam, scene, sc.make_cotangents(cam, 3 if variant == "light" else 2), "cuda:0")
    f.zero_grad(); f.step_e2e(True); f.zero_grad(); f.step_e2e(True)
torch.cuda.synchronize()
print("rasterizer done")
# the device-side tracker (CUDA graph of 8 kernels + memset, fused loss, on-chip pose update)
import test_tracking_gpu as tt
s = tt._setup(P=3000, W=160, H=96)
t = tt._tracker(s, max_iterations=8)
r = t.run(4)
print("tracker done", r["loss"])
