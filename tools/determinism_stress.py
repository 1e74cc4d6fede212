"""Stress probe for the N +-1 flake seen in bench.py: per forward compare the returned N with the sum
of tiles_touched decoded from that forward's own geometry buffer, across fresh small tensors, with
and without intervening backward passes and synchronisation."""
import ctypes, os, sys, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import parity_util as pu
ge = pu.ge
sc = ge.load_scene_module()
cam, scene = sc.config("C3")
mod = ge.load_variant("full")
lib = ctypes.CDLL(ge.core_library_path())
dev = "cuda:0"
E = torch.Tensor([])
d = lambda t: t.to(dev)
P = scene.means3D.shape[0]
fixed = dict(means=d(scene.means3D), op=d(scene.opacities), sc=d(scene.scales), rot=d(scene.rotations),
             shs=d(scene.shs), gt=d(scene.gt_depth))
cot = sc.make_cotangents(cam, 2)
tiles = torch.empty(P, dtype=torch.int32, device=dev)
random.seed(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
hist = {}
bad = 0
for it in range(120):
    mode = random.randrange(3)
    if mode == 1:
        pu.run_variant(mod, "full", cam, scene, cot)          # fwd+bwd, ends with a sync (.cpu())
    if mode == 2:
        torch.cuda.synchronize()
    args = [d(scene.bg), fixed["means"], E, fixed["op"], fixed["sc"], fixed["rot"], 1.0, E, d(cam.viewmatrix),
            fixed["gt"], d(cam.projmatrix), cam.tanfovx, cam.tanfovy, cam.H, cam.W, fixed["shs"], 3,
            d(cam.campos), False]
    r = mod._C.rasterize_gaussians(*args)
    n = int(r[0])
    lib.gsr_decode_geometry(ctypes.c_void_p(r[6].data_ptr()), P, None, None, None, None, None,
                            ctypes.c_void_p(tiles.data_ptr()), None, None)
    s = int(tiles.sum(dtype=torch.int64))
    hist[(n, s)] = hist.get((n, s), 0) + 1
    if n != s:
        bad += 1
print("(N returned, sum tiles_touched) histogram:", hist, "inconsistent:", bad)
