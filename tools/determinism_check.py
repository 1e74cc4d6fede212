"""Determinism probe: N (num_rendered) and output checksums over repeated forwards of config C3,
before and after autograd backward passes; scene hash to detect host-side input differences."""
import hashlib, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import parity_util as pu
ge = pu.ge
sc = ge.load_scene_module()
cam, scene = sc.config("C3")
h = hashlib.sha1()
for t in scene:
    h.update(t.numpy().tobytes())
for t in (cam.viewmatrix, cam.projmatrix, cam.campos):
    h.update(t.numpy().tobytes())
print("scene+camera sha1", h.hexdigest()[:16], "threads", torch.get_num_threads())
mod = ge.load_variant("full")
dev = "cuda:0"
E = torch.Tensor([])
d = lambda t: t.to(dev)
args = [d(scene.bg), d(scene.means3D), E, d(scene.opacities), d(scene.scales), d(scene.rotations), 1.0, E,
        d(cam.viewmatrix), d(scene.gt_depth), d(cam.projmatrix), cam.tanfovx, cam.tanfovy, cam.H, cam.W,
        d(scene.shs), 3, d(cam.campos), False]
def probe(tag, n=10):
    seen = {}
    for i in range(n):
        r = mod._C.rasterize_gaussians(*args)
        key = (int(r[0]), float(r[2].double().sum()), int(r[5].sum()))
        seen[key] = seen.get(key, 0) + 1
    print(tag, seen)
probe("before backward")
cot = sc.make_cotangents(cam, 2)
for i in range(4):
    pu.run_variant(mod, "full", cam, scene, cot)
probe("after 4 fwd+bwd")
inp = hashlib.sha1()
for t in args:
    if isinstance(t, torch.Tensor) and t.is_cuda:
        inp.update(t.cpu().numpy().tobytes())
print("device inputs sha1 after", inp.hexdigest()[:16])
