"""Same-process and cross-process determinism probe: N (num_rendered) and output checksums over
repeated forwards of config C3."""
import hashlib, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import __graft_entry__ as ge
sc = ge.load_scene_module()
cam, scene = sc.config("C3")
h = hashlib.sha1()
for t in scene:
    h.update(t.numpy().tobytes())
print("scene sha1", h.hexdigest()[:16])
mod = ge.load_variant("full")
dev = "cuda:0"
E = torch.Tensor([])
d = lambda t: t.to(dev)
args = [d(scene.bg), d(scene.means3D), E, d(scene.opacities), d(scene.scales), d(scene.rotations), 1.0, E,
        d(cam.viewmatrix), d(scene.gt_depth), d(cam.projmatrix), cam.tanfovx, cam.tanfovy, cam.H, cam.W,
        d(scene.shs), 3, d(cam.campos), False]
seen = {}
for i in range(40):
    r = mod._C.rasterize_gaussians(*args)
    key = (int(r[0]), float(r[2].double().sum()), int(r[5].sum()))
    seen[key] = seen.get(key, 0) + 1
print("distinct (N, sum(color), sum(radii)) over 40 forwards:", seen)
