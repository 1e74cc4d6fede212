"""Find which Gaussian's tile count differs between processes (N +-1 flake)."""
import ctypes, hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import parity_util as pu
ge = pu.ge
sc = ge.load_scene_module()
cam, scene = sc.config("C3")
h = hashlib.sha1()
for t in list(scene) + [cam.viewmatrix, cam.projmatrix, cam.campos]:
    h.update(t.numpy().tobytes())
mod = ge.load_variant("full")
lib = ctypes.CDLL(ge.core_library_path())
dev = "cuda:0"
E = torch.Tensor([]); d = lambda t: t.to(dev)
P = scene.means3D.shape[0]
args = [d(scene.bg), d(scene.means3D), E, d(scene.opacities), d(scene.scales), d(scene.rotations), 1.0, E,
        d(cam.viewmatrix), d(scene.gt_depth), d(cam.projmatrix), cam.tanfovx, cam.tanfovy, cam.H, cam.W,
        d(scene.shs), 3, d(cam.campos), False]
r = mod._C.rasterize_gaussians(*args)
f = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)
tiles = torch.empty(P, dtype=torch.int32, device=dev)
m2d, co = f(P, 2), f(P, 4)
vp = lambda t: ctypes.c_void_p(t.data_ptr())
lib.gsr_decode_geometry(vp(r[6]), P, None, vp(m2d), vp(co), None, None, vp(tiles), None, None)
torch.cuda.synchronize()
t = tiles.cpu().numpy(); m = m2d.cpu().numpy(); c = co.cpu().numpy(); rad = r[5].cpu().numpy()
print("scene sha1", h.hexdigest()[:12], "N", int(r[0]), "tiles sha1", hashlib.sha1(t.tobytes()).hexdigest()[:12],
      "means2D sha1", hashlib.sha1(m[rad > 0].tobytes()).hexdigest()[:12], "conic sha1", hashlib.sha1(c[rad > 0].tobytes()).hexdigest()[:12])
ref_path = "gpurun_out/tiles_ref.npz"
if not os.path.exists(ref_path):
    np.savez(ref_path, t=t, m=m, c=c)
else:
    z = np.load(ref_path)
    idx = np.nonzero(z["t"] != t)[0]
    for i in idx[:5]:
        print("  gaussian", i, "tiles ref/now", z["t"][i], t[i], "means2D", z["m"][i], m[i], "conic_op", z["c"][i], c[i],
              "radius", rad[i], "opacity", float(scene.opacities[i]))
