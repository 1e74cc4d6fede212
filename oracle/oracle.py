"""ctypes driver of the CPU oracle (oracle/gsr_oracle.c).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg, never by the product path (diff-gaussian-rasterization_b200/ does not import it).

The oracle restates the reference's CUDA algorithm on the CPU (see the header of gsr_oracle.c for
the reference file:line of each stage).  Two builds of the same C file exist:
  liboracle_f32.so  REAL = float   (same rounding class as the reference; thresholds comparable)
  liboracle_f64.so  REAL = double  (adjudicates summation-order differences)
`run()` mirrors tests/parity_util.run_variant(): same inputs, same output / gradient dict keys.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_libs = {}


class OracleIn(C.Structure):
    _fields_ = [("variant", C.c_int), ("P", C.c_int), ("D", C.c_int), ("M", C.c_int),
                ("W", C.c_int), ("H", C.c_int),
                ("bg", C.c_void_p), ("means3D", C.c_void_p), ("shs", C.c_void_p),
                ("colors_precomp", C.c_void_p), ("opacities", C.c_void_p), ("scales", C.c_void_p),
                ("rotations", C.c_void_p), ("cov3D_precomp", C.c_void_p),
                ("scale_modifier", C.c_float),
                ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("campos", C.c_void_p),
                ("perspec", C.c_void_p),
                ("tanfovx", C.c_float), ("tanfovy", C.c_float),
                ("gt_depth", C.c_void_p)]


def lib(precision="f32"):
    assert precision in ("f32", "f64")
    if precision in _libs:
        return _libs[precision]
    path = os.path.join(HERE, "liboracle_%s.so" % precision)
    src = os.path.join(HERE, "gsr_oracle.c")
    if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", HERE], check=True)
    L = C.CDLL(path)
    L.oracle_forward.restype = C.c_void_p
    L.oracle_forward.argtypes = [C.POINTER(OracleIn)] + [C.c_void_p] * 10
    L.oracle_backward.restype = C.c_int
    L.oracle_backward.argtypes = [C.c_void_p] + [C.c_void_p] * 5 + [C.c_int, C.c_int] + [C.c_void_p] * 11
    L.oracle_geom.restype = None
    L.oracle_geom.argtypes = [C.c_void_p] + [C.c_void_p] * 7
    L.oracle_free.restype = None
    L.oracle_free.argtypes = [C.c_void_p]
    L.oracle_real_bytes.restype = C.c_int
    assert L.oracle_real_bytes() == (4 if precision == "f32" else 8)
    _libs[precision] = L
    return L


def _f32(t):
    """torch tensor / ndarray / None -> contiguous float32 ndarray (or None)."""
    if t is None:
        return None
    if hasattr(t, "detach"):
        t = t.detach().cpu().numpy()
    return np.ascontiguousarray(t, dtype=np.float32)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Run:
    """One oracle forward (kept alive for backward / geometry queries)."""

    def __init__(self, variant, W, H, tanfovx, tanfovy, bg, means3D, opacities, viewmatrix,
                 projmatrix, campos, perspec, gt_depth, shs=None, colors_precomp=None, scales=None,
                 rotations=None, cov3D_precomp=None, scale_modifier=1.0, sh_degree=3,
                 precision="f32"):
        assert variant in ("light", "full")
        self.L = lib(precision)
        self.variant = variant
        self.W, self.H = int(W), int(H)
        a = self.arrays = dict(
            bg=_f32(bg), means3D=_f32(means3D), shs=_f32(shs), colors_precomp=_f32(colors_precomp),
            opacities=_f32(opacities), scales=_f32(scales), rotations=_f32(rotations),
            cov3D_precomp=_f32(cov3D_precomp), viewmatrix=_f32(viewmatrix),
            projmatrix=_f32(projmatrix), campos=_f32(campos), perspec=_f32(perspec),
            gt_depth=_f32(gt_depth))
        P = self.P = a["means3D"].shape[0]
        self.M = 0 if a["shs"] is None else a["shs"].shape[1]
        inp = OracleIn()
        inp.variant = 0 if variant == "light" else 1
        inp.P, inp.D, inp.M, inp.W, inp.H = P, int(sh_degree), self.M, self.W, self.H
        for k in ("bg", "means3D", "shs", "colors_precomp", "opacities", "scales", "rotations",
                  "cov3D_precomp", "viewmatrix", "projmatrix", "campos", "perspec", "gt_depth"):
            setattr(inp, k, _ptr(a[k]))
        inp.scale_modifier = float(scale_modifier)
        inp.tanfovx, inp.tanfovy = float(tanfovx), float(tanfovy)
        self.inp = inp
        HW = self.W * self.H
        d = lambda *s: np.zeros(s, dtype=np.float64)
        self.color, self.depth, self.aux = d(3, self.H, self.W), d(1, self.H, self.W), d(1, self.H, self.W)
        self.median, self.var = d(1, self.H, self.W), d(1, self.H, self.W)
        self.radii = np.zeros(P, dtype=np.int32)
        self.gau_unc = d(P, 1)
        self.gau_px = np.zeros((P, 1), dtype=np.int32)
        nr, ng = C.c_int64(0), C.c_int64(0)
        self.ctx = self.L.oracle_forward(
            C.byref(inp), _ptr(self.color), _ptr(self.depth), _ptr(self.aux), _ptr(self.median),
            _ptr(self.var), _ptr(self.radii), _ptr(self.gau_unc), _ptr(self.gau_px),
            C.cast(C.byref(nr), C.c_void_p), C.cast(C.byref(ng), C.c_void_p))
        if not self.ctx:
            raise MemoryError("oracle_forward failed")
        self.num_rendered, self.num_related = int(nr.value), int(ng.value)
        assert HW == self.color.size // 3

    def outputs(self):
        if self.variant == "light":
            return dict(color=self.color, radii=self.radii, depth=self.depth,
                        depth_median=self.median, depth_var=self.var, opacity_map=self.aux,
                        gau_uncertainty=self.gau_unc, gau_related_pixels=self.gau_px)
        return dict(color=self.color, radii=self.radii, depth=self.depth, uncertainty=self.aux)

    def geometry(self):
        P = self.P
        g = dict(depth=np.zeros(P), means2D=np.zeros((P, 2)), conic_opacity=np.zeros((P, 4)),
                 rgb=np.zeros((P, 3)), cov3D=np.zeros((P, 6)),
                 tiles_touched=np.zeros(P, dtype=np.uint32), clamped=np.zeros((P, 3), dtype=np.uint8))
        self.L.oracle_geom(self.ctx, _ptr(g["depth"]), _ptr(g["means2D"]), _ptr(g["conic_opacity"]),
                           _ptr(g["rgb"]), _ptr(g["cov3D"]), _ptr(g["tiles_touched"]), _ptr(g["clamped"]))
        return g

    def backward(self, dL_dcolor, dL_ddepth, dL_dmedian=None, dL_dvar=None, alphas=None,
                 track_off=False, map_off=False):
        """dL_dvar = cotangent of depth_var (light) / uncertainty (full)."""
        P, M = self.P, self.M
        gc, gd, gm, gv, al = _f32(dL_dcolor), _f32(dL_ddepth), _f32(dL_dmedian), _f32(dL_dvar), _f32(alphas)
        d = lambda *s: np.zeros(s, dtype=np.float64)
        o = dict(means2D=d(P, 3), conic=d(P, 2, 2), opacities=d(P, 1), colors=d(P, 3), depths=d(P, 1),
                 means3D=d(P, 3), cov3D=d(P, 6), shs=d(P, max(M, 1), 3), scales=d(P, 3),
                 rotations=d(P, 4), viewmatrix=d(4, 4))
        rc = self.L.oracle_backward(
            self.ctx, _ptr(gc), _ptr(gd), _ptr(gm), _ptr(gv), _ptr(al), int(track_off), int(map_off),
            _ptr(o["means2D"]), _ptr(o["conic"]), _ptr(o["opacities"]), _ptr(o["colors"]),
            _ptr(o["depths"]), _ptr(o["means3D"]), _ptr(o["cov3D"]), _ptr(o["shs"]),
            _ptr(o["scales"]), _ptr(o["rotations"]), _ptr(o["viewmatrix"]))
        if rc != 0:
            raise RuntimeError("oracle_backward failed (%d)" % rc)
        if M == 0:
            o["shs"] = d(P, 0, 3)
        return o

    def close(self):
        if self.ctx:
            self.L.oracle_free(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def run(variant, cam, scene, cot, use_sh=True, sh_degree=3, track_off=False, map_off=False,
        backward=True, precision="f32", cov_precomp=None):
    """Same contract as tests/parity_util.run_variant: returns (outputs, grads) dicts of numpy."""
    r = Run(variant, cam.W, cam.H, cam.tanfovx, cam.tanfovy, scene.bg, scene.means3D,
            scene.opacities, cam.viewmatrix, cam.projmatrix, cam.campos, cam.perspec_matrix,
            scene.gt_depth, shs=scene.shs if use_sh else None,
            colors_precomp=None if use_sh else scene.colors,
            scales=None if cov_precomp is not None else scene.scales,
            rotations=None if cov_precomp is not None else scene.rotations,
            cov3D_precomp=cov_precomp, sh_degree=sh_degree, precision=precision)
    outs = r.outputs()
    grads = {}
    if backward:
        ccol, caux = cot
        if variant == "light":
            g = r.backward(ccol, caux[0], caux[1], caux[2],
                           alphas=outs["opacity_map"].astype(np.float32),
                           track_off=track_off, map_off=map_off)
        else:
            g = r.backward(ccol, caux[0], None, caux[1])
        grads = dict(means3D=g["means3D"], means2D=g["means2D"], opacities=g["opacities"],
                     viewmatrix=g["viewmatrix"])
        if cov_precomp is None:
            grads.update(scales=g["scales"], rotations=g["rotations"])
        else:
            grads["cov3D"] = g["cov3D"]
        if use_sh:
            grads["shs"] = g["shs"]
        else:
            grads["colors"] = g["colors"]
    outs = dict(outs)
    outs["_num_rendered"] = r.num_rendered
    outs["_num_related"] = r.num_related
    r.close()
    return outs, grads
