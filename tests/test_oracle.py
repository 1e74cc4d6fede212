"""CPU tests of the oracle itself (oracle/gsr_oracle.c through oracle/oracle.py).

1. pinned against the golden vectors = outputs of the reference's CUDA build on a B200
   (tests/golden/case_*.npz, made by tests/golden/make_golden.py);
2. internal consistency: f32 vs f64 builds, closed-form single-splat cases, the documented
   reference quirks, analytic gradient vs finite differences of the f64
   oracle for the parts of the reference that ARE self-consistent."""
import numpy as np
import pytest
import torch

import parity_util as pu

ge = pu.ge


def _strip(outs):
    return {k: v for k, v in outs.items() if not k.startswith("_")}


# ---- 1. golden vectors -----------------------------------------------------------------------

CASES = pu.golden_cases()


@pytest.mark.skipif(not CASES, reason="no golden vectors committed yet")
@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    g = pu.load_golden(name)
    for (track_off, map_off) in g["modes"]:
        exp_o, exp_g = pu.golden_expected(g["data"], track_off, map_off)
        outs, grads = pu.run_oracle(g["variant"], g["cam"], g["scene"], g["cot"], use_sh=g["use_sh"],
                                    sh_degree=g["sh_degree"], track_off=track_off, map_off=map_off,
                                    cov_precomp=g["cov"])
        assert outs["_num_rendered"] == int(g["data"]["num_rendered"][0])
        if g["variant"] == "full":
            # identical hard decisions <=> identical number of valid (pixel, Gaussian) pairs
            assert abs(outs["_num_related"] - int(g["data"]["num_rendered"][1])) <= 4
        ok, lines = pu.compare_runs(_strip(outs), grads, exp_o, exp_g, flip_budget=5e-4,
                                    grad_budget=5e-3, label_a="oracle", label_b="reference")
        assert ok, "oracle vs reference golden (%s, track_off=%s map_off=%s):\n%s" % (
            name, track_off, map_off, "\n".join(lines))


@pytest.mark.skipif(not CASES, reason="no golden vectors committed yet")
@pytest.mark.parametrize("name", CASES)
def test_oracle_geometry_matches_reference_golden(name):
    g = pu.load_golden(name)
    orc = ge.load_oracle()
    cam, scene = g["cam"], g["scene"]
    r = orc.Run(g["variant"], cam.W, cam.H, cam.tanfovx, cam.tanfovy, scene.bg, scene.means3D,
                scene.opacities, cam.viewmatrix, cam.projmatrix, cam.campos, cam.perspec_matrix,
                scene.gt_depth, shs=scene.shs if g["use_sh"] else None,
                colors_precomp=None if g["use_sh"] else scene.colors,
                scales=None if g["cov"] is not None else scene.scales,
                rotations=None if g["cov"] is not None else scene.rotations,
                cov3D_precomp=g["cov"], sh_degree=g["sh_degree"])
    mine, ref = r.geometry(), pu.golden_geometry(g["data"])
    exp_o, _ = pu.golden_expected(g["data"], *g["modes"][0])
    vis = exp_o["radii"] > 0
    assert (r.radii == exp_o["radii"]).all()
    assert (mine["tiles_touched"][vis] == ref["tiles_touched"][vis]).all()
    if g["use_sh"]:  # with precomputed colours the reference never writes `clamped` (uninitialised)
        assert (mine["clamped"][vis].astype(bool) == ref["clamped"][vis].astype(bool)).all()
    for k, tol in (("depth", 1e-5), ("means2D", 1e-3), ("conic_opacity", 1e-4), ("rgb", 1e-5)):
        if k == "rgb" and not g["use_sh"]:
            continue  # precomputed colours are read in place: the reference leaves geom.rgb uninitialised
        a, b = np.asarray(mine[k])[vis], np.asarray(ref[k])[vis]
        scale = np.maximum(np.abs(b), 1.0)
        assert np.max(np.abs(a - b) / scale) < tol, k
    r.close()


# ---- 2. internal consistency -------------------------------------------------------------------

def _small(seed=3, P=400, W=64, H=48, backdrop=False):
    sc = ge.load_scene_module()
    cam = sc.make_camera(W, H)
    return sc, cam, sc.make_scene(P, cam, (2.0, 8.0), seed=seed, backdrop=backdrop)


@pytest.mark.parametrize("variant", ["light", "full"])
def test_f32_and_f64_builds_agree(variant):
    sc, cam, scene = _small()
    cot = sc.make_cotangents(cam, 3 if variant == "light" else 2)
    o32, g32 = pu.run_oracle(variant, cam, scene, cot, precision="f32")
    o64, g64 = pu.run_oracle(variant, cam, scene, cot, precision="f64")
    assert o32["_num_rendered"] == o64["_num_rendered"]
    ok, lines = pu.compare_runs(_strip(o32), g32, o64, g64, flip_budget=2e-3, grad_budget=2e-2)
    assert ok, "\n".join(lines)


def test_single_splat_closed_form():
    """One isotropic Gaussian in front of an identity camera: centre pixel alpha, colour, depth."""
    sc = ge.load_scene_module()
    W = H = 32
    cam = sc.make_camera(W, H, rot_deg=0.0, trans=(0.0, 0.0, 0.0))
    orc = ge.load_oracle()
    z, s, op = 2.0, 0.05, 0.8
    means = torch.tensor([[0.0, 0.0, z]])
    r = orc.Run("light", W, H, cam.tanfovx, cam.tanfovy, torch.tensor([0.1, 0.2, 0.3]), means,
                torch.tensor([[op]]), cam.viewmatrix, cam.projmatrix, cam.campos, cam.perspec_matrix,
                torch.full((1, H, W), 2.5), colors_precomp=torch.tensor([[0.9, 0.5, 0.25]]),
                scales=torch.tensor([[s, s, s]]), rotations=torch.tensor([[1.0, 0, 0, 0]]),
                precision="f64")
    fx = W / (2 * cam.tanfovx)
    var = (fx * s / z) ** 2 + 0.3  # EWA + 0.3 px^2 dilation (forward.cu:106-108)
    # splat centre is at pixel coordinate ((0+1)*W-1)/2 = 15.5; pixel 15 and 16 are 0.5 away
    d2 = 0.5 ** 2 + 0.5 ** 2
    alpha = op * np.exp(-0.5 * d2 / var)
    o = r.outputs()
    assert abs(o["opacity_map"][0, 15, 15] - alpha) < 1e-6
    assert abs(o["depth"][0, 16, 16] - alpha * z) < 1e-6
    assert abs(o["color"][0, 15, 16] - (0.9 * alpha + (1 - alpha) * 0.1)) < 1e-6
    assert o["depth_var"].max() == 0.0  # light never updates D_var (L/forward.cu:317,410)
    # isotropic splat: lambda = mid + sqrt(max(0.1, mid^2 - det)) = var + sqrt(0.1) (forward.cu:222-225)
    assert r.radii[0] == int(np.ceil(3 * np.sqrt(var + np.sqrt(0.1))))
    # median depth is set where T crosses 0.5: alpha > 0.5 at the centre
    assert o["depth_median"][0, 15, 15] == z and o["gau_related_pixels"][0, 0] > 0
    r.close()


def test_alpha_threshold_is_15_over_255():
    """Pairs with alpha < 15/255 are skipped (forward.cu:360; Inria uses 1/255)."""
    sc = ge.load_scene_module()
    W = H = 32
    cam = sc.make_camera(W, H, rot_deg=0.0, trans=(0.0, 0.0, 0.0))
    orc = ge.load_oracle()
    for op, expect_hit in ((0.058, False), (0.0595, True)):  # 15/255 = 0.05882
        r = orc.Run("full", W, H, cam.tanfovx, cam.tanfovy, torch.zeros(3), torch.tensor([[0.0, 0.0, 2.0]]),
                    torch.tensor([[op]]), cam.viewmatrix, cam.projmatrix, cam.campos,
                    cam.perspec_matrix, torch.ones(1, H, W), colors_precomp=torch.ones(1, 3),
                    scales=torch.full((1, 3), 0.5), rotations=torch.tensor([[1.0, 0, 0, 0]]),
                    precision="f64")
        assert (r.outputs()["uncertainty"].max() > 0) == expect_hit
        r.close()


def test_light_vs_full_termination_rule():
    """-full blends the Gaussian that drives T below 1e-4, -light does not."""
    sc = ge.load_scene_module()
    W = H = 16
    cam = sc.make_camera(W, H, rot_deg=0.0, trans=(0.0, 0.0, 0.0))
    orc = ge.load_oracle()
    n = 4  # alpha clamps at 0.99f: T = 0.0099999905 after one splat, 9.99998e-5 < 1e-4 after two
    means = torch.tensor([[0.0, 0.0, 1.0 + 0.5 * i] for i in range(n)])
    kw = dict(colors_precomp=torch.ones(n, 3), scales=torch.full((n, 3), 20.0),
              rotations=torch.tensor([[1.0, 0, 0, 0]]).repeat(n, 1), precision="f64")
    res = {}
    for v in ("light", "full"):
        r = orc.Run(v, W, H, cam.tanfovx, cam.tanfovy, torch.zeros(3), means, torch.full((n, 1), 0.995),
                    cam.viewmatrix, cam.projmatrix, cam.campos, cam.perspec_matrix, torch.ones(1, H, W), **kw)
        o = r.outputs()
        res[v] = o["opacity_map" if v == "light" else "uncertainty"][0, 8, 8]
        r.close()
    a = float(np.float32(0.99))
    assert abs(res["light"] - a) < 1e-6                  # second splat would end the pixel: rejected
    assert abs(res["full"] - (a + a * (1 - a))) < 1e-6   # second splat blended, then the pixel ends


def test_empty_and_all_culled_scene():
    sc, cam, scene = _small(P=50)
    orc = ge.load_oracle()
    behind = scene.means3D.clone()
    behind[:] = torch.tensor([0.0, 0.0, -5.0])
    r = orc.Run("light", cam.W, cam.H, cam.tanfovx, cam.tanfovy, scene.bg, behind, scene.opacities,
                cam.viewmatrix, cam.projmatrix, cam.campos, cam.perspec_matrix, scene.gt_depth,
                shs=scene.shs, scales=scene.scales, rotations=scene.rotations)
    o = r.outputs()
    assert r.num_rendered == 0 and (r.radii == 0).all()
    assert np.allclose(o["color"], scene.bg.numpy()[:, None, None])
    g = r.backward(torch.ones(3, cam.H, cam.W), torch.ones(1, cam.H, cam.W), torch.ones(1, cam.H, cam.W),
                   torch.ones(1, cam.H, cam.W))
    assert all(np.abs(v).max() == 0 for v in g.values())
    r.close()


def test_light_colour_gradients_match_finite_differences():
    """The -light colour/opacity path is a consistent forward/backward pair (unlike depth_var /
    -full uncertainty): check dL/dcolour and dL/dopacity of the f64 oracle against
    central differences of its own forward."""
    sc, cam, scene = _small(P=60, W=32, H=32, seed=11)
    orc = ge.load_oracle()
    ccol, caux = sc.make_cotangents(cam, 3)

    def fwd(colors, opac):
        r = orc.Run("light", cam.W, cam.H, cam.tanfovx, cam.tanfovy, scene.bg, scene.means3D, opac,
                    cam.viewmatrix, cam.projmatrix, cam.campos, cam.perspec_matrix, scene.gt_depth,
                    colors_precomp=colors, scales=scene.scales, rotations=scene.rotations,
                    precision="f64")
        return r

    r0 = fwd(scene.colors, scene.opacities)
    o = r0.outputs()
    zero = torch.zeros(1, cam.H, cam.W)
    g = r0.backward(ccol, zero, zero, zero, alphas=o["opacity_map"].astype(np.float32))
    vis = np.nonzero(r0.radii > 0)[0][:6]
    loss = lambda rr: float((rr.outputs()["color"] * ccol.numpy()).sum())
    eps = 1e-3
    for i in vis:
        for ch in range(3):
            cp, cm = scene.colors.clone(), scene.colors.clone()
            cp[i, ch] += eps
            cm[i, ch] -= eps
            fd = (loss(fwd(cp, scene.opacities)) - loss(fwd(cm, scene.opacities))) / (2 * eps)
            assert abs(fd - g["colors"][i, ch]) <= 1e-4 * max(1.0, abs(fd))
    r0.close()


# ---- 3. size-independent properties (the same ones the GPU suite checks on the CUDA path at full
#         size, here on the oracle at a size it finishes in a second) --------------------------------

def _small(seed=11, P=400, W=96, H=64):
    sc = ge.load_scene_module()
    cam = sc.make_camera(W, H)
    return sc, cam, sc.make_scene(P, cam, (2.0, 10.0), seed=seed)


@pytest.mark.parametrize("variant", ["light", "full"])
def test_backward_is_linear_in_the_cotangents(variant):
    sc, cam, scene = _small()
    n_aux = 3 if variant == "light" else 2
    c1 = sc.make_cotangents(cam, n_aux, seed=1)
    c2 = sc.make_cotangents(cam, n_aux, seed=2)
    if variant == "light":      # depth_var's backward term is quadratic in nothing but carries gt: keep it linear
        c1[1][2].zero_()
        c2[1][2].zero_()
    mix = (c1[0] + 2.0 * c2[0], [a + 2.0 * b for a, b in zip(c1[1], c2[1])])
    _, g1 = pu.run_oracle(variant, cam, scene, c1, precision="f64")
    _, g2 = pu.run_oracle(variant, cam, scene, c2, precision="f64")
    _, gm = pu.run_oracle(variant, cam, scene, mix, precision="f64")
    for k in g1:
        if g1[k] is None:
            continue
        want = g1[k].astype(np.float64) + 2.0 * g2[k].astype(np.float64)
        scale = max(1e-30, float(np.abs(want).max()))
        assert float(np.abs(gm[k] - want).max()) <= 1e-6 * scale, k   # results come back as float32


def test_background_enters_only_through_the_final_transmittance():
    sc, cam, scene = _small(seed=12)
    cot = sc.make_cotangents(cam, 3)
    o1, _ = pu.run_oracle("light", cam, scene, cot, backward=False, precision="f64")
    scene2 = scene._replace(bg=torch.tensor([0.9, 0.1, 0.5]))
    o2, _ = pu.run_oracle("light", cam, scene2, cot, backward=False, precision="f64")
    T = 1.0 - o1["opacity_map"].astype(np.float64)          # light: alpha map = sum alpha*T = 1 - T_final
    dbg = (scene2.bg - scene.bg).numpy().astype(np.float64)
    np.testing.assert_allclose(o2["color"] - o1["color"], dbg[:, None, None] * T, atol=1e-6)
    for k in ("depth", "depth_median", "opacity_map", "radii"):
        np.testing.assert_array_equal(o1[k], o2[k])


def test_gaussians_behind_the_camera_get_no_radius_and_no_gradient():
    sc, cam, scene = _small(seed=13, P=300)
    cot = sc.make_cotangents(cam, 2)
    outs, grads = pu.run_oracle("full", cam, scene, cot)
    w2c = cam.w2c.numpy().astype(np.float64)
    z = scene.means3D.numpy().astype(np.float64) @ w2c[2, :3] + w2c[2, 3]
    behind = z <= 0.2
    assert behind.any() and (~behind).any()
    assert (outs["radii"][behind] == 0).all()
    for k in ("means3D", "scales", "rotations", "opacities", "shs", "means2D"):
        assert float(np.abs(grads[k][behind]).max()) == 0.0, k
