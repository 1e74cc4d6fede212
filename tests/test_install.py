"""The install contract of the reference (README.md:26,63; diff-gaussian-rasterization-light/setup.py:15-36):
`pip install .` in either package directory builds the extension and `import diff_gaussian_rasterization`
works from site-packages.  Here: pip-install each variant into a temporary target directory (no index,
no build isolation — torch and nvcc come from the environment), import it from there in a fresh
interpreter and check the reference's surface; on a GPU box additionally run one forward+backward from
the INSTALLED copy and compare it bit for bit with the in-tree package."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "diff-gaussian-rasterization_b200")


@pytest.fixture(scope="module", params=["light", "full"])
def installed(request, built, tmp_path_factory):
    variant = request.param
    target = str(tmp_path_factory.mktemp("site_" + variant))
    r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                        "--quiet", "--target", target, os.path.join(PKG, variant)],
                       capture_output=True, text=True, cwd="/tmp")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return variant, target


def _run_py(target, code):
    env = dict(os.environ, PYTHONPATH=target)
    return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp", env=env)


def test_pip_installed_package_imports_from_site_packages(installed):
    variant, target = installed
    files = os.listdir(os.path.join(target, "diff_gaussian_rasterization"))
    assert "libgsr_b200.so" in files and any(f.startswith("_C.") and f.endswith(".so") for f in files)
    r = _run_py(target, """
import diff_gaussian_rasterization as m
assert m.__file__.startswith(%r), m.__file__
for fn in ("rasterize_gaussians", "rasterize_gaussians_backward", "mark_visible"):
    assert callable(getattr(m._C, fn))
assert callable(m.rasterize_gaussians)
print(len(m.GaussianRasterizationSettings._fields), m.GaussianRasterizer.__name__)
""" % target)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.split() == ["15" if variant == "light" else "12", "GaussianRasterizer"]


@pytest.mark.gpu
def test_pip_installed_package_renders_like_the_in_tree_one(installed):
    variant, target = installed
    r = _run_py(target, """
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import diff_gaussian_rasterization as inst
assert inst.__file__.startswith(%r)
import parity_util as pu
sc = pu.ge.load_scene_module()
cam = sc.make_camera(160, 96)
scene = sc.make_scene(3000, cam, (1.0, 10.0), seed=11, backdrop=True)
cot = sc.make_cotangents(cam, 3 if %r == "light" else 2)
o_i, g_i = pu.run_variant(inst, %r, cam, scene, cot)
o_t, g_t = pu.run_variant(pu.ge.load_variant(%r), %r, cam, scene, cot)
for k in o_t:
    if k != "gau_uncertainty":
        assert np.array_equal(o_i[k], o_t[k]), k
for k in g_t:
    rel, _ = pu.grad_mismatch(g_i[k], g_t[k], rtol=1e-4)
    assert rel < 1e-4, (k, rel)
print("installed == in-tree")
""" % (ROOT, os.path.join(ROOT, "tests"), target, variant, variant, variant, variant))
    assert r.returncode == 0 and "installed == in-tree" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
