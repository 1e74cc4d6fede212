"""pip-installable `diff_gaussian_rasterization` (-full surface) backed by the B200-native core.

    pip install --no-build-isolation .        (from this directory; needs torch + nvcc, sm_100a only)

Same install contract as the reference's diff-gaussian-rasterization-full/setup.py:15-36 (package name
`diff_gaussian_rasterization`, extension module `diff_gaussian_rasterization._C`), so CG-SLAM's
"pip install submodules/diff-gaussian-rasterization-full" step works unchanged when the submodule points
here.  The build is ../build.py's: libgsr_b200.so (all CUDA, nvcc -gencode arch=compute_100a,code=sm_100a)
plus the thin torch shim `_C`; both are installed INSIDE the package directory (the shim finds the
core through an $ORIGIN rpath), nothing else is written to site-packages.
"""
import importlib.util
import os
import shutil

from setuptools import Extension, setup
from setuptools.command.build_ext import build_ext

HERE = os.path.dirname(os.path.abspath(__file__))
VARIANT = "full"


def _native_build():
    spec = importlib.util.spec_from_file_location("gsr_b200_build", os.path.join(HERE, "..", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class BuildNative(build_ext):
    """Builds the CUDA core and the torch shim with ../build.py and places both in the package."""

    def run(self):
        b = _native_build()
        outs = [b.build_core(), b.build_shim(VARIANT)]
        dst = os.path.join(self.build_lib, "diff_gaussian_rasterization")
        os.makedirs(dst, exist_ok=True)
        for o in outs:
            shutil.copy2(o, dst)
        if self.inplace:   # pip install -e / build_ext --inplace: the shim is already in place, add the core
            shutil.copy2(outs[0], os.path.join(HERE, "diff_gaussian_rasterization"))

    def get_outputs(self):
        b = _native_build()
        dst = os.path.join(self.build_lib, "diff_gaussian_rasterization")
        return [os.path.join(dst, os.path.basename(b.core_lib_path())),
                os.path.join(dst, os.path.basename(b.module_path(VARIANT)))]


setup(
    name="diff_gaussian_rasterization",
    version="0.2.0+b200.full",
    description="B200-native (sm_100a) drop-in for hjr37/diff-gaussian-rasterization-full",
    packages=["diff_gaussian_rasterization"],
    ext_modules=[Extension("diff_gaussian_rasterization._C", sources=[])],
    cmdclass={"build_ext": BuildNative},
)
