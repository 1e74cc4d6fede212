#!/usr/bin/env python
"""In-tree build of the B200-native rasterizer.

  lib/libgsr_b200.so                                  CUDA core + C ABI (nvcc, sm_100a only)
  light/diff_gaussian_rasterization/_C.<abi>.so       torch shim, -DGSR_VARIANT_LIGHT
  full/diff_gaussian_rasterization/_C.<abi>.so        torch shim, -DGSR_VARIANT_FULL

The shims are plain C++ (no kernels) linked against libgsr_b200.so with an $ORIGIN rpath, so the
built tree is relocatable (it travels to the GPU box as is).  Nothing is installed into
site-packages.  Usage: python build.py [--force] [--core-only]
"""
import argparse
import concurrent.futures as cf
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BIND = os.path.join(HERE, "binding")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
CORE_SOURCES = ["api.cu", "preprocess_fwd.cu", "binning.cu", "render_fwd.cu", "render_bwd.cu", "sh_grad_views.cu",
                "preprocess_bwd.cu", "tracker.cu", "loss.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CXX = os.environ.get("CXX", "g++")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + "\n")
        raise RuntimeError("build step failed: " + cmd[-1])
    return r.stdout


def core_lib_path():
    return os.path.join(LIBDIR, "libgsr_b200.so")


def ext_suffix():
    return sysconfig.get_config_var("EXT_SUFFIX")


def module_path(variant):
    return os.path.join(HERE, variant, "diff_gaussian_rasterization", "_C" + ext_suffix())


def build_core(force=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    headers = [os.path.join(CSRC, "gsr_common.cuh"),
               os.path.join(HERE, "..", "include", "gsr_b200.h")]
    objs, jobs = [], []
    for src in CORE_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _newer(o, [s] + headers):
            jobs.append([NVCC] + NVCC_FLAGS + ["-c", s, "-o", o])
    with cf.ThreadPoolExecutor(max_workers=6) as ex:
        list(ex.map(_run, jobs))
    lib = core_lib_path()
    if force or jobs or _newer(lib, objs):
        _run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs + ["-ldl"])
    return lib


def build_shim(variant, force=False):
    import torch
    from torch.utils import cpp_extension as ce
    out = module_path(variant)
    srcs = [os.path.join(BIND, "ext.cpp"), os.path.join(BIND, "rasterize_points.cpp")]
    deps = srcs + [os.path.join(BIND, "rasterize_points.h"),
                   os.path.join(HERE, "..", "include", "gsr_b200.h")]
    if not (force or _newer(out, deps)):
        return out
    inc = []
    for p in ce.include_paths("cuda"):
        inc += ["-isystem", p]
    inc += ["-isystem", sysconfig.get_paths()["include"]]
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    defs = ["-DGSR_VARIANT_" + variant.upper(), "-DTORCH_EXTENSION_NAME=_C",
            "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    objs = []
    jobs = []
    for s in srcs:
        o = os.path.join(OBJDIR, "%s_%s.o" % (variant, os.path.basename(s).replace(".cpp", "")))
        objs.append(o)
        jobs.append([CXX, "-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden"] + defs + inc + ["-c", s, "-o", o])
    with cf.ThreadPoolExecutor(max_workers=2) as ex:
        list(ex.map(_run, jobs))
    _run([CXX, "-shared", "-o", out] + objs +
         ["-L" + LIBDIR, "-lgsr_b200", "-L" + torch_lib, "-ltorch", "-ltorch_cpu", "-ltorch_cuda",
          "-lc10", "-lc10_cuda", "-ltorch_python",
          # in-tree: the core sits in ../../lib; pip-installed (light|full/setup.py): next to the shim
          "-Wl,-rpath,$ORIGIN:$ORIGIN/../../lib", "-Wl,-rpath," + torch_lib, "-ldl"])
    return out


def build_all(force=False, core_only=False):
    os.makedirs(OBJDIR, exist_ok=True)
    lib = build_core(force)
    outs = [lib]
    if not core_only:
        with cf.ThreadPoolExecutor(max_workers=2) as ex:
            outs += list(ex.map(lambda v: build_shim(v, force), ["light", "full"]))
    return outs


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--core-only", action="store_true")
    a = ap.parse_args()
    for o in build_all(a.force, a.core_only):
        print("built", os.path.relpath(o, HERE))
