"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/gsr_b200.h declares; the torch shims import and expose the reference's three functions
(ext.cpp:15-19); argument validation of the C entry points works without a GPU."""
import ctypes
import inspect
import os
import re

import pytest


def declared_symbols():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "include", "gsr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"GSR_API\s+[\w\s\*]+?\b(gsr_\w+)\s*\(", src)))


def test_header_declares_the_five_operations():
    syms = declared_symbols()
    for s in ("gsr_light_forward", "gsr_light_backward", "gsr_full_forward", "gsr_full_backward",
              "gsr_mark_visible", "gsr_abi_version", "gsr_last_error", "gsr_backward_scratch_floats"):
        assert s in syms


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built.core_library_path())
    for s in declared_symbols():
        assert hasattr(lib, s), "libgsr_b200.so does not export %s" % s
    assert lib.gsr_abi_version() == 5


def test_scratch_size_and_options(built):
    lib = ctypes.CDLL(built.core_library_path())
    lib.gsr_backward_scratch_floats.restype = ctypes.c_size_t
    assert lib.gsr_backward_scratch_floats(0) >= 0
    n1, n2 = lib.gsr_backward_scratch_floats(1000), lib.gsr_backward_scratch_floats(2000)
    assert n2 > n1 >= 1000 * 16
    lib.gsr_last_error.restype = ctypes.c_char_p
    assert lib.gsr_get_option(b"exact_ng") == 0 and lib.gsr_get_option(b"tight_tiles") == 1
    assert lib.gsr_set_option(b"exact_ng", 1) == 0
    assert lib.gsr_get_option(b"exact_ng") == 1
    lib.gsr_set_option(b"exact_ng", 0)
    assert lib.gsr_set_option(b"no_such_option", 1) == -1
    assert b"unknown option" in lib.gsr_last_error()


def test_options_are_snapshotted_per_call_and_launch_counter_exists(built):
    """gsr_set_option changes the process-wide default only; gsr_launch_count counts on the host."""
    lib = ctypes.CDLL(built.core_library_path())
    lib.gsr_launch_count.restype = ctypes.c_longlong
    assert lib.gsr_launch_count(1) >= 0 and lib.gsr_launch_count(0) == 0
    old = lib.gsr_set_option(b"cnt_stride", 4)
    assert lib.gsr_get_option(b"cnt_stride") == 4
    lib.gsr_set_option(b"cnt_stride", old)
    assert lib.gsr_get_option(b"bwd_prefetch") == -1   # removed in round 2 (measured slower)


def test_misaligned_128bit_operands_are_rejected_without_a_gpu(built):
    """rotations / dL_dconic / dL_drot are accessed with 128-bit loads and stores: a pointer that is
    not 16-byte aligned is refused with GSR_E_INVALID instead of faulting on the device."""
    lib = ctypes.CDLL(built.core_library_path())
    lib.gsr_last_error.restype = ctypes.c_char_p
    ok, bad = ctypes.c_void_p(4096), ctypes.c_void_p(4096 + 4)
    cf, nr = ctypes.c_float, ctypes.c_int(0)
    alloc = ctypes.CFUNCTYPE(ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)(lambda c, n: None)
    rc = lib.gsr_light_forward(alloc, None, alloc, None, alloc, None, 4, 0, 0, ok, 16, 16, ok, None, ok, ok, ok,
                               cf(1.0), bad, None, ok, ok, ok, cf(1.0), cf(1.0), 0, ok, ok, ok, ok, ok, ok,
                               ok, ok, ok, 0, None, ctypes.byref(nr))
    assert rc == -1 and b"16-byte aligned" in lib.gsr_last_error()
    rc = lib.gsr_full_backward(4, 0, 0, 0, ok, 16, 16, ok, None, ok, ok, cf(1.0), ok, None, ok, ok, ok, cf(1.0),
                               cf(1.0), ok, ok, ok, ok, ok, ok, ok, ok, bad, ok, ok, ok, ok, ok, ok, ok, ok, ok, ok,
                               ok, ok, None, None)
    assert rc == -1 and b"16-byte aligned" in lib.gsr_last_error()


def test_invalid_arguments_are_rejected_without_a_gpu(built):
    lib = ctypes.CDLL(built.core_library_path())
    lib.gsr_last_error.restype = ctypes.c_char_p
    # negative P / null pointers: rejected before any CUDA call
    assert lib.gsr_mark_visible(-1, None, None, None, None, None) == -1
    assert lib.gsr_mark_visible(5, None, None, None, None, None) == -1
    assert lib.gsr_mark_visible(0, None, None, None, None, None) == 0
    assert lib.gsr_decode_geometry(None, 3, None, None, None, None, None, None, None, None) == -1


def test_extension_entry_points_validate_arguments_without_a_gpu(built):
    """Tracker, P2P SH rebuild and NVLS slice all-reduce reject bad arguments before any CUDA call."""
    lib = ctypes.CDLL(built.core_library_path())
    lib.gsr_last_error.restype = ctypes.c_char_p
    lib.gsr_tracker_create.restype = ctypes.c_void_p
    assert lib.gsr_tracker_create(0, 3, 16, 64, 48, ctypes.c_float(1.0), ctypes.c_float(1.0), None, 8) is None
    assert b"bad arguments" in lib.gsr_last_error()
    assert lib.gsr_tracker_run(None, None, 1, None, None, None) == -1
    assert lib.gsr_tracker_set_scene(None, None, None, None, None, None, ctypes.c_float(1.0), None, None, None) == -1
    assert lib.gsr_tracker_set_frame(None, None, None) == -1
    assert lib.gsr_tracker_set_pose(None, None, None) == -1
    lib.gsr_tracker_destroy(None)  # no-op
    # offset not a multiple of 4 floats / NULL multicast pointer / rank out of range
    f = ctypes.c_void_p(256)
    assert lib.gsr_nvls_allreduce_slice(None, ctypes.c_size_t(0), ctypes.c_size_t(16), 0, 2, 0, None) == -1
    assert lib.gsr_nvls_allreduce_slice(f, ctypes.c_size_t(2), ctypes.c_size_t(16), 0, 2, 0, None) == -1
    assert lib.gsr_nvls_allreduce_slice(f, ctypes.c_size_t(0), ctypes.c_size_t(16), 2, 2, 0, None) == -1
    assert lib.gsr_p2p_allreduce_slice(None, ctypes.c_size_t(0), ctypes.c_size_t(16), 0, 2, 0, None) == -1
    assert lib.gsr_p2p_gather(None, 2, ctypes.c_size_t(16), f, ctypes.c_size_t(16), 0, None) == -1
    assert lib.gsr_p2p_gather(f, 2, ctypes.c_size_t(18), f, ctypes.c_size_t(20), 0, None) == -1   # count not a multiple of 4
    assert lib.gsr_sh_grad_from_view_ptrs(8, 3, 16, None, 1, None, None, None, None) == -1
    assert lib.gsr_sh_grad_from_view_ptrs(8, 3, 16, f, 17, f, f, f, None) == -1   # more than 16 views
    assert lib.gsr_sh_grad_from_view_ptrs(0, 3, 16, None, 0, None, None, None, None) == 0
    # RGB-D L1 loss helper
    lib.gsr_rgbd_l1_scratch_floats.restype = ctypes.c_size_t
    assert lib.gsr_rgbd_l1_scratch_floats(1920, 1080) >= 1920 * 1080 // 256
    assert lib.gsr_rgbd_l1_loss(1, 16, 16, None, f, f, f, f, 1, f, 1, f, f, f, f, f, f, f, None) == -1
    assert lib.gsr_rgbd_l1_loss(2, 16, 16, f, f, f, f, f, 1, f, 1, f, f, f, f, f, f, f, None) == -1


@pytest.mark.parametrize("variant", ["light", "full"])
def test_shim_exports_reference_surface(built, variant):
    mod = built.load_variant(variant)
    for fn in ("rasterize_gaussians", "rasterize_gaussians_backward", "mark_visible"):
        assert callable(getattr(mod._C, fn))
    # GaussianRasterizationSettings field order (F/__init__.py:153-165, L/__init__.py:180-195)
    fields = list(mod.GaussianRasterizationSettings._fields)
    common = ["image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier",
              "viewmatrix", "projmatrix", "sh_degree", "campos", "prefiltered"]
    if variant == "light":
        assert fields == common + ["debug", "perspec_matrix", "track_off", "map_off"]
    else:
        assert fields == common + ["perspec_matrix"]
    sig = inspect.signature(mod.GaussianRasterizer.forward)
    assert list(sig.parameters) == ["self", "means3D", "means2D", "opacities", "shs", "colors_precomp",
                                    "scales", "rotations", "cov3D_precomp", "viewmatrix", "gt_depth"]


@pytest.mark.parametrize("variant", ["light", "full"])
def test_wrapper_argument_rules(built, variant):
    """'exactly one of' rules of GaussianRasterizer.forward (L/__init__.py:217-221)."""
    import torch
    mod = built.load_variant(variant)
    kw = dict(image_height=16, image_width=16, tanfovx=1.0, tanfovy=1.0, bg=torch.zeros(3),
              scale_modifier=1.0, viewmatrix=torch.eye(4), projmatrix=torch.eye(4), sh_degree=0,
              campos=torch.zeros(3), prefiltered=False, perspec_matrix=torch.eye(4))
    if variant == "light":
        kw.update(debug=False, track_off=False, map_off=False)
    rast = mod.GaussianRasterizer(mod.GaussianRasterizationSettings(**kw))
    m = torch.zeros(4, 3)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        rast(m, m, torch.ones(4, 1), shs=None, colors_precomp=None, scales=m, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        rast(m, m, torch.ones(4, 1), shs=torch.zeros(4, 1, 3), colors_precomp=m, scales=m,
             rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        rast(m, m, torch.ones(4, 1), colors_precomp=m)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        rast(m, m, torch.ones(4, 1), colors_precomp=m, scales=m, rotations=torch.zeros(4, 4),
             cov3D_precomp=torch.zeros(4, 6))
    # CPU tensors never reach a kernel: the shim refuses them loudly (no CPU fallback)
    with pytest.raises(Exception):
        rast(m, m, torch.ones(4, 1), colors_precomp=m, scales=m, rotations=torch.zeros(4, 4),
             viewmatrix=torch.eye(4), gt_depth=torch.zeros(1, 16, 16))
    with pytest.raises(Exception, match="num_points, 3"):
        mod._C.rasterize_gaussians(*_fwd_args(variant, torch.zeros(4, 2)))


def _fwd_args(variant, means):
    import torch
    E = torch.Tensor([])
    a = [torch.zeros(3), means, torch.zeros(4, 3), torch.ones(4, 1), torch.ones(4, 3), torch.zeros(4, 4),
         1.0, E, torch.eye(4), torch.zeros(1, 16, 16), torch.eye(4), 1.0, 1.0, 16, 16, E, 0,
         torch.zeros(3), False]
    if variant == "light":
        a.append(False)
    return a
