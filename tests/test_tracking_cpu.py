"""CPU tests of the tracker's host-side pieces: the pose oracle (oracle/pose_oracle.py) pinned
against torch autograd + torch.optim.Adam, the torch tracking-loop helpers, and the C-ABI exports."""
import ctypes
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


po = _load("pose_oracle", os.path.join(ROOT, "oracle", "pose_oracle.py"))
trk = _load("gsr_b200_tracking", os.path.join(ge.PKG, "tracking.py"))


def _random_pose(seed):
    rng = np.random.default_rng(seed)
    q = rng.normal(size=4)
    q = q / np.linalg.norm(q) * rng.uniform(0.7, 1.4)  # deliberately not unit length
    return q, rng.normal(size=3)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_pose_gradient_matches_autograd(seed):
    q0, t0 = _random_pose(seed)
    rng = np.random.default_rng(100 + seed)
    dview = rng.normal(size=(4, 4))
    dview[:, 3] = 0.0  # entries 3, 7, 11, 15 are always zero in the reference's dL_dview
    q = torch.tensor(q0, dtype=torch.float64, requires_grad=True)
    t = torch.tensor(t0, dtype=torch.float64, requires_grad=True)
    R = trk.quat_to_rotation(q)
    w2c = torch.cat([torch.cat([R, t[:, None]], dim=1),
                     torch.tensor([[0.0, 0.0, 0.0, 1.0]], dtype=torch.float64)], dim=0)
    (w2c.t() * torch.tensor(dview)).sum().backward()
    gq, gt = po.pose_gradient(q0, dview.reshape(-1))
    np.testing.assert_allclose(gq, q.grad.numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(gt, t.grad.numpy(), rtol=1e-10, atol=1e-12)


def _so3_exp(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * (K @ K)


@pytest.mark.parametrize("seed", [0, 1])
def test_twist_gradient_matches_finite_differences(seed):
    """dL/d(omega, v) of L = <dview, (exp(xi^) W2C)^T> at xi = 0, by central differences."""
    q0, t0 = _random_pose(seed)
    rng = np.random.default_rng(200 + seed)
    dview = rng.normal(size=(4, 4))
    dview[:, 3] = 0.0
    R = po.quat_to_R(q0)

    def L(xi):
        Re = _so3_exp(xi[:3])
        w2c = np.eye(4)
        w2c[:3, :3] = Re @ R
        w2c[:3, 3] = Re @ t0 + xi[3:]   # first order in xi (V(omega) v = v + O(|xi|^2))
        return float((w2c.T * dview).sum())
    fd = np.zeros(6)
    for k in range(6):
        e = np.zeros(6)
        e[k] = 1e-6
        fd[k] = (L(e) - L(-e)) / 2e-6
    np.testing.assert_allclose(po.twist_gradient(q0, t0, dview.reshape(-1)), fd, rtol=1e-6, atol=1e-8)


def test_adam_matches_torch():
    rng = np.random.default_rng(7)
    p0 = rng.normal(size=7)
    q = torch.tensor(p0[:4], dtype=torch.float64, requires_grad=True)
    t = torch.tensor(p0[4:], dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam([{"params": [q], "lr": 4e-4}, {"params": [t], "lr": 2e-3}],
                           betas=(0.9, 0.999), eps=1e-8)
    ad = po.Adam(4e-4, 2e-3)
    p = p0.copy()
    for k in range(25):
        g = rng.normal(size=7) * (10.0 ** rng.integers(-3, 3))
        q.grad = torch.tensor(g[:4])
        t.grad = torch.tensor(g[4:])
        opt.step()
        p = ad.step(p, g)
        np.testing.assert_allclose(p, np.concatenate([q.detach().numpy(), t.detach().numpy()]),
                                   rtol=1e-9, atol=1e-12)


def test_camera_from_pose_matches_scene_generator():
    sc = ge.load_scene_module()
    cam = sc.make_camera(64, 48)
    q = trk.rotation_to_quat(cam.w2c[:3, :3])
    view, proj, campos = po.camera_from_pose(q.numpy(), cam.w2c[:3, 3].numpy(), cam.perspec_matrix.numpy())
    np.testing.assert_allclose(view, cam.viewmatrix.numpy(), atol=2e-6)
    np.testing.assert_allclose(proj, cam.projmatrix.numpy(), atol=2e-5)
    np.testing.assert_allclose(campos, cam.campos.numpy(), atol=2e-6)
    np.testing.assert_allclose(po.quat_to_R(q.numpy()), trk.quat_to_rotation(q.double()).numpy(), atol=1e-7)


def test_masked_l1_matches_autograd():
    rng = np.random.default_rng(3)
    H, W = 12, 20
    color = torch.tensor(rng.uniform(0, 1, size=(3, H, W)), requires_grad=True)
    depth = torch.tensor(rng.uniform(0.5, 5, size=(H, W)), requires_grad=True)
    alpha = rng.uniform(0.9, 1.0, size=(H, W))
    gtc = rng.uniform(0, 1, size=(3, H, W))
    gtd = rng.uniform(0.5, 5, size=(H, W)) * (rng.uniform(size=(H, W)) > 0.2)
    mask = torch.tensor(((alpha > 0.95) & (gtd > 0)).astype(np.float64))
    loss = 0.5 * (mask * (color - torch.tensor(gtc)).abs()).sum() + 1.0 * (mask * (depth - torch.tensor(gtd)).abs()).sum()
    loss.backward()
    l, dc, dd = po.masked_l1(color.detach().numpy(), depth.detach().numpy(), alpha, gtc, gtd, 0.5, 1.0, 0.95, True)
    assert abs(l - float(loss)) < 1e-9
    np.testing.assert_allclose(dc, color.grad.numpy())
    np.testing.assert_allclose(dd, depth.grad.numpy())


@pytest.mark.parametrize("axis,deg", [((1, 0, 0), 10), ((0, 1, 0), 179), ((0, 0, 1), 179), ((1, 0, 0), 179),
                                      ((1, 2, 3), 120), ((-1, 1, 0.5), 65)])
def test_rotation_quaternion_round_trip(axis, deg):
    """rotation_to_quat covers all four branches (trace > 0 and the three dominant diagonals)."""
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    th = np.radians(deg)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)
    q = trk.rotation_to_quat(torch.tensor(R))
    assert abs(float(q.norm()) - 1.0) < 1e-6
    np.testing.assert_allclose(trk.quat_to_rotation(q.double()).numpy(), R, atol=1e-6)
    np.testing.assert_allclose(po.quat_to_R(q.numpy()), R, atol=1e-6)


def test_default_params_and_struct_layout():
    prm = trk.default_params(w_color=0.25, lr_rot=1e-3)
    assert prm["w_color"] == 0.25 and prm["lr_rot"] == 1e-3 and prm["beta2"] == 0.999 and prm["use_depth_mask"] is True
    cp = trk.TrackParams(1.0, 2.0, 0.5, 1, 1e-3, 2e-3, 0.9, 0.999, 1e-8)
    assert abs(cp.w_depth - 2.0) < 1e-7 and cp.use_depth_mask == 1 and abs(cp.eps - 1e-8) < 1e-12


def test_tracker_symbols_exported():
    lib = ctypes.CDLL(ge.core_library_path())
    for name in ("gsr_tracker_create", "gsr_tracker_destroy", "gsr_tracker_set_scene",
                 "gsr_tracker_set_frame", "gsr_tracker_set_pose", "gsr_tracker_run"):
        assert hasattr(lib, name), name
    assert ctypes.sizeof(trk.TrackParams) == 36
    assert ctypes.sizeof(trk.TrackResult) == 4 * (4 + 3 + 16 + 7 + 6 + 4)
