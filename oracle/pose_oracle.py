"""CPU restatement (numpy, float64) of the tracker's pose update — TEST INFRASTRUCTURE ONLY
(imported by tests/ alone; the product path is csrc/tracker.cu and never calls this).

What it restates: the chain rule from the reference's pose gradient dL/dviewmatrix
(L/cuda_rasterizer/backward.cu:701-751 + L/diff_gaussian_rasterization/__init__.py:160-176: 16 floats,
flat[4c + r] = dL/dW2C[r][c]) to the (quaternion, translation) parametrisation
W2C = [R(q/|q|) t; 0 1], the torch.optim.Adam update, and the camera tensors the rasterizer takes
(viewmatrix = W2C^T, projmatrix = viewmatrix @ perspec_matrix, campos = -R^T t).
The reference repository has no such code (it lives in CG-SLAM's tracker); this oracle is pinned
against torch autograd + torch.optim.Adam in tests/test_tracking_cpu.py.
"""
import numpy as np


def quat_to_R(q):
    q = np.asarray(q, dtype=np.float64)
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def camera_from_pose(q, t, perspec_matrix):
    """-> (viewmatrix [4,4] = W2C^T, projmatrix [4,4], campos [3]) as float64 arrays."""
    R = quat_to_R(q)
    w2c = np.eye(4)
    w2c[:3, :3] = R
    w2c[:3, 3] = np.asarray(t, dtype=np.float64)
    view = w2c.T
    return view, view @ np.asarray(perspec_matrix, dtype=np.float64), -(R.T @ w2c[:3, 3])


def pose_gradient(q, dL_dview):
    """dL/dq [4], dL/dt [3] from the 16-float dL/dviewmatrix (reference layout)."""
    q = np.asarray(q, dtype=np.float64)
    g = np.asarray(dL_dview, dtype=np.float64).reshape(4, 4)  # g[c][r] = dL/dW2C[r][c]
    dR = g[:3, :3].T
    dt = g[3, :3].copy()
    n = np.linalg.norm(q)
    w, x, y, z = q / n
    dRdw = 2 * np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]])
    dRdx = 2 * np.array([[0, y, z], [y, -2 * x, -w], [z, w, -2 * x]])
    dRdy = 2 * np.array([[-2 * y, x, w], [x, 0, z], [-w, z, -2 * y]])
    dRdz = 2 * np.array([[-2 * z, -w, x], [w, -2 * z, y], [x, y, 0]])
    gh = np.array([(dR * d).sum() for d in (dRdw, dRdx, dRdy, dRdz)])
    qh = q / n
    return (gh - qh * (qh @ gh)) / n, dt


def twist_gradient(q, t, dL_dview):
    """dL/d(omega, v) for the left perturbation W2C' = exp(xi^) W2C (SE(3) tangent space)."""
    g = np.asarray(dL_dview, dtype=np.float64).reshape(4, 4)
    dR, dt = g[:3, :3].T, g[3, :3]
    R = quat_to_R(q)
    om = sum(np.cross(R[:, c], dR[:, c]) for c in range(3)) + np.cross(np.asarray(t, dtype=np.float64), dt)
    return np.concatenate([om, dt])


class Adam:
    """torch.optim.Adam (no weight decay, no amsgrad) on a 7-vector with two learning rates."""

    def __init__(self, lr_rot, lr_trans, beta1=0.9, beta2=0.999, eps=1e-8):
        self.lr = np.array([lr_rot] * 4 + [lr_trans] * 3, dtype=np.float64)
        self.b1, self.b2, self.eps = beta1, beta2, eps
        self.m = np.zeros(7)
        self.v = np.zeros(7)
        self.k = 0

    def step(self, p, g):
        self.k += 1
        self.m = self.b1 * self.m + (1 - self.b1) * g
        self.v = self.b2 * self.v + (1 - self.b2) * g * g
        bc1, bc2 = 1 - self.b1 ** self.k, 1 - self.b2 ** self.k
        return p - (self.lr / bc1) * self.m / (np.sqrt(self.v) / np.sqrt(bc2) + self.eps)


def masked_l1(color, depth, alpha, gt_color, gt_depth, w_color, w_depth, alpha_thresh, use_depth_mask):
    """Loss value and cotangents (dL/dcolor [3,H,W], dL/ddepth [H,W]) of the tracker's fused loss."""
    color, depth, alpha = (np.asarray(a, dtype=np.float64) for a in (color, depth, alpha))
    gt_color, gt_depth = np.asarray(gt_color, dtype=np.float64), np.asarray(gt_depth, dtype=np.float64)
    m = alpha > alpha_thresh
    if use_depth_mask:
        m = m & (gt_depth > 0)
    m = m.astype(np.float64)
    ec, ed = color - gt_color, depth - gt_depth
    loss = w_color * (m * np.abs(ec)).sum() + w_depth * (m * np.abs(ed)).sum()
    return loss, w_color * m * np.sign(ec), w_depth * m * np.sign(ed)
