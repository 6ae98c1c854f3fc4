"""Oracle: se(3) knots -> interpolated camera-to-world poses.

Test infrastructure (see oracle/__init__.py).  Restates, in functional form,
  spline.py:16-26   (se3_2_qt_parallel)        -> knot_to_quat_trans
  spline.py:28-34   (skew_symmetric)           -> _hat
  spline.py:46-62   (taylor_B / taylor_C)      -> _series
  spline.py:79-100  (exp_r2q_parallel)         -> rotvec_to_quat
  spline.py:167-192 (log_q2r_parallel)         -> quat_to_rotvec
  spline.py:130-148 (q_to_Q_parallel, conj)    -> quat_mul / quat_conj
  spline.py:111-118 (q_to_R_parallel)          -> quat_to_matrix
  spline.py:247-303 (cubic_spline_pose_unit_time)
  spline.py:305-331 (linear_pose_unit_time)
  model/optimize.py:58-111 (get_pose_evt / get_pose_rgb)
Quaternions are (x, y, z, w).  All arithmetic is fp32 and differentiable.
"""
import math
import torch


def _hat(w):
    """[..., 3] -> [..., 3, 3] cross-product matrix (spline.py:28-34)."""
    a, b, c = w.unbind(-1)
    z = torch.zeros_like(a)
    rows = [torch.stack([z, -c, b], -1), torch.stack([c, z, -a], -1), torch.stack([-b, a, z], -1)]
    return torch.stack(rows, -2)


def _series(x, first, nth=10):
    """sum_i (-1)^i x^(2i) / d_i with d_i = prod_{j<=i} (2j+first)(2j+first+1).

    first=1 is taylor_B ((1-cos x)/x^2, spline.py:46-53); first=2 is taylor_C
    ((x-sin x)/x^3, spline.py:55-62).  The running denominator is a Python
    float (double) exactly as in the reference.
    """
    total = torch.zeros_like(x)
    den = 1.0
    for i in range(nth + 1):
        den *= (2 * i + first) * (2 * i + first + 1)
        total = total + (-1) ** i * x ** (2 * i) / den
    return total


def rotvec_to_quat(r, eps=1e-9):
    """exp map so(3) -> unit quaternion, both branches evaluated (spline.py:79-100)."""
    x, y, z = r[..., 0], r[..., 1], r[..., 2]
    half = 0.5 * torch.sqrt(x ** 2 + y ** 2 + z ** 2)
    small = (half < eps).unsqueeze(-1).repeat(1, 1, 4)
    k = 1.0 / 2.0 - 1.0 / 12.0 * half ** 2 - 1.0 / 240.0 * half ** 4
    series = torch.stack([k * x, k * y, k * z, 1.0 - 1.0 / 2.0 * half ** 2 + 1.0 / 24.0 * half ** 4], -1)
    lam = torch.sin(half) / (2.0 * half)
    closed = torch.stack([lam * x, lam * y, lam * z, torch.cos(half)], -1)
    return torch.where(small, series, closed)


def quat_to_rotvec(q, eps_theta=1e-20, eps_w=1e-10):
    """log map, atan (not atan2) form with its three regimes (spline.py:167-192)."""
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    th = torch.sqrt(x ** 2 + y ** 2 + z ** 2)
    w_zero = torch.abs(w) < eps_w
    w_zero_neg = torch.logical_and(w_zero, w < 0)
    lam = torch.where(
        w_zero,
        torch.where(w_zero_neg, -torch.pi / th, torch.pi / th),
        torch.where(th < eps_theta,
                    2.0 / w - 2.0 / 3.0 * (th ** 2) / (w * w * w),
                    2.0 * (torch.arctan(th / w)) / th),
    )
    return torch.stack([lam * x, lam * y, lam * z], -1)


def quat_conj(q):
    return torch.stack([-q[..., 0], -q[..., 1], -q[..., 2], q[..., 3]], -1)


def _left_matrix(q):
    """Left-multiplication matrix Q(a): a (x) b = Q(a) b (spline.py:130-138)."""
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    rows = [torch.stack([w, -z, y, x], -1), torch.stack([z, w, -x, y], -1),
            torch.stack([-y, x, w, z], -1), torch.stack([-x, -y, -z, w], -1)]
    return torch.stack(rows, -2)


def quat_mul(a, b):
    return (_left_matrix(a) @ b[..., None]).squeeze(-1)


def quat_to_matrix(q):
    """(x,y,z,w) -> 3x3 rotation (spline.py:111-118)."""
    b, c, d, a = q.unbind(-1)
    r0 = torch.stack([1 - 2 * (c ** 2 + d ** 2), 2 * (b * c - a * d), 2 * (a * c + b * d)], -1)
    r1 = torch.stack([2 * (b * c + a * d), 1 - 2 * (b ** 2 + d ** 2), 2 * (c * d - a * b)], -1)
    r2 = torch.stack([2 * (b * d - a * c), 2 * (a * b + c * d), 1 - 2 * (b ** 2 + c ** 2)], -1)
    return torch.stack([r0, r1, r2], -2)


def knot_to_quat_trans(wu):
    """[1,1,6] (rotation first) -> quaternion [1,1,4], translation V(w) u [1,1,3] (spline.py:16-26)."""
    w, u = wu.split([3, 3], dim=-1)
    wx = _hat(w)
    th = w.norm(dim=-1)[..., None, None]
    eye = torch.eye(3, dtype=torch.float32)
    V = eye + _series(th, 1) * wx + _series(th, 2) * wx @ wx
    t = (V @ u[..., None]).squeeze(-1)
    return rotvec_to_quat(w), t


def _nudge(ts):
    """u==0 -> 1e-6, u==1 -> 1-1e-6 (spline.py:249-252; done out of place here)."""
    ts = ts.clone()
    ts[ts == 0] = ts[ts == 0] + 0.000001
    ts[ts == 1] = ts[ts == 1] - 0.000001
    return ts


def cubic_poses(k0, k1, k2, k3, ts):
    """Four knots [1,1,6] + P timestamps in [0,1] -> [P,3,4] (spline.py:247-303)."""
    u = _nudge(ts).unsqueeze(-1)
    q0, t0 = knot_to_quat_trans(k0)
    q1, t1 = knot_to_quat_trans(k1)
    q2, t2 = knot_to_quat_trans(k2)
    q3, t3 = knot_to_quat_trans(k3)
    uu, uuu = u ** 2, u ** 3
    s6, h = 1.0 / 6.0, 0.5
    c0 = s6 - h * u + h * uu - s6 * uuu
    c1 = 4 * s6 - uu + h * uuu
    c2 = s6 + h * u + h * uu - h * uuu
    c3 = s6 * uuu
    trans = c0 * t0 + c1 * t1 + c2 * t2 + c3 * t3
    b1 = 5 * s6 + h * u - h * uu + s6 * uuu
    b2 = s6 + h * u + h * uu - 2 * s6 * uuu
    b3 = s6 * uuu
    d01 = quat_to_rotvec(quat_mul(quat_conj(q0), q1)) * b1
    d12 = quat_to_rotvec(quat_mul(quat_conj(q1), q2)) * b2
    d23 = quat_to_rotvec(quat_mul(quat_conj(q2), q3)) * b3
    e0, e1, e2 = rotvec_to_quat(d01), rotvec_to_quat(d12), rotvec_to_quat(d23)
    q = quat_mul(q0, quat_mul(e0, quat_mul(e1, e2)))
    R = quat_to_matrix(q)
    return torch.cat([R, trans.unsqueeze(-1)], -1).reshape(-1, 3, 4)


def linear_poses(k_start, k_end, ts):
    """Two knots + P timestamps -> [P,3,4] (spline.py:305-331)."""
    ts = _nudge(ts)
    qs, t_s = knot_to_quat_trans(k_start)
    qe, t_e = knot_to_quat_trans(k_end)
    trans = (1 - ts)[..., None] * t_s + ts[..., None] * t_e
    r = ts[..., None] * quat_to_rotvec(quat_mul(quat_conj(qs), qe))
    q = quat_mul(qs, rotvec_to_quat(r))
    R = quat_to_matrix(q)
    return torch.cat([R, trans.unsqueeze(-1)], -1).reshape(-1, 3, 4)


def poses_from_knots(knots, transform, t_lo, t_hi, num, traj="spline"):
    """model/optimize.py:58-111.  knots [4,6]; transform [1,6] or None.

    The RGB camera's knots are the event knots plus ``transform`` *added in
    se(3)* (optimize.py:86-89); pass transform=None for the event camera.
    Timestamps are linspace(t_lo, t_hi, num).
    """
    ks = [knots[i].reshape(1, 1, 6) for i in range(4)]
    if transform is not None:
        ks = [k + transform.reshape(1, 1, 6) for k in ks]
    ts = torch.linspace(float(t_lo), float(t_hi), num)
    if traj == "linear":
        return linear_poses(ks[0], ks[3], ts)
    if traj == "spline":
        return cubic_poses(ks[0], ks[1], ks[2], ks[3], ts)
    raise ValueError(f"unknown traj {traj!r}")


PI = math.pi
