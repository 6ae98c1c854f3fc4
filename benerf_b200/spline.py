"""Mirror of the reference's spline.py entry points used on the hot path (spline.py:247-331).

Same signatures: knots are [1,1,6] se(3) vectors (rotation first), sample_time a 1-D tensor in
[0,1]; returns [P,3,4] camera-to-world poses.  One CUDA launch (csrc/pose.cu) instead of ~300.
Unlike the reference, sample_time is not modified in place (Q7's nudge happens in registers).
Both entry points are differentiable w.r.t. the knots (bnrf_spline_poses_backward), exactly like the
reference's torch ops: model/optimize.py:58-111 optimises the knots through them.
"""
import torch

from .engine import Engine, TRAJ

_engine = None


def _eng():
    global _engine
    if _engine is None:
        _engine = Engine()          # pose interpolation does not depend on the render configuration
    return _engine


class SplineFn(torch.autograd.Function):
    """poses [P,3,4] = interpolate(knots [4,6] (+ transform [1,6] added in se(3), model/optimize.py:86-89), ts [P])."""

    @staticmethod
    def forward(ctx, eng, traj, ts, knots, transform):
        k = knots.detach().to(eng.device, torch.float32).contiguous()
        t = transform.detach().reshape(6).to(eng.device, torch.float32).contiguous() if transform is not None else None
        ctx.eng, ctx.traj, ctx.ts, ctx.k, ctx.t = eng, traj, ts, k, t
        return eng.spline_poses(k, t, ts, traj)

    @staticmethod
    def backward(ctx, d_poses):
        d_knots, d_transform = ctx.eng.spline_poses_backward(ctx.k, ctx.t, ctx.ts, d_poses.contiguous(), ctx.traj)
        return None, None, None, d_knots, (d_transform.reshape(1, 6) if d_transform is not None else None)


def _poses(knots, sample_time, traj, engine):
    """knots: four [..., 6] tensors.  The stack keeps the autograd graph of the individual knots; the kernel's gradient
    w.r.t. the stacked [4,6] tensor flows back through it."""
    eng = engine or _eng()
    if traj not in TRAJ:
        raise ValueError(traj)
    stacked = torch.cat([k.reshape(1, 6) for k in knots], 0)
    ts = sample_time.detach().to(device=eng.device, dtype=torch.float32).contiguous()
    if torch.is_grad_enabled() and stacked.requires_grad:
        return SplineFn.apply(eng, traj, ts, stacked, None)
    return eng.spline_poses(stacked.detach().to(device=eng.device, dtype=torch.float32).contiguous(), None, ts, traj)


def cubic_spline_pose_unit_time(pose0, pose1, pose2, pose3, sample_time, engine=None):
    return _poses([pose0, pose1, pose2, pose3], sample_time, "spline", engine)


def linear_pose_unit_time(start_pose, end_pose, sample_time, engine=None):
    z = torch.zeros_like(start_pose)
    return _poses([start_pose, z, z, end_pose], sample_time, "linear", engine)     # the linear path reads knots 0 and 3 only
