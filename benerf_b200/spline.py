"""Mirror of the reference's spline.py entry points used on the hot path (spline.py:247-331).

Same signatures: knots are [1,1,6] se(3) vectors (rotation first), sample_time a 1-D tensor in
[0,1]; returns [P,3,4] camera-to-world poses.  One CUDA launch (csrc/pose.cu) instead of ~300.
Unlike the reference, sample_time is not modified in place (Q7's nudge happens in registers).
"""
import torch

from .engine import Engine

_engine = None


def _eng():
    global _engine
    if _engine is None:
        _engine = Engine()          # pose interpolation does not depend on the render configuration
    return _engine


def _stack(knots, device):
    return torch.cat([k.reshape(1, 6) for k in knots], 0).to(device=device, dtype=torch.float32).contiguous()


def cubic_spline_pose_unit_time(pose0, pose1, pose2, pose3, sample_time, engine=None):
    eng = engine or _eng()
    knots = _stack([pose0, pose1, pose2, pose3], eng.device)
    ts = sample_time.to(device=eng.device, dtype=torch.float32).contiguous()
    return eng.spline_poses(knots, None, ts, "spline")


def linear_pose_unit_time(start_pose, end_pose, sample_time, engine=None):
    eng = engine or _eng()
    z = torch.zeros_like(start_pose)
    knots = _stack([start_pose, z, z, end_pose], eng.device)     # the linear path reads knots 0 and 3 only
    ts = sample_time.to(device=eng.device, dtype=torch.float32).contiguous()
    return eng.spline_poses(knots, None, ts, "linear")
