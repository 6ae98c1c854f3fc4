"""Seeded synthetic inputs shared by the golden generator and the parity tests.

Every input (weights, knots, pixels, the four RNG draws of a render, events,
blurry pixels) comes from ``numpy.random.default_rng(seed)`` so that the SAME
inputs can be rebuilt on the GPU box, where /root/reference does not exist, and
compared with the reference outputs stored under tests/golden/ (written by
tools/make_golden.py in the build container).

Camera intrinsics are the ones the reference's configs ship.
"""
import math
from dataclasses import dataclass, field

import numpy as np
import torch


@dataclass(frozen=True)
class Case:
    name: str
    dataset: str
    channels: int
    H: int
    W: int
    fx: float
    cx: float
    cy: float
    n_poses: int            # virtual poses across the exposure (blur model)
    r_rgb: int              # blur pixels (the reference uses sampling_rgb_rays // n_poses)
    r_evt: int              # event pixels (rendered at 2 poses)
    n_samples: int = 64
    n_importance: int = 64
    event_threshold: float = 0.1
    traj: str = "spline"
    knot_scale: float = 0.01
    transform_scale: float = 0.0
    bias_scale: float = 0.05
    exposure: tuple = (0.2, 0.8)
    window: tuple = (0.3, 0.4)
    training: bool = True
    seed: int = 0
    n_events: int = 4000
    barf_iter: int = -1     # >= 0: args.use_barf_c2f with iter_step = barf_iter of max_iter = 1000, start 0.1, end 0.5 (model/nerf.py:16-26)

    @property
    def K(self):
        return np.array([[self.fx, 0.0, self.cx], [0.0, self.fx, self.cy], [0.0, 0.0, 1.0]], dtype=np.float64)


# BASELINE.json configs at test scale (few pixels; full pose counts and sample counts).
CASES = {c.name: c for c in [
    # config 2: benerf_unreal RGB + events, 19 poses, 64+128
    Case("unreal_rgb", "BeNeRF_Unreal", 3, 480, 768, 548.409, 384.0, 240.0, 19, 6, 24, seed=11),
    # config 1: benerf_blender gray 200x200, 7 poses, coarse only, eval branch of Graph.render
    Case("blender_gray_coarse", "BeNeRF_Blender", 1, 200, 200, 541.850232 * 200 / 600, 100.0, 100.0,
         7, 16, 16, n_importance=0, training=False, seed=12),
    # config 3: e2nerf_synthetic, lin-log brightness, thresholded event loss
    Case("e2nerf_syn", "E2NeRF_Synthetic", 3, 800, 800, 1111.111, 400.0, 400.0, 19, 4, 20,
         event_threshold=0.2, transform_scale=0.005, seed=13),
    # config 4: e2nerf_real, 31 poses, normalised event loss
    Case("e2nerf_real", "E2NeRF_Real", 3, 260, 346, 653.98456, 173.0, 130.0, 31, 3, 24,
         event_threshold=-1.0, seed=14),
    # gray + fine network, linear trajectory, larger camera motion, zero biases (init_nerf state)
    Case("gray_linear", "BeNeRF_Blender", 1, 400, 600, 541.850232, 300.0, 200.0, 5, 8, 8,
         traj="linear", knot_scale=0.2, transform_scale=0.05, bias_scale=0.0, seed=15),
    # BARF coarse-to-fine weighting of the positional encodings, mid-schedule (frequencies 0..3 on, 4 partly, 5..9 off)
    Case("barf_c2f", "BeNeRF_Unreal", 3, 480, 768, 548.409, 384.0, 240.0, 5, 6, 12, seed=16, barf_iter=270),
]}

BARF_MAX_ITER, BARF_START, BARF_END = 1000, 0.1, 0.5


def barf_of(case):
    """(progress, start, end) of a BARF case for the oracle, else None."""
    return (case.barf_iter / BARF_MAX_ITER, BARF_START, BARF_END) if case.barf_iter >= 0 else None


def layer_shapes(channels):
    s = {"pts_linears.0": (256, 63)}
    for i in range(1, 8):
        s[f"pts_linears.{i}"] = (256, 319 if i == 5 else 256)
    s["views_linears.0"] = (128, 283)
    s["feature_linear"] = (256, 256)
    s["alpha_linear"] = (1, 256)
    s["rgb_linear"] = (channels, 128)
    return s


def make_params(rng, channels, bias_scale):
    """Xavier-uniform weights (run_nerf_helpers.py:194-197 bound) + optional small biases."""
    p = {}
    for name, (o, i) in layer_shapes(channels).items():
        bound = math.sqrt(6.0 / (o + i))
        p[name + ".weight"] = torch.from_numpy(rng.uniform(-bound, bound, (o, i)).astype(np.float32))
        p[name + ".bias"] = torch.from_numpy(rng.uniform(-bias_scale, bias_scale, (o,)).astype(np.float32))
    return p


def make_rng_draws(rng, n_rays, n_samples, n_importance):
    d = {"t_rand": torch.from_numpy(rng.random((n_rays, n_samples), dtype=np.float32)),
         "noise_c": torch.from_numpy(rng.standard_normal((n_rays, n_samples), dtype=np.float32))}
    if n_importance > 0:
        d["u"] = torch.from_numpy(rng.random((n_rays, n_importance), dtype=np.float32))
        d["noise_f"] = torch.from_numpy(rng.standard_normal((n_rays, n_samples + n_importance), dtype=np.float32))
    return d


def make_inputs(case: Case):
    """All inputs of one training-shaped iteration for ``case`` (deterministic)."""
    rng = np.random.default_rng(case.seed)
    inp = {
        "coarse": make_params(rng, case.channels, case.bias_scale),
        "fine": make_params(rng, case.channels, case.bias_scale) if case.n_importance > 0 else None,
        "knots": torch.from_numpy((rng.random((4, 6)) * case.knot_scale).astype(np.float32)),
        "transform": torch.from_numpy(((rng.random((1, 6)) - 0.5) * 2 * case.transform_scale).astype(np.float32)),
        "idx_rgb": torch.from_numpy(rng.permutation(case.H * case.W)[:case.r_rgb].astype(np.int64)),
        "idx_evt": torch.from_numpy(rng.permutation(case.H * case.W)[:case.r_evt].astype(np.int64)),
    }
    inp["rng_evt"] = make_rng_draws(rng, 2 * case.r_evt, case.n_samples, case.n_importance)
    inp["rng_rgb"] = make_rng_draws(rng, case.n_poses * case.r_rgb, case.n_samples, case.n_importance)
    inp["blur_target"] = torch.from_numpy(rng.random((case.r_rgb, case.channels), dtype=np.float32))
    ev = {"x": rng.integers(0, case.W, case.n_events), "y": rng.integers(0, case.H, case.n_events),
          "ts": np.sort(rng.random(case.n_events)), "pol": rng.integers(0, 2, case.n_events) * 2.0 - 1.0}
    # make sure the sampled event pixels see some events
    flat = inp["idx_evt"].numpy()
    ev["x"][:flat.size * 4] = np.tile(flat % case.W, 4)
    ev["y"][:flat.size * 4] = np.tile(flat // case.W, 4)
    inp["events"] = ev
    return inp


def golden_path(case_name):
    import os
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", case_name + ".npz")


def load_golden(case_name):
    with np.load(golden_path(case_name)) as f:
        return {k: torch.from_numpy(f[k]) for k in f.files}
