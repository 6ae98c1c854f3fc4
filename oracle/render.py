"""Oracle: Graph.render orchestration (coarse -> resample -> fine).

Test infrastructure (see oracle/__init__.py).  Restates model/nerf.py:236-343.
The four random draws of one render (order rand[N,S_c], randn[N,S_c],
rand[N,N_i], randn[N,S_c+N_i]) are explicit inputs in ``rng``:
  rng = {"t_rand": [N,S_c], "noise_c": [N,S_c], "u": [N,N_i], "noise_f": [N,S_c+N_i]}
"""
import torch

from . import rays as _rays
from .mlp import mlp_forward
from .composite import composite
from .resample import fine_depths


def draw_rng(n_rays, n_samples, n_importance, generator=None):
    """Generate the four draws in the reference's order and shapes."""
    g = generator
    rng = {"t_rand": torch.rand(n_rays, n_samples, generator=g),
           "noise_c": torch.randn(n_rays, n_samples, generator=g)}
    if n_importance > 0:
        rng["u"] = torch.rand(n_rays, n_importance, generator=g)
        rng["noise_f"] = torch.randn(n_rays, n_samples + n_importance, generator=g)
    return rng


def render(params_coarse, params_fine, poses, ray_idx, H, W, K, rng, *,
           n_samples=64, n_importance=64, channels=3, remap=None, ndc=True,
           near=0.0, far=1.0, return_intermediates=False, barf=None):
    """poses [P,3,4], ray_idx [R] -> dict like Graph.render (pose-major, N = P*R rays).

    Keys: rgb_map, disp_map, acc_map (+ rgb0, disp0, acc0, sigma if n_importance>0).
    The training and eval branches of the reference produce identical rays
    (SURVEY 8-a3), so one implementation serves both.
    """
    K = torch.as_tensor(K, dtype=torch.float32)
    o, d, view = _rays.ray_batch(poses, ray_idx, H, W, K, remap=remap, ndc=ndc)
    n = o.shape[0]
    z = _rays.stratified_depths(n, n_samples, rng["t_rand"], near, far)
    raw = mlp_forward(params_coarse, _rays.sample_points(o, d, z), view, barf)
    c = composite(raw, z, d, rng["noise_c"], channels)
    out = {"rgb_map": c["rgb_map"], "disp_map": c["disp_map"], "acc_map": c["acc_map"]}
    extra = {"rays_o": o, "rays_d": d, "viewdirs": view, "z_coarse": z, "raw_coarse": raw,
             "weights_coarse": c["weights"], "depth0": c["depth_map"], "sigma0": c["sigma"]}
    if n_importance > 0:
        zf = fine_depths(z, c["weights"], rng["u"])
        raw_f = mlp_forward(params_fine, _rays.sample_points(o, d, zf), view, barf)
        f = composite(raw_f, zf, d, rng["noise_f"], channels)
        out = {"rgb_map": f["rgb_map"], "disp_map": f["disp_map"], "acc_map": f["acc_map"],
               "rgb0": c["rgb_map"], "disp0": c["disp_map"], "acc0": c["acc_map"], "sigma": f["sigma"]}
        extra.update({"z_fine": zf, "raw_fine": raw_f, "weights_fine": f["weights"], "depth_map": f["depth_map"]})
    if return_intermediates:
        out["_extra"] = extra
    return out


def unstable_last_sample(raw, noise, eps=1e-4):
    """Rays whose last-sample density sits within eps of the relu kink.

    With dists[-1] = 1e10 the last alpha is 0 or ~1 depending on the SIGN of
    raw_sigma + noise (SURVEY 7 hard part 2 / Q13): a 1e-6 difference there moves
    rgb_map by O(1).  Parity tests report such rays separately instead of
    widening the tolerance.  raw [N,S,C+1], noise [N,S] -> bool mask [N].
    """
    return (raw[:, -1, -1] + noise[:, -1]).abs() < eps
