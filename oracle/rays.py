"""Oracle: pixels + poses -> NDC rays, view directions, stratified depths.

Test infrastructure (see oracle/__init__.py).  Restates
  run_nerf_helpers.py:35-44  (get_specific_rays; get_rays :13-32 is its eval twin
                              and gives bit-identical rays)
  run_nerf_helpers.py:46-71  (ndc_rays, called with focal=K[0][0], near=1.0)
  model/nerf.py:241-254      (pose-major pixel x pose expansion)
  model/nerf.py:272-308      (viewdirs BEFORE ndc; near=0, far=1; jitter always on)
"""
import torch


def expand_pixels(poses, ray_idx, W, remap=None):
    """[P,3,4], [R] -> per-ray pixel coords (i, j) and c2w [P*R,3,4], pose-major.

    Ray n belongs to pose n // R and pixel n % R (model/nerf.py:242-245).  With
    a TUM-VIE undistortion LUT ``remap`` [H,W,2] the integer pixel is replaced by
    remap[j, i] (model/nerf.py:247-250).
    """
    P, R = poses.shape[0], ray_idx.shape[0]
    idx = ray_idx.repeat(P)
    c2w = poses.unsqueeze(1).repeat(1, R, 1, 1).reshape(-1, 3, 4)
    j = idx // W
    i = idx % W
    if remap is not None:
        rect = remap[j, i]
        i, j = rect[..., 0], rect[..., 1]
    return i, j, c2w


def camera_rays(i, j, K, c2w):
    """run_nerf_helpers.py:35-44.  K is a 3x3 tensor; returns (origins, directions)."""
    d_cam = torch.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -torch.ones_like(i)], -1)
    d_world = torch.sum(d_cam[..., None, :] * c2w[..., :3, :3], -1)
    return c2w[..., :3, -1], d_world


def to_ndc(H, W, focal, near, o, d):
    """run_nerf_helpers.py:46-71."""
    t = -(near + o[..., 2]) / d[..., 2]
    o = o + t[..., None] * d
    sx = -1.0 / (W / (2.0 * focal))
    sy = -1.0 / (H / (2.0 * focal))
    o0 = sx * o[..., 0] / o[..., 2]
    o1 = sy * o[..., 1] / o[..., 2]
    o2 = 1.0 + 2.0 * near / o[..., 2]
    d0 = sx * (d[..., 0] / d[..., 2] - o[..., 0] / o[..., 2])
    d1 = sy * (d[..., 1] / d[..., 2] - o[..., 1] / o[..., 2])
    d2 = -2.0 * near / o[..., 2]
    return torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)


def ray_batch(poses, ray_idx, H, W, K, remap=None, ndc=True):
    """model/nerf.py:241-279 -> (rays_o [N,3], rays_d [N,3], viewdirs [N,3]).

    viewdirs are normalised from the PRE-ndc direction (model/nerf.py:272-275).
    """
    i, j, c2w = expand_pixels(poses, ray_idx, W, remap)
    o, d = camera_rays(i, j, K, c2w)
    view = (d / torch.norm(d, dim=-1, keepdim=True)).reshape(-1, 3).float()
    if ndc:
        o, d = to_ndc(H, W, K[0][0], 1.0, o, d)
    return o.reshape(-1, 3).float(), d.reshape(-1, 3).float(), view


def stratified_depths(n_rays, n_samples, t_rand, near=0.0, far=1.0):
    """model/nerf.py:285-307.  t_rand [N,S] ~ U[0,1) is RNG draw #1 (always applied)."""
    near_t = near * torch.ones(n_rays, 1)
    far_t = far * torch.ones(n_rays, 1)
    t = torch.linspace(0.0, 1.0, steps=n_samples)
    z = (near_t * (1.0 - t) + far_t * t).expand([n_rays, n_samples])
    mids = 0.5 * (z[..., 1:] + z[..., :-1])
    upper = torch.cat([mids, z[..., -1:]], -1)
    lower = torch.cat([z[..., :1], mids], -1)
    return lower + (upper - lower) * t_rand


def sample_points(o, d, z):
    """model/nerf.py:308,327: pts = o + d * z  -> [N,S,3]."""
    return o[..., None, :] + d[..., None, :] * z[..., :, None]
