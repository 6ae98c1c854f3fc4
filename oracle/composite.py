"""Oracle: alpha compositing of raw MLP outputs along a ray.

Test infrastructure (see oracle/__init__.py).  Restates model/nerf.py:118-148
(NeRF.raw2output).  Quirks kept on purpose: sigma noise of std 1.0
is ALWAYS added (Q1), last interval is 1e10 (Q13), dists are scaled by the
POST-ndc direction norm (Q5), disparity is NaN when acc == 0 (Q14), the CRF
arguments of the reference are ignored there and do not exist here.
"""
import torch
import torch.nn.functional as F


def composite(raw, z, rays_d, noise, channels=3):
    """raw [N,S,C+1], z [N,S], rays_d [N,3], noise [N,S] ~ N(0,1) (RNG draw #2/#4).

    Returns dict(rgb_map [N,C], disp_map [N], acc_map [N], weights [N,S],
    depth_map [N], sigma [N,S]).
    """
    dists = z[..., 1:] - z[..., :-1]
    dists = torch.cat([dists, torch.tensor([1e10]).expand(dists[..., :1].shape)], -1)
    dists = dists * torch.norm(rays_d[..., None, :], dim=-1)
    rgb = torch.sigmoid(raw[..., :channels])
    dens = raw[..., channels] + noise
    alpha = 1.0 - torch.exp(-F.relu(dens) * dists)
    trans = torch.cumprod(torch.cat([torch.ones((alpha.shape[0], 1)), 1.0 - alpha + 1e-10], -1), -1)[:, :-1]
    weights = alpha * trans
    rgb_map = torch.sum(weights[..., None] * rgb, -2)
    depth_map = torch.sum(weights * z, -1)
    acc_map = torch.sum(weights, -1)
    disp_map = 1.0 / torch.max(1e-10 * torch.ones_like(depth_map), depth_map / torch.sum(weights, -1))
    return {"rgb_map": rgb_map, "disp_map": disp_map, "acc_map": acc_map,
            "weights": weights, "depth_map": depth_map, "sigma": F.relu(dens)}
