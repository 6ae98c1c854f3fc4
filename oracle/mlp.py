"""Oracle: the 8x256 NeRF MLP with view branch.

Test infrastructure (see oracle/__init__.py).  Restates model/nerf.py:40-116
(NeRF.__init__ / NeRF.forward with use_viewdirs=True, skips=[4], BARF off).
Parameters use the reference's state-dict names and (out, in) layout:
  pts_linears.{0..7}  (256,63) (256,256)x4 (256,319) (256,256)x2
  views_linears.0     (128,283)     feature_linear (256,256)
  alpha_linear        (1,256)       rgb_linear     (C,128)
"""
import math
import torch
import torch.nn.functional as F

from .encode import positional_encoding, encode

D, WIDTH, SKIP = 8, 256, 4
PTS_FREQS, DIR_FREQS = 10, 4
PTS_CH, DIR_CH = 63, 27


def layer_shapes(channels=3):
    """Ordered {name: (out, in)} of the 12 linears (model/nerf.py:53-64)."""
    shapes = {"pts_linears.0": (WIDTH, PTS_CH)}
    for i in range(1, D):
        shapes[f"pts_linears.{i}"] = (WIDTH, WIDTH + PTS_CH if i == SKIP + 1 else WIDTH)
    shapes["views_linears.0"] = (WIDTH // 2, DIR_CH + WIDTH)
    shapes["feature_linear"] = (WIDTH, WIDTH)
    shapes["alpha_linear"] = (1, WIDTH)
    shapes["rgb_linear"] = (channels, WIDTH // 2)
    return shapes


def xavier_params(channels=3, generator=None, bias_scale=0.0):
    """Xavier-uniform weights, zero biases: run_nerf_helpers.py:194-208 (init_nerf).

    ``bias_scale`` > 0 draws small non-zero biases instead so tests exercise the
    bias path (the reference zeroes them at iteration 0 only).
    """
    params = {}
    for name, (fan_out, fan_in) in layer_shapes(channels).items():
        bound = math.sqrt(6.0 / (fan_in + fan_out))
        w = (torch.rand(fan_out, fan_in, generator=generator) * 2 - 1) * bound
        b = (torch.rand(fan_out, generator=generator) * 2 - 1) * bias_scale
        params[name + ".weight"], params[name + ".bias"] = w, b
    return params


def mlp_forward(params, pts, viewdirs, barf=None):
    """pts [N,S,3], viewdirs [N,3] -> raw [N,S,C+1] = cat(rgb, sigma); no output activation.

    model/nerf.py:67-116: encode points and (per-sample broadcast) view
    directions, 8 ReLU layers with the encoded point re-concatenated IN FRONT of
    the hidden state after layer 4, sigma head on h, feature head (no ReLU),
    view layer on cat(feature, encoded dir), rgb head.
    """
    lin = lambda name, x: F.linear(x, params[name + ".weight"], params[name + ".bias"])
    flat = pts.reshape(-1, 3)
    enc_p = encode(flat, PTS_FREQS, barf)                 # barf = (progress, start, end): model/nerf.py:75-88
    dirs = viewdirs[:, None].expand(pts.shape).reshape(-1, 3)
    enc_d = encode(dirs, DIR_FREQS, barf)
    h = enc_p
    for i in range(D):
        h = F.relu(lin(f"pts_linears.{i}", h))
        if i == SKIP:
            h = torch.cat([enc_p, h], -1)
    sigma = lin("alpha_linear", h)
    feat = lin("feature_linear", h)
    h = F.relu(lin("views_linears.0", torch.cat([feat, enc_d], -1)))
    rgb = lin("rgb_linear", h)
    out = torch.cat([rgb, sigma], -1)
    return out.reshape(list(pts.shape[:-1]) + [out.shape[-1]])


# GEMM work per sample, forward: the figure roofline.achieved uses.
def macs_per_sample(channels=3):
    return sum(o * i for (o, i) in layer_shapes(channels).values())
