"""Oracle: sin/cos positional encoding.

Test infrastructure (see oracle/__init__.py).  Restates model/embedder.py:9-52:
[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)], every term a
3-vector block, frequencies exact powers of two (2 ** linspace(0, L-1, L)).
"""
import math

import torch


def positional_encoding(x, n_freqs, include_input=True):
    """[M,3] -> [M, 3*(include_input + 2*n_freqs)]  (63 for L=10, 27 for L=4)."""
    bands = 2.0 ** torch.linspace(0.0, n_freqs - 1, steps=n_freqs)
    parts = [x] if include_input else []
    for f in bands:
        parts.append(torch.sin(x * f))
        parts.append(torch.cos(x * f))
    return torch.cat(parts, -1)


def barf_c2f_weight(embedded, n_freqs, progress, start, end):
    """model/nerf.py:16-26 (use_barf_c2f): the sin/cos part of an encoding ([M, 6L], no raw input) times a raised-cosine ramp.
    Upstream views the [M, 6L] tensor as (-1, L) before multiplying by the L weights, so channel e of a sample gets
    weight[e % L] -- NOT the weight of its own frequency e // 6; restated literally."""
    L = n_freqs
    alpha = (progress - start) / (end - start) * L
    k = torch.arange(L)
    weight = (1 - (alpha - k).clamp_(min=0, max=1).mul_(math.pi).cos_()) / 2
    shape = embedded.shape
    return (embedded.reshape(-1, L) * weight).reshape(*shape)


def encode(x, n_freqs, barf=None):
    """Encoding fed to the network: model/embedder.py (+ model/nerf.py:75-88 when barf = (progress, start, end))."""
    if barf is None:
        return positional_encoding(x, n_freqs)
    return torch.cat([x, barf_c2f_weight(positional_encoding(x, n_freqs, include_input=False), n_freqs, *barf)], -1)
