"""Oracle: sin/cos positional encoding.

Test infrastructure (see oracle/__init__.py).  Restates model/embedder.py:9-52:
[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)], every term a
3-vector block, frequencies exact powers of two (2 ** linspace(0, L-1, L)).
"""
import torch


def positional_encoding(x, n_freqs, include_input=True):
    """[M,3] -> [M, 3*(include_input + 2*n_freqs)]  (63 for L=10, 27 for L=4)."""
    bands = 2.0 ** torch.linspace(0.0, n_freqs - 1, steps=n_freqs)
    parts = [x] if include_input else []
    for f in bands:
        parts.append(torch.sin(x * f))
        parts.append(torch.cos(x * f))
    return torch.cat(parts, -1)
