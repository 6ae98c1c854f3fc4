"""Oracle: camera-response tone mappers.

Test infrastructure (see oracle/__init__.py).  Restates model/component.py:38-149 for input_type "Gray":
ColorToneMapper.forward / LuminanceToneMapper.forward = sigmoid(Sequential(Linear(1, w), ReLU, [Linear(w, w), ReLU] * hidden,
Linear(w, 1))(x)) on an [N, 1] tensor.  `params` is the Sequential's parameter list: weight, bias per Linear, in order.
"""
import torch


def tone_map(params, x):
    h = x
    n = len(params) // 2
    for l in range(n):
        h = torch.nn.functional.linear(h, params[2 * l], params[2 * l + 1])
        if l < n - 1:
            h = torch.relu(h)
    return torch.sigmoid(h)
