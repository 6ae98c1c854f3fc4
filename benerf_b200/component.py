"""Mirror of model/component.py: the pose / transform parameter tables and the two camera-response tone mappers.

ControlKnotLieAlgebra / TransformationLieAlgebra are nn.Embedding tables exactly as upstream (model/component.py:7-15).
ColorToneMapper / LuminanceToneMapper own the same nn.Sequential of Linear / ReLU modules under the same names (so state
dicts round-trip) and forward through bnrf_crf_forward / bnrf_crf_backward (csrc/crf.cu).  Upstream they are
constructed with input_type "Gray" only (model/optimize.py:15-20) and applied when args.optimize_rgb_crf /
optimize_event_crf is set (train.py:176-192, run_nerf_helpers.py:125-126,152-153); the "RGB" variant indexes a 1-D slice into
Linear(1, width) and cannot run upstream either, so it is rejected here.
"""
import ctypes as C

import torch
import torch.nn as nn
import torch.nn.init as init


class ControlKnotLieAlgebra(nn.Module):
    def __init__(self, knot_num):
        super().__init__()
        self.params = nn.Embedding(knot_num, 6)


class TransformationLieAlgebra(nn.Module):
    def __init__(self, trans_num):
        super().__init__()
        self.params = nn.Embedding(trans_num, 6)


class _CrfFn(torch.autograd.Function):
    """sigmoid(mlp(x)) element-wise over an [N, 1] tensor; parameters in nn.Sequential order (weight, bias per Linear)."""

    @staticmethod
    def forward(ctx, x, width, hidden, *params):
        from . import _lib
        from .engine import _ptr, _stream, _rc
        lib = _lib.load()
        xs = x.detach().to(torch.float32).contiguous()
        ps = [p.detach().contiguous() for p in params]
        ws = (C.c_void_p * (hidden + 2))(*[_ptr(p, name="crf weight").value for p in ps[0::2]])
        bs = (C.c_void_p * (hidden + 2))(*[_ptr(p, name="crf bias").value for p in ps[1::2]])
        y = torch.empty_like(xs)
        _rc(lib.bnrf_crf_forward(width, hidden, ws, bs, _ptr(xs), xs.numel(), _ptr(y), _stream()), "bnrf_crf_forward")
        ctx.save_for_backward(xs, *ps)
        ctx.dims = (width, hidden)
        return y.view_as(x)

    @staticmethod
    def backward(ctx, g):
        from . import _lib
        from .engine import _ptr, _stream, _rc
        lib = _lib.load()
        width, hidden = ctx.dims
        xs, *ps = ctx.saved_tensors
        gs = g.to(torch.float32).contiguous()
        ws = (C.c_void_p * (hidden + 2))(*[_ptr(p).value for p in ps[0::2]])
        bs = (C.c_void_p * (hidden + 2))(*[_ptr(p).value for p in ps[1::2]])
        d = [torch.zeros_like(p) for p in ps]
        dws = (C.c_void_p * (hidden + 2))(*[_ptr(t).value for t in d[0::2]])
        dbs = (C.c_void_p * (hidden + 2))(*[_ptr(t).value for t in d[1::2]])
        dx = torch.empty_like(xs)
        _rc(lib.bnrf_crf_backward(width, hidden, ws, bs, _ptr(xs), _ptr(gs), xs.numel(), _ptr(dx), dws, dbs, _stream()), "bnrf_crf_backward")
        return (dx.view_as(g), None, None) + tuple(d)


class _ToneMapper(nn.Module):
    attr = None            # name of the nn.Sequential (state-dict prefix): "mlp_gray" / "mlp_luminance"
    bias_init = staticmethod(init.zeros_)

    def __init__(self, hidden=0, width=128, input_type="Gray"):
        super().__init__()
        self.net_hidden, self.net_width, self.input_type = hidden, width, str(input_type)
        if self.input_type != "Gray":
            raise NotImplementedError('tone mapper input_type "RGB" feeds a 1-D slice into Linear(1, width) upstream '
                                      "(model/component.py:88-96) and cannot run there either; only \"Gray\" is built (model/optimize.py:15-20)")
        layers = [nn.Linear(1, width), nn.ReLU()]
        for _ in range(hidden):
            layers += [nn.Linear(width, width), nn.ReLU()]
        layers.append(nn.Linear(width, 1))
        setattr(self, self.attr, nn.Sequential(*layers))

    def _linears(self):
        return [m for m in getattr(self, self.attr) if isinstance(m, nn.Linear)]

    def weights_biases_init(self):
        for layer in self._linears():
            init.xavier_uniform_(layer.weight)
            self.bias_init(layer.bias)

    def forward(self, radience):
        if radience.shape[-1] != 1:
            raise ValueError(f"tone mappers take [N, 1] (gray) input, got {tuple(radience.shape)}: Linear(1, width) (model/component.py:51)")
        params = [t for m in self._linears() for t in (m.weight, m.bias)]
        return _CrfFn.apply(radience, self.net_width, self.net_hidden, *params)


class ColorToneMapper(_ToneMapper):
    """A network for color tone-mapping (model/component.py:38-104): Xavier weights, zero biases."""
    attr = "mlp_gray"

    def constraint_radience_scale(self, fixed_value=0.5):
        dev = next(self.parameters()).device
        color_0 = self.forward(torch.zeros(1, 1, device=dev))
        return torch.mean((color_0 - fixed_value) ** 2)


class LuminanceToneMapper(_ToneMapper):
    """A network for luminance tone-mapping (model/component.py:106-149): Xavier weights, biases of one."""
    attr = "mlp_luminance"
    bias_init = staticmethod(init.ones_)
