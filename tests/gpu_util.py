"""Helpers shared by the -m gpu parity tests (CUDA engine through the C ABI vs oracle/golden)."""
import torch

from tests.cases import CASES, make_inputs, load_golden

DEV = "cuda"


def to_dev(x):
    if isinstance(x, dict):
        return {k: to_dev(v) for k, v in x.items()}
    if isinstance(x, torch.Tensor):
        return x.to(DEV).contiguous()
    return x


def make_engine(case, mode):
    from benerf_b200.engine import Engine
    eng = Engine(n_samples=case.n_samples, n_importance=case.n_importance, channels=case.channels, mlp_mode=mode)
    eng.set_sample_grid(torch.linspace(0.0, 1.0, steps=case.n_samples))   # the oracle's own grid (see bnrf_set_sample_grid)
    if case.barf_iter >= 0:                                                # BARF c2f case: channel weights of its iter_step
        from argparse import Namespace
        from benerf_b200.nerf import barf_c2f_channel_weights
        from tests.cases import BARF_MAX_ITER, BARF_START, BARF_END
        eng.set_encoding_weights(*barf_c2f_channel_weights(case.barf_iter, Namespace(
            multires=10, multires_views=4, max_iter=BARF_MAX_ITER, barf_c2f_start=BARF_START, barf_c2f_end=BARF_END)))
    return eng


def max_abs(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    both_nan = torch.isnan(a) & torch.isnan(b)
    return float(torch.where(both_nan, torch.zeros_like(a), (a - b).abs()).max())
