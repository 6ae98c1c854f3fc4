"""Mirror of the run_nerf_helpers.py entry points the callers use around the hot path."""
import torch


def init_weights(linear):
    """run_nerf_helpers.py:194-197: Xavier-uniform weight, zero bias."""
    torch.nn.init.xavier_uniform_(linear.weight)
    torch.nn.init.zeros_(linear.bias)


def init_nerf(nerf):
    """run_nerf_helpers.py:199-208."""
    for linear_pt in nerf.pts_linears:
        init_weights(linear_pt)
    for linear_view in nerf.views_linears:
        init_weights(linear_view)
    init_weights(nerf.feature_linear)
    init_weights(nerf.alpha_linear)
    init_weights(nerf.rgb_linear)
