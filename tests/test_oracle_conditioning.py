"""The reference's inverse-CDF resampling is ill-conditioned in near-empty bins.

This is a property of the reference algorithm (run_nerf_helpers.py:74-115), shown here on the
oracle alone: perturbing the coarse weights by 2e-7 relative -- the size of an exp/GEMM rounding
difference between two valid fp32 evaluations -- moves some fine depths by > 1e-5 and the fine
sigma at those samples by > 1e-4, while every sample drawn from a bin of mass >= 2e-3 stays put.
The GPU parity tests rely on exactly this split (tests/test_gpu_render.py).
"""
import torch

from oracle import rays, resample, mlp, composite
from tests.cases import CASES, make_inputs, load_golden


def test_reference_fine_depths_are_chaotic_only_in_near_empty_bins():
    case, gold = CASES["unreal_rgb"], load_golden("unreal_rgb")
    inp = make_inputs(case)
    z, w, u = gold["rgb_z_c"], gold["rgb_weights_c"], inp["rng_rgb"]["u"]
    g = torch.Generator().manual_seed(0)
    w2 = w * (1 + 2e-7 * torch.randn(w.shape, generator=g))
    mass, width = resample.conditioning(z, w, u)
    zs1 = resample.inverse_cdf_samples(0.5 * (z[..., 1:] + z[..., :-1]), w[..., 1:-1], u)
    zs2 = resample.inverse_cdf_samples(0.5 * (z[..., 1:] + z[..., :-1]), w2[..., 1:-1], u)
    dz = (zs1 - zs2).abs()
    well = mass >= 2e-3
    assert dz[well].max() < 2e-6
    assert dz[~well].max() > 1e-5, "expected visible chaos in near-empty bins"
    # analytic bound used by the GPU tests: |dz| <= 2e-7 + 4e-7 / mass * width
    assert bool((dz <= 2e-7 + 4e-7 / mass.clamp_min(1e-5) * width.abs() + 1e-9).all())
    frac = float((~well).any(-1).float().mean())
    assert 0.02 < frac < 0.5     # a sizeable share of rays holds at least one such sample


def test_fine_outputs_follow_the_same_split():
    case, gold = CASES["unreal_rgb"], load_golden("unreal_rgb")
    inp = make_inputs(case)
    K = torch.tensor(case.K, dtype=torch.float32)
    o, d, v = rays.ray_batch(gold["poses_rgb"], inp["idx_rgb"], case.H, case.W, K)
    z, w, draws = gold["rgb_z_c"], gold["rgb_weights_c"], inp["rng_rgb"]
    g = torch.Generator().manual_seed(1)
    w2 = w * (1 + 2e-7 * torch.randn(w.shape, generator=g))
    outs = []
    for ww in (w, w2):
        zf = resample.fine_depths(z, ww, draws["u"])
        raw = mlp.mlp_forward(inp["fine"], rays.sample_points(o, d, zf), v)
        outs.append(composite.composite(raw, zf, d, draws["noise_f"], case.channels))
    mass, _ = resample.conditioning(z, w, draws["u"])
    well = mass.min(-1)[0] >= 2e-3
    err_rgb = (outs[0]["rgb_map"] - outs[1]["rgb_map"]).abs().amax(-1)
    err_sig = (outs[0]["sigma"] - outs[1]["sigma"]).abs().amax(-1)
    assert err_rgb[well].max() < 1e-5 and err_sig[well].max() < 1e-4
    assert err_sig[~well].max() > 1e-4


def test_feature_linear_composes_with_the_view_layer():
    """feature_linear has no activation (model/nerf.py:102-105), so views_linears.0(cat([feature_linear(h), dirs])) is ONE
    linear map of cat([h, dirs]): W_m = W_views[:, :256] @ W_feature, b_m = b_views + W_views[:, :256] @ b_feature.  The CUDA
    engine runs that merged step (csrc/common.cuh: wt9m); this pins the algebra, in fp32, on the fixtures' own weights:
    the two evaluation orders agree to ~1e-6, two orders of magnitude inside the 1e-4 parity bound."""
    import torch.nn.functional as F
    from tests.cases import CASES, make_inputs
    for name in ("unreal_rgb", "gray_linear"):
        p = make_inputs(CASES[name])["coarse"]
        g = torch.Generator().manual_seed(3)
        h = torch.relu(torch.randn(4096, 256, generator=g))
        d = torch.randn(4096, 27, generator=g).clamp(-1, 1)
        feat = F.linear(h, p["feature_linear.weight"], p["feature_linear.bias"])
        want = torch.relu(F.linear(torch.cat([feat, d], -1), p["views_linears.0.weight"], p["views_linears.0.bias"]))
        wv, wd = p["views_linears.0.weight"][:, :256], p["views_linears.0.weight"][:, 256:]
        w_m = wv @ p["feature_linear.weight"]
        b_m = p["views_linears.0.bias"] + wv @ p["feature_linear.bias"]
        got = torch.relu(F.linear(h, w_m, b_m) + F.linear(d, wd))
        assert float((got - want).abs().max()) < 5e-6
        # and in float64 the identity is exact to rounding of the inputs
        got64 = torch.relu(F.linear(h.double(), (wv.double() @ p["feature_linear.weight"].double()),
                                    p["views_linears.0.bias"].double() + wv.double() @ p["feature_linear.bias"].double()) + F.linear(d.double(), wd.double()))
        want64 = torch.relu(F.linear(torch.cat([F.linear(h.double(), p["feature_linear.weight"].double(), p["feature_linear.bias"].double()), d.double()], -1),
                                     p["views_linears.0.weight"].double(), p["views_linears.0.bias"].double()))
        assert float((got64 - want64).abs().max()) < 1e-12
