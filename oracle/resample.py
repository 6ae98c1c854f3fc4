"""Oracle: hierarchical (inverse-CDF) resampling and merge with the coarse depths.

Test infrastructure (see oracle/__init__.py).  Restates
run_nerf_helpers.py:74-115 (sample_pdf with det=False; the uniform draw u is an
explicit input = RNG draw #3) and model/nerf.py:322-326 (mid-points as bins,
weights[1:-1], detach, concatenate with the coarse depths, sort).
"""
import torch


def inverse_cdf_samples(bins, weights, u):
    """bins [N,B], weights [N,B-1], u [N,K] in [0,1) -> samples [N,K]."""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_lo, cdf_hi = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bin_lo, bin_hi = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = cdf_hi - cdf_lo
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_lo) / denom
    return bin_lo + t * (bin_hi - bin_lo)


def fine_depths(z_coarse, weights, u):
    """model/nerf.py:322-326 -> sorted [N, S_c + K] depths (no gradient flows through)."""
    mids = 0.5 * (z_coarse[..., 1:] + z_coarse[..., :-1])
    extra = inverse_cdf_samples(mids, weights[..., 1:-1], u).detach()
    z, _ = torch.sort(torch.cat([z_coarse, extra], -1), -1)
    return z


def conditioning(z_coarse, weights, u):
    """Per new sample: (bin mass, bin width) of the inverse-CDF bin it is drawn from.

    t = (u - cdf_lo) / mass: an fp32 rounding difference eps (~6e-8, one ulp of a cdf value) between
    two valid evaluations of the coarse pass moves the sample by eps / mass * width.  With the 1e-5
    weight floor, empty-space bins have mass ~2e-5, i.e. an amplification of ~3000: the reference's
    fine depths are only reproducible to ~1e-4 there (and the 2^9 positional-encoding frequency
    multiplies that again).  Parity harnesses use this to bound / flag such samples explicitly.
    Returns (mass [N,K], width [N,K]).
    """
    mids = 0.5 * (z_coarse[..., 1:] + z_coarse[..., :-1])
    w = weights[..., 1:-1] + 1e-5
    pdf = w / torch.sum(w, -1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[..., :1]), torch.cumsum(pdf, -1)], -1)
    inds = torch.searchsorted(cdf, u.contiguous(), right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    mass = torch.gather(cdf, 1, above) - torch.gather(cdf, 1, below)
    width = torch.gather(mids, 1, above) - torch.gather(mids, 1, below)
    return mass, width
