"""Image formation of the blur and event models (train.py:163-177,205-331) on the CUDA library.

blur_mean / event_logdiff / accumulate_events are the C-ABI kernels (csrc/image_formation.cu).
The two scalar losses are tiny reductions over [R] tensors done with torch ops on the device,
laid out exactly as the reference's loss block so a training loop can call them in its place.
"""
import torch

from .engine import blur_mean, event_logdiff, accumulate_events, LOG_MODE  # noqa: F401


def mse(a, b):
    """loss/imgloss.py:3-5."""
    return torch.mean((a - b) ** 2)


def event_loss(diff, target, event_threshold, event_coeff_syn=0.1, event_coeff_real=2.0):
    """One level of train.py:207-292.  diff [R,1] fp32, target [R,1] float64 (Q10)."""
    if event_threshold > 0:
        return mse(diff, target * event_threshold) * event_coeff_syn
    # the norms run over the whole ray batch (dim 0): with pixel-sharded ranks the squared sums are all-reduced
    # (differentiably), so every rank normalises by the global norm exactly as a single process would
    from .parallel import global_sum
    dnorm = torch.sqrt(global_sum((diff * diff).sum(dim=0, keepdim=True)))
    tnorm = torch.sqrt(global_sum((target * target).sum(dim=0, keepdim=True)))
    dn = diff / (dnorm + 1e-9)
    tn = target / (tnorm + 1e-9)
    return mse(dn, tn) * event_coeff_real


def training_loss(ret_event, ret_rgb, events_accu, ray_idx_event, blur_target, args):
    """The loss block of train.py:163-337 (CRFs off, as in every shipped config) on the fused image-formation kernels.

    ret_event: Graph.render of the event pose pair ([2*R_e, C] pose-major), ret_rgb: of the N blur poses;
    events_accu [H_ev, W_ev] float64; blur_target [R_b, C].  Returns (loss, parts); differentiable end to end.
    """
    target = events_accu.reshape(-1, 1)[ray_idx_event]
    parts = {}
    for level in ("rgb_map", "rgb0"):
        diff = event_logdiff(ret_event[level], 1, args.dataset).reshape(-1, 1)
        parts["event_" + level] = event_loss(diff, target, args.event_threshold, getattr(args, "event_coeff_syn", 0.1),
                                             getattr(args, "event_coeff_real", 2.0))
        parts["blur_" + level] = mse(blur_mean(ret_rgb[level], args.num_interpolated_pose), blur_target) * getattr(args, "rgb_coeff", 1.0)
    loss = (parts["event_rgb0"] + parts["event_rgb_map"]) + (parts["blur_rgb_map"] + parts["blur_rgb0"])
    return loss, parts


def accumulate_events_on_gpu(out, xs, ys, ps, device="cuda"):
    """Signature of utils/event_utils.py:247-259: numpy x/y/polarity of the window -> float64 [H,W]."""
    import numpy as np
    H, W = out.shape
    x = torch.as_tensor(np.asarray(xs), dtype=torch.int32, device=device)
    y = torch.as_tensor(np.asarray(ys), dtype=torch.int32, device=device)
    p = torch.as_tensor(np.asarray(ps), dtype=torch.float32, device=device)
    base = torch.as_tensor(out, dtype=torch.float64, device=device).clone()
    return accumulate_events(x, y, p, H, W, out=base)
