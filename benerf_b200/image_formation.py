"""Image formation of the blur and event models (train.py:163-177,205-331) on the CUDA library.

blur_mean / event_logdiff / accumulate_events are the C-ABI kernels (csrc/image_formation.cu); the loss block of an
iteration -- event pair split, target gather, both event losses, blur mean, rgb loss, and the gradients w.r.t. the four
rendered tensors -- is bnrf_training_loss (csrc/loss.cu), one launch (two for the normalised event loss).
"""
import torch

from .engine import (blur_mean, event_logdiff, accumulate_events, accumulate_events_binned, LOG_MODE, loss_cfg, training_loss_fused)  # noqa: F401

PART_KEYS = ("event_rgb_map", "event_rgb0", "blur_rgb_map", "blur_rgb0")


def mse(a, b):
    """loss/imgloss.py:3-5."""
    return torch.mean((a - b) ** 2)


def _rank_sum():
    """In-place sum over the data-parallel ranks (None in a single process): the five batch sums of the normalised loss."""
    from .parallel import world
    if world() == 1:
        return None
    import torch.distributed as dist
    return lambda t: dist.all_reduce(t)


class _TrainingLossFn(torch.autograd.Function):
    """loss_out [5] = bnrf_training_loss(...); the gradients w.r.t. the four renders come out of the same launch and are
    handed to autograd scaled by d L / d loss_out[0] (the other four entries are reporting values)."""

    @staticmethod
    def forward(ctx, cfg, events_accu, idx_evt, blur_target, evt_fine, evt_coarse, blur_fine, blur_coarse):
        loss_out, grads = training_loss_fused(cfg, evt_fine.contiguous(), evt_coarse.contiguous(), events_accu, idx_evt,
                                              blur_fine.contiguous(), blur_coarse.contiguous(), blur_target, all_reduce=_rank_sum())
        ctx.save_for_backward(*grads)
        return loss_out

    @staticmethod
    def backward(ctx, g):
        s = g[0].to(torch.float32)
        return (None, None, None, None) + tuple(d * s for d in ctx.saved_tensors)


def training_loss(ret_event, ret_rgb, events_accu, ray_idx_event, blur_target, args):
    """The loss block of train.py:163-337 on the fused kernel.  ret_event: Graph.render of the event pose pair ([2*R_e, C]
    pose-major), ret_rgb: of the N blur poses -- or what graph.event_crf / graph.rgb_crf made of them (train.py:176-192);
    events_accu [H_ev, W_ev] float64; blur_target [R_b, C].  Honours args.event_loss / args.rgb_loss.  Returns (loss, parts):
    loss is a float64 scalar, differentiable w.r.t. the renders; parts are detached reporting values."""
    dev = ret_event["rgb_map"].device
    idx = torch.as_tensor(ray_idx_event).reshape(-1).to(device=dev, dtype=torch.int64).contiguous()
    accu = torch.as_tensor(events_accu).to(device=dev, dtype=torch.float64).contiguous()
    tgt = torch.as_tensor(blur_target).to(device=dev, dtype=torch.float32).contiguous()
    out = _TrainingLossFn.apply(loss_cfg(args), accu, idx, tgt, ret_event["rgb_map"], ret_event["rgb0"], ret_rgb["rgb_map"], ret_rgb["rgb0"])
    parts = dict(zip(PART_KEYS, (out[1].detach(), out[2].detach(), out[3].detach(), out[4].detach())))
    return out[0], parts


def event_loss(diff, target, event_threshold, event_coeff_syn=0.1, event_coeff_real=2.0):
    """One level of train.py:207-292 from an already formed difference, as differentiable torch ops (operator-level helper for
    callers that form the difference themselves; the training path uses bnrf_training_loss).  diff [R,1] fp32, target [R,1]
    float64 (Q10).  With pixel-sharded ranks the squared norms are all-reduced (differentiably, parallel.global_sum), so every
    rank normalises by the norm of the whole batch exactly as a single process would."""
    if event_threshold > 0:
        return mse(diff, target * event_threshold) * event_coeff_syn
    from .parallel import global_sum
    dnorm = torch.sqrt(global_sum((diff * diff).sum(dim=0, keepdim=True)))
    tnorm = torch.sqrt(global_sum((target * target).sum(dim=0, keepdim=True)))
    return mse(diff / (dnorm + 1e-9), target / (tnorm + 1e-9)) * event_coeff_real


def accumulate_events_on_gpu(out, xs, ys, ps, device="cuda"):
    """Signature of utils/event_utils.py:247-259: numpy x/y/polarity of the window -> float64 [H,W]."""
    import numpy as np
    H, W = out.shape
    x = torch.as_tensor(np.asarray(xs), dtype=torch.int32, device=device)
    y = torch.as_tensor(np.asarray(ys), dtype=torch.int32, device=device)
    p = torch.as_tensor(np.asarray(ps), dtype=torch.float32, device=device)
    base = torch.as_tensor(out, dtype=torch.float64, device=device).clone()
    return accumulate_events(x, y, p, H, W, out=base)
