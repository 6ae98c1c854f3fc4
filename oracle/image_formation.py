"""Oracle: blur and event image formation + losses.

Test infrastructure (see oracle/__init__.py).  Restates
  train.py:163-177   split of the event render into start/end halves, target gather
  train.py:205-292   event loss, synthetic (threshold > 0) and real (normalised)
  train.py:299-331   blur = mean of the P sharp renders, rgb loss
  utils/img_utils.py:7-16   RGB2Gray (0.299, 0.587, 0.114)
  utils/math_utils.py:4-23  safe_log / lin_log / rgb2brightlog dataset switch
  loss/imgloss.py:3-5       MSELoss
"""
import torch

GRAY = (0.299, 0.587, 0.114)
LINLOG_DATASETS = ("E2NeRF_Synthetic", "E2NeRF_Real")
SAFELOG_DATASETS = ("BeNeRF_Blender", "BeNeRF_Unreal")


def to_gray(rgb):
    """[R,3] -> [R,1] (utils/img_utils.py:13-16)."""
    w = torch.tensor(GRAY)
    x = torch.sum(rgb * w[None, :], dim=-1)
    return x.reshape(x.shape[0], 1)


def log_brightness(x, dataset):
    """utils/math_utils.py:18-23."""
    if dataset in SAFELOG_DATASETS:
        return torch.log(x + 1e-9)
    if dataset in LINLOG_DATASETS:
        c = x * 255
        thres = 20
        slope = torch.log(torch.tensor(thres) + 1e-9) / thres
        return torch.where(c < thres, slope * c, torch.log(c + 1e-9))
    raise ValueError(dataset)


def mse(a, b):
    return torch.mean((a - b) ** 2)


def blur_mean(rgb, n_poses):
    """[P*R,C] pose-major -> [R,C]: running sum j = 0..P-1 then one divide (train.py:307-318)."""
    r = rgb.shape[0] // n_poses
    acc = 0
    for j in range(n_poses):
        acc = acc + rgb[j * r:(j + 1) * r]
    return acc / n_poses


def event_log_diff(rgb, dataset, channels=3):
    """[2R,C] (start poses first, train.py:166-173) -> [R,1] log-brightness difference."""
    r = rgb.shape[0] // 2
    first, second = rgb[:r], rgb[r:]
    if channels == 3:
        first, second = to_gray(first), to_gray(second)
    return log_brightness(second, dataset) - log_brightness(first, dataset)


def event_loss(diff, target, threshold, coeff_syn=0.1, coeff_real=2.0):
    """train.py:207-292 for one level (fine or coarse).  target [R,1] is float64 (Q10).

    threshold > 0: mse(diff, target * threshold) * coeff_syn.  Otherwise both
    sides are divided by their L2 norm over the whole ray batch (+1e-9) and the
    mse is scaled by coeff_real.
    """
    if threshold > 0:
        return mse(diff, target * torch.tensor(threshold)) * coeff_syn
    dn = diff / (torch.linalg.norm(diff, dim=0, keepdim=True) + 1e-9)
    tn = target / (torch.linalg.norm(target, dim=0, keepdim=True) + 1e-9)
    return mse(dn, tn) * coeff_real


def training_loss(ret_event, ret_rgb, events_accu, ray_idx_event, blur_target, *, n_poses,
                  dataset, channels=3, threshold=0.1, coeff_syn=0.1, coeff_real=2.0, rgb_coeff=1.0):
    """Total loss of one iteration (train.py:163-337, CRFs off as in all shipped configs).

    events_accu [H_ev,W_ev] float64; blur_target [R_b,C] fp32 (the blurry pixels at
    ray_idx_rgb).  Returns (loss, parts dict).
    """
    target = events_accu.reshape(-1, 1)[ray_idx_event]
    parts = {}
    for level in ("rgb_map", "rgb0"):
        diff = event_log_diff(ret_event[level], dataset, channels)
        parts["event_" + level] = event_loss(diff, target, threshold, coeff_syn, coeff_real)
        parts["blur_" + level] = mse(blur_mean(ret_rgb[level], n_poses), blur_target) * rgb_coeff
    loss = (parts["event_rgb0"] + parts["event_rgb_map"]) + (parts["blur_rgb_map"] + parts["blur_rgb0"])
    return loss, parts
