"""Host-side logic of the product package that needs no GPU: the BARF coarse-to-fine channel weights, the pixel sharding of the
data-parallel path, the loss configuration handed to the fused loss kernel, and the learning-rate schedule the fused optimiser
tail evaluates on the device (restated here from the reference's train.py:355-394 and checked against the host mirror)."""
import math
from argparse import Namespace

import pytest
import torch


@pytest.mark.parametrize("iter_step", [0, 50, 100, 137, 300, 499, 500, 800, 1000])
def test_barf_channel_weights_equal_the_oracle_weighting(iter_step):
    """nerf.barf_c2f_channel_weights folds model/nerf.py:16-26 into per-channel factors (applied to weight-matrix columns by
    bnrf_set_encoding_weights).  Weighting an encoding channel by channel with them must equal the oracle's literal restatement
    (the [M, 6L] encoding viewed as (-1, L)) -- including before the ramp starts and after it ends."""
    from benerf_b200.nerf import barf_c2f_channel_weights
    from oracle import encode
    args = Namespace(multires=10, multires_views=4, max_iter=1000, barf_c2f_start=0.1, barf_c2f_end=0.5)
    pts_w, dir_w = barf_c2f_channel_weights(iter_step, args)
    assert len(pts_w) == 63 and len(dir_w) == 27 and pts_w[:3] == (1.0, 1.0, 1.0) and dir_w[:3] == (1.0, 1.0, 1.0)
    g = torch.Generator().manual_seed(iter_step)
    for w, L in ((pts_w, 10), (dir_w, 4)):
        x = torch.rand(7, 3, generator=g) * 2 - 1
        want = encode.encode(x, L, barf=(iter_step / args.max_iter, args.barf_c2f_start, args.barf_c2f_end))
        got = encode.encode(x, L) * torch.tensor(w)
        assert torch.equal(got, want)
        assert all(0.0 <= v <= 1.0 for v in w)
    if iter_step / args.max_iter <= args.barf_c2f_start:
        assert set(pts_w[3:]) == {0.0}                      # nothing but the raw input before the ramp
    if iter_step / args.max_iter >= args.barf_c2f_end:
        assert set(pts_w[3:]) == {1.0}                      # the plain encoding after it


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_shard_partitions_the_batch_into_equal_disjoint_slices(world):
    from benerf_b200.parallel import shard
    t = torch.arange(4081 * 2).reshape(4081, 2)
    parts = [shard(t, r, world) for r in range(world)]
    per = 4081 // world
    assert all(p.shape == (per, 2) for p in parts)
    cat = torch.cat(parts)
    assert torch.equal(cat, t[:per * world])                 # contiguous, ordered, no overlap; only the tail (< world rows) is dropped
    assert 4081 - per * world < world


def test_loss_cfg_follows_the_reference_flags():
    """engine.loss_cfg: args.event_loss / args.rgb_loss (store_true flags, config.py:215-218) default to on as in the shipped configs;
    the log mode follows the dataset (utils/math_utils.py:4-23); coefficients are passed through."""
    from benerf_b200.engine import loss_cfg, LOG_MODE
    base = dict(channels=3, num_interpolated_pose=19, dataset="BeNeRF_Unreal", event_threshold=0.1)
    c = loss_cfg(Namespace(**base))
    assert (c.channels, c.n_poses, c.log_mode, c.event_loss, c.rgb_loss) == (3, 19, LOG_MODE["BeNeRF_Unreal"], 1, 1)
    assert abs(c.event_threshold - 0.1) < 1e-7 and abs(c.event_coeff_syn - 0.1) < 1e-7 and c.event_coeff_real == 2.0 and c.rgb_coeff == 1.0
    c = loss_cfg(Namespace(**dict(base, dataset="E2NeRF_Real", event_threshold=-1.0, event_loss=False, rgb_loss=True,
                                  event_coeff_real=3.0, rgb_coeff=0.5)))
    assert (c.log_mode, c.event_loss, c.rgb_loss) == (LOG_MODE["E2NeRF_Real"], 0, 1) and c.event_coeff_real == 3.0 and c.rgb_coeff == 0.5
    assert LOG_MODE["E2NeRF_Real"] != LOG_MODE["BeNeRF_Unreal"]     # real data: piecewise-linear toe below 20 / 255


def test_device_side_learning_rate_schedule_equals_the_reference_loop():
    """bnrf_adam_step_sched derives the learning rate of iteration g (0-based global_step) on the device as
    lr0 for g = 0 and lr0 * rate ** ((g - 1) / decay_steps) afterwards.  That is the reference's loop: optimizer.step() of
    iteration g uses the rate set at the END of iteration g - 1 from the not yet incremented counter (train.py:343-394)."""
    lr0, rate, decay_steps = 5e-4, 0.1, 200 * 1000
    lr, global_step, used = lr0, 0, []
    for _ in range(6):
        used.append(lr)                                      # optimizer.step()
        lr = lr0 * rate ** (global_step / decay_steps)        # "new_lrate" written into param_groups after the step
        global_step += 1
    device_rule = [lr0 if g == 0 else lr0 * rate ** ((g - 1) / decay_steps) for g in range(6)]
    assert all(math.isclose(a, b, rel_tol=0, abs_tol=0) for a, b in zip(used, device_rule))
