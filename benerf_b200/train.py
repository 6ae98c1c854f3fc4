"""One optimisation step of the reference's loop (train.py:153-394) on the CUDA engine, data-parallel over pixels.

train.py itself stays the caller's script; this helper is the step it performs per iteration, used by
bench.py and the tests: poses -> two Graph.render calls -> image formation + the four loss terms -> backward ->
ONE gradient all-reduce -> the reference's Adam steps and exponential learning-rate decay (Q17).
"""
import torch

from . import image_formation as IF
from .engine import adam_step
from .parallel import FlatGrads, world


class Trainer:
    def __init__(self, model, args):
        """model: benerf_b200.optimize.Model after build_network(); args: the reference's flag namespace."""
        self.model, self.graph, self.args = model, model.graph, args
        self.optims = model.setup_optimizer(args)                       # nerf, pose, transform, rgb_crf, event_crf (optimize.py:36-55)
        g = self.graph
        params = list(g.nerf.parameters()) + (list(g.nerf_fine.parameters()) if hasattr(g, "nerf_fine") else [])
        params += [g.evt_knot_pose_se3.params.weight, g.transform.params.weight]
        self.fused = bool(getattr(args, "fused_optimizer", True))
        if self.fused:
            # parameters, like their gradients, become views of ONE flat buffer (same order), so that the optimiser tail is a
            # single launch (bnrf_adam_step); state_dict / load_state_dict / init_nerf keep working on the views
            self.flat_params = torch.cat([p.data.reshape(-1) for p in params])
            off, n_nerf = 0, sum(p.numel() for p in params[:-2])
            for p in params:
                p.data = self.flat_params[off:off + p.numel()].view_as(p)
                off += p.numel()
            self.exp_avg, self.exp_avg_sq = torch.zeros_like(self.flat_params), torch.zeros_like(self.flat_params)
            self.group_ranges = [(0, n_nerf), (n_nerf, n_nerf + 24), (n_nerf + 24, n_nerf + 30)]      # nerf(s), knots [4,6], transform [1,6]
        self.flat = FlatGrads(params)
        self.base_lr = [[grp["lr"] for grp in o.param_groups] for o in self.optims]
        self.global_step = 0
        self.phase_ms = None           # set to a list to get synchronised per-phase wall times (debug aid; serialises the step)

    def step(self, events_accu, idx_evt, idx_rgb, blur_target, ts_evt, ts_rgb, H, W, K, K_event, H_ev=None, W_ev=None):
        """idx_* are THIS rank's pixels; events_accu [H_ev, W_ev] float64; blur_target [R_rgb, C].  Returns (loss, parts)."""
        g, a = self.graph, self.args
        mark = self._mark
        mark(None)
        poses_evt = g.get_pose_evt(a, ts_evt)
        poses_rgb = g.get_pose_rgb(a, ts_rgb)
        ret_evt = g.render(self.global_step, poses_evt, idx_evt, H_ev or H, W_ev or W, K_event, a, enable_crf=True, sensor_type="event",
                           remap=None, training=True)
        ret_rgb = g.render(self.global_step, poses_rgb, idx_rgb, H, W, K, a, enable_crf=True, sensor_type="rgb", remap=None, training=True)
        mark("forward")
        loss, parts = IF.training_loss(ret_evt, ret_rgb, events_accu, idx_evt, blur_target, a)
        if not self.fused:
            self.flat.zero()
        mark("loss")
        loss.backward()
        mark("backward")
        opt_nerf, opt_pose, opt_trans = self.optims[0], self.optims[1], self.optims[2]
        flags = [getattr(a, "optimize_nerf", True), getattr(a, "optimize_pose", True), getattr(a, "optimize_trans", False)]
        if self.fused:
            self.flat.all_reduce_sum()                                   # the single exchange of the step
            mark("all_reduce")
            groups = [(b, e, o.param_groups[0]["lr"], f) for (b, e), o, f in zip(self.group_ranges, self.optims[:3], flags)]
            adam_step(self.flat_params, self.flat.flat, self.exp_avg, self.exp_avg_sq, groups, self.global_step + 1,
                      grad_scale=1.0 / world(), zero_grads=True)         # averages, steps all three optimisers, clears the gradients
            g.engine(a).invalidate_weights()                             # in-place update torch's version counters do not see
        else:
            self.flat.all_reduce_mean()
            mark("all_reduce")
            if flags[0]:
                opt_nerf.step()
            if flags[1]:
                opt_pose.step()
            if flags[2]:
                opt_trans.step()
        # lr = lr0 * rate ** (step / (lrate_decay * 1000)) with one rate per optimiser, applied after the step (train.py:355-394)
        decay_steps = getattr(a, "lrate_decay", 200) * 1000
        rates = [getattr(a, k, d) for k, d in (("decay_rate", 0.1), ("decay_rate_pose", 0.01), ("decay_rate_transform", 0.01),
                                               ("decay_rate_rgb_crf", 0.1), ("decay_rate_event_crf", 0.1))]
        for o, base, rate in zip(self.optims, self.base_lr, rates):
            for grp, lr0 in zip(o.param_groups, base):
                grp["lr"] = lr0 * rate ** (self.global_step / decay_steps)
        mark("optimizer")
        self.global_step += 1
        return loss.detach(), parts

    def _mark(self, name):
        if self.phase_ms is None:
            return
        import time
        torch.cuda.synchronize()
        now = time.perf_counter()
        if name is not None:
            self.phase_ms.append((self.global_step, name, round((now - self._t_prev) * 1e3, 2)))
        self._t_prev = now
