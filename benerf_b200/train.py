"""One optimisation step of the reference's loop (train.py:153-394) on the CUDA engine, data-parallel over pixels.

train.py itself stays the caller's script; this helper is the step it performs per iteration, used by
bench.py and the tests: poses -> two Graph.render calls -> image formation + the four loss terms -> backward ->
ONE gradient all-reduce -> the reference's Adam steps and exponential learning-rate decay (Q17).
"""
import torch

from . import image_formation as IF
from .parallel import FlatGrads


class Trainer:
    def __init__(self, model, args):
        """model: benerf_b200.optimize.Model after build_network(); args: the reference's flag namespace."""
        self.model, self.graph, self.args = model, model.graph, args
        self.optims = model.setup_optimizer(args)                       # nerf, pose, transform, rgb_crf, event_crf (optimize.py:36-55)
        g = self.graph
        params = list(g.nerf.parameters()) + (list(g.nerf_fine.parameters()) if hasattr(g, "nerf_fine") else [])
        params += [g.evt_knot_pose_se3.params.weight, g.transform.params.weight]
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
        self.flat.zero()
        mark("loss")
        loss.backward()
        mark("backward")
        self.flat.all_reduce_mean()                                      # the single exchange of the step
        mark("all_reduce")
        opt_nerf, opt_pose, opt_trans = self.optims[0], self.optims[1], self.optims[2]
        if getattr(a, "optimize_nerf", True):
            opt_nerf.step()
        if getattr(a, "optimize_pose", True):
            opt_pose.step()
        if getattr(a, "optimize_trans", False):
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
