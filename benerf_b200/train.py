"""One optimisation step of the reference's loop (train.py:153-394) on the CUDA engine, data-parallel over pixels.

train.py itself stays the caller's script; this helper is the step it performs per iteration, used by
bench.py and the tests: poses -> the two Graph.render calls of an iteration -> image formation + the four loss terms -> backward ->
the gradient all-reduce (one exchange of the flat buffer, issued in two parts so that the fine network's half overlaps the coarse
network's backward pass) -> the reference's Adam steps and exponential learning-rate decay (Q17).

Two executions of the same arithmetic:
  * the DIRECT step (default, fused_optimizer=True): every stage is an explicit call into libbenerf_b200.so -- bnrf_set_weights_pair,
    bnrf_spline_poses_pair (on a second stream, beside the packing), bnrf_render_forward_multi (event pose pair + blur poses as one
    ray batch), bnrf_training_loss (loss + the gradients w.r.t. the renders), bnrf_render_backward_multi,
    bnrf_spline_poses_pair_backward, all-reduce, bnrf_adam_step_sched -- with no autograd graph in between: 28 kernels.
    Everything that changes from iteration to iteration (global_step -> Adam bias corrections, the three decayed learning rates,
    the Philox stream offset) lives in DEVICE memory, so after two eager iterations the step is captured in a CUDA graph and
    replayed: the host enqueues one graph launch (round 1's strong-scaled 8-GPU step was bounded by the 2.6 ms the host needed
    for ~80 launches).
  * the AUTOGRAD step (fused_optimizer=False, or a tone-mapper being optimised): Graph.render / image_formation.training_loss
    under torch autograd and the reference's own torch.optim.Adam objects; the reference-shaped cross-check of the former.
"""
import torch

from . import image_formation as IF
from .engine import adam_step_sched, loss_cfg, training_loss_fused
from .parallel import FlatGrads, world, rank
from ._lib import LINEAR_NAMES


class Trainer:
    def __init__(self, model, args):
        """model: benerf_b200.optimize.Model after build_network(); args: the reference's flag namespace."""
        self.model, self.graph, self.args = model, model.graph, args
        self.optims = model.setup_optimizer(args)                       # nerf, pose, transform, rgb_crf, event_crf (optimize.py:36-55)
        g = self.graph
        self.nets = [g.nerf] + ([g.nerf_fine] if hasattr(g, "nerf_fine") else [])
        # flat layout: fine network, coarse network, knots, transform -- the backward pass finishes the fine network first, so its
        # gradients are one leading range that can be all-reduced early, and everything else is one trailing range
        params = [p for m in reversed(self.nets) for p in m.parameters()]
        params += [g.evt_knot_pose_se3.params.weight, g.transform.params.weight]
        if not all(p.requires_grad for p in params):
            raise ValueError("Trainer lays parameters, gradients and moments out as flat buffers in one order: freeze an optimiser "
                             "with args.optimize_nerf / optimize_pose / optimize_trans, not with requires_grad=False")
        self.params = params
        self.crf = bool(getattr(args, "optimize_rgb_crf", False) or getattr(args, "optimize_event_crf", False))
        self.fused = bool(getattr(args, "fused_optimizer", True))
        # BARF c2f changes the packed weights with iter_step on the host: such runs step eagerly
        self.use_graph = self.fused and not self.crf and bool(getattr(args, "cuda_graph", True)) and not getattr(args, "use_barf_c2f", False)
        if self.fused:
            # parameters, like their gradients, become views of ONE flat buffer (same order), so that the optimiser tail is a
            # single launch (bnrf_adam_step_sched); state_dict / load_state_dict / init_nerf keep working on the views
            self.flat_params = torch.cat([p.data.reshape(-1) for p in params])
            off, n_nerf = 0, sum(p.numel() for p in params[:-2])
            for p in params:
                p.data = self.flat_params[off:off + p.numel()].view_as(p)
                off += p.numel()
            self.exp_avg, self.exp_avg_sq = torch.zeros_like(self.flat_params), torch.zeros_like(self.flat_params)
            self.group_ranges = [(0, n_nerf), (n_nerf, n_nerf + 24), (n_nerf + 24, n_nerf + 30)]      # nerf(s), knots [4,6], transform [1,6]
            self.step_dev = torch.zeros(1, device=self.flat_params.device, dtype=torch.int64)        # train.py's global_step, on the device
            self.step_scratch = torch.zeros(1, device=self.flat_params.device, dtype=torch.int64)    # block arrival count of the Adam launch
        self.flat = FlatGrads(params)
        self._side = torch.cuda.Stream(device=params[0].device) if params[0].is_cuda else None   # issue point of the early all-reduce
        self.base_lr = [[grp["lr"] for grp in o.param_groups] for o in self.optims]
        self.global_step = 0
        self.phase_ms = None           # set to a list to get synchronised per-phase wall times (debug aid; serialises the step)
        self._cg = None                # captured iteration: (CUDAGraph, signature, static inputs, outputs)
        self._eager_direct_steps = 0
        self.launches_per_step = None  # kernels in one captured iteration (library launches + the torch ops of the step)

    # ---------------------------------------------------------------------------------------------------------------------------
    def step(self, events_accu, idx_evt, idx_rgb, blur_target, ts_evt, ts_rgb, H, W, K, K_event, H_ev=None, W_ev=None, remap_evt=None,
             remap_rgb=None):
        """idx_* are THIS rank's pixels; events_accu [H_ev, W_ev] float64; blur_target [R_rgb, C].  Returns (loss, parts)."""
        if self.fused:
            p0, p1 = self.params[0], self.params[-1]
            if p0.data_ptr() != self.flat_params.data_ptr() or p1.data_ptr() != self.flat_params[-p1.numel():].data_ptr():
                raise RuntimeError("a parameter no longer aliases Trainer.flat_params (graph.to() / p.data = ... after Trainer was built)")
        if not self.fused or self.crf:
            return self._step_autograd(events_accu, idx_evt, idx_rgb, blur_target, ts_evt, ts_rgb, H, W, K, K_event, H_ev, W_ev, remap_evt, remap_rgb)
        H_ev, W_ev = H_ev or H, W_ev or W
        a = self.args
        dev = self.flat_params.device
        use_remap = a.dataset == "TUM_VIE"
        ts_e, ts_r = self._linspace(ts_evt, 2), self._linspace(ts_rgb, a.num_interpolated_pose)
        inputs = {"events_accu": events_accu, "idx_evt": idx_evt, "idx_rgb": idx_rgb, "blur_target": blur_target,
                  "ts": torch.cat([ts_e.to(ts_r.device), ts_r]),                        # the 2 event timestamps first, then the exposure's
                  "remap_evt": remap_evt if use_remap else None, "remap_rgb": remap_rgb if use_remap else None}
        consts = (int(H), int(W), int(H_ev), int(W_ev), _as_tuple(K), _as_tuple(K_event))
        if self.use_graph and self.phase_ms is None:
            sig = (consts, tuple((k, None if v is None else (tuple(v.shape), v.dtype)) for k, v in inputs.items()))
            if self._cg is not None and self._cg[1] == sig:
                out = self._replay(inputs)
            elif self._eager_direct_steps >= 2:
                try:
                    out = self._capture(sig, inputs, consts)
                except RuntimeError as e:      # capture refused here (driver / NCCL combination): the same launches, enqueued eagerly
                    import warnings
                    warnings.warn(f"benerf_b200.train.Trainer: CUDA-graph capture of the step failed ({e}); stepping eagerly")
                    self.use_graph, self._cg = False, None
                    torch.cuda.synchronize()
                    out = self._step_direct({k: _dev(v, dev) for k, v in inputs.items()}, consts)
            else:
                out = self._step_direct({k: _dev(v, dev) for k, v in inputs.items()}, consts)
                self._eager_direct_steps += 1
        else:
            out = self._step_direct({k: _dev(v, dev) for k, v in inputs.items()}, consts)
        self._host_bookkeeping()
        loss_out = out
        return loss_out[0], dict(zip(IF.PART_KEYS, (loss_out[1], loss_out[2], loss_out[3], loss_out[4])))

    def _linspace(self, ts2, num):
        """get_pose_evt / get_pose_rgb (model/optimize.py:58-111): linspace(ts[0], ts[1], num); device tensors pass through."""
        if isinstance(ts2, torch.Tensor) and ts2.is_cuda and ts2.numel() == num:
            return ts2
        return torch.linspace(float(ts2[0]), float(ts2[1]), num)

    def _host_bookkeeping(self):
        # lr = lr0 * rate ** (step / (lrate_decay * 1000)) with one rate per optimiser, applied after the step (train.py:355-394);
        # the fused tail evaluates the same expression on the device, these mirrors keep optimizer.param_groups readable
        a = self.args
        decay_steps = getattr(a, "lrate_decay", 200) * 1000
        for o, base, rate in zip(self.optims, self.base_lr, self._rates()):
            for grp, lr0 in zip(o.param_groups, base):
                grp["lr"] = lr0 * rate ** (self.global_step / decay_steps)
        self.global_step += 1

    def _rates(self):
        a = self.args
        return [getattr(a, k, d) for k, d in (("decay_rate", 0.1), ("decay_rate_pose", 0.01), ("decay_rate_transform", 0.01),
                                              ("decay_rate_rgb_crf", 0.1), ("decay_rate_event_crf", 0.1))]

    # ---------------------------------------------------------------------------------------------------------------------------
    def _tables(self, eng):
        """ctypes pointer tables into the flat buffers (they never move): weights for bnrf_set_weights, gradients for
        bnrf_render_backward; built once."""
        if getattr(self, "_tab", None) is None:
            names = [n + sfx for n in LINEAR_NAMES for sfx in (".weight", ".bias")]
            w, gt = [], []
            for m in self.nets:
                table = dict(m.named_parameters())
                w.append({n: table[n] for n in names})
                gt.append(eng._grad_table({n: table[n].grad for n in names}))
            self._tab = (w, gt)
        return self._tab

    def _step_direct(self, t, consts):
        """One iteration as explicit library calls on the current stream; t: device tensors, consts: host constants."""
        g, a = self.graph, self.args
        H, W, H_ev, W_ev, K, K_event = consts
        eng = g.engine(a)
        mark = self._mark
        mark(None)
        weights, grad_tabs = self._tables(eng)
        if getattr(a, "use_barf_c2f", False):
            g._sync_barf(eng, self.global_step, a)
        knots = g.evt_knot_pose_se3.params.weight.data
        transform = g.transform.params.weight.data.reshape(6)
        seed = g.seed()
        n_pe = 2                                                                       # get_pose_evt: window start / end (model/optimize.py:58-82)
        # get_pose_evt + get_pose_rgb, one launch -- on the side stream: a 26 us chain of dependent latency that needs nothing of the
        # four weight-packing launches (another 80 us of small dependent kernels), so the two run side by side
        poses = torch.empty(t["ts"].numel(), 3, 4, device=eng.device, dtype=torch.float32)
        cur = torch.cuda.current_stream()
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            eng.spline_poses_pair(knots, transform, t["ts"], n_pe, a.traj, out=poses)
        if len(weights) == 2:                                            # the parameters changed (bnrf_adam_step_sched wrote them in place)
            eng.set_weights_pair(weights[0], weights[1])
        else:
            eng.set_weights(0, weights[0])
        cur.wait_stream(self._side)
        poses_evt, poses_rgb = poses[:n_pe], poses[n_pe:]
        fine = len(self.nets) > 1
        # the event pose pair and the N blur poses (model/nerf.py:217,227) as two segments of ONE ray batch: every stage of the
        # render and of its backward pass runs once over all 2 R_e + P R_b rays
        segs = [(poses_evt, t["idx_evt"], H_ev, W_ev, K_event, t["remap_evt"]), (poses_rgb, t["idx_rgb"], H, W, K, t["remap_rgb"])]
        n_evt = poses_evt.shape[0] * t["idx_evt"].numel()
        n = n_evt + poses_rgb.shape[0] * t["idx_rgb"].numel()
        saved = self._buffer("saved", eng.saved_bytes(n))
        ret = eng.render_multi(segs, seed=seed, offset=1, saved=saved, offset_dev=self.step_dev)
        mark("forward")
        coarse_key = "rgb0" if fine else "rgb_map"
        loss_out, (d_fine, d_coarse) = training_loss_fused(
            loss_cfg(a), ret["rgb_map"][:n_evt], ret[coarse_key][:n_evt], t["events_accu"], t["idx_evt"], ret["rgb_map"][n_evt:],
            ret[coarse_key][n_evt:], t["blur_target"], all_reduce=IF._rank_sum(), out_like=(ret["rgb_map"], ret[coarse_key]), split=n_evt)
        mark("loss")
        d_poses = torch.zeros(poses_evt.shape[0] + poses_rgb.shape[0], 3, 4, device=eng.device)
        d_pe, d_pr = d_poses[:poses_evt.shape[0]], d_poses[poses_evt.shape[0]:]
        gc, gf = grad_tabs[0], grad_tabs[1] if fine else None
        eng.render_backward_multi(segs, saved, d_fine, d_coarse if fine else None, gc, gf, [d_pe, d_pr])
        # The exchange of the step, overlapped with what is left of the backward pass: the fine network's gradients (its backward
        # pass runs first) are reduced while the coarse network's still runs, the coarse network's while the pose gradients are
        # formed.  NCCL runs on its own stream; the only join is in front of the optimiser.
        pending, n_early = [], 0
        if world() > 1 and fine and getattr(a, "overlap_all_reduce", True):
            import torch.distributed as dist
            n_early = sum(p.numel() for p in self.nets[1].parameters())      # the fine network's gradients lead the flat buffer
            eng.wait_fine_gradients(self._side)
            with torch.cuda.stream(self._side):
                pending.append(dist.all_reduce(self.flat.flat[:n_early], async_op=True))
            torch.cuda.current_stream().wait_stream(self._side)          # (joins the side stream for graph capture; NCCL is not on it)
        gk, gt = g.evt_knot_pose_se3.params.weight.grad, g.transform.params.weight.grad.reshape(6)
        eng.spline_poses_pair_backward(knots, transform, t["ts"], n_pe, d_poses, gk, gt, a.traj)
        mark("backward")
        if world() > 1:
            import torch.distributed as dist
            pending.append(dist.all_reduce(self.flat.flat[n_early:], async_op=True))     # coarse network, knots, transform
            for work in pending:
                work.wait()
        mark("all_reduce")
        flags = [getattr(a, "optimize_nerf", True), getattr(a, "optimize_pose", True), getattr(a, "optimize_trans", False)]
        rates = self._rates()
        groups = [(b, e, self.base_lr[k][0], rates[k], flags[k]) for k, (b, e) in enumerate(self.group_ranges)]
        adam_step_sched(self.flat_params, self.flat.flat, self.exp_avg, self.exp_avg_sq, groups, self.step_dev,
                        getattr(a, "lrate_decay", 200) * 1000, grad_scale=1.0 / world(), zero_grads=True,
                        advance_scratch=self.step_scratch)               # ... and global_step += 1
        eng.invalidate_weights()                                         # for Graph.render callers outside this step (evaluation)
        mark("optimizer")
        return loss_out

    def _buffer(self, name, nbytes):
        bufs = self.__dict__.setdefault("_bufs", {})
        if name not in bufs or bufs[name].numel() < nbytes:
            bufs[name] = torch.empty(nbytes, device=self.flat_params.device, dtype=torch.uint8)
        return bufs[name]

    # ---------------------------------------------------------------------------------------------------------------------------
    def _capture(self, sig, inputs, consts):
        dev = self.flat_params.device
        static = {k: (None if v is None else _dev(v, dev).clone()) for k, v in inputs.items()}
        eng = self.graph.engine(self.args)
        torch.cuda.synchronize()
        before = eng.launch_count()
        cg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cg, capture_error_mode="thread_local"):
            out = self._step_direct(static, consts)
        # kernels launched through the context + the ones that do not go through it: the loss (two stages when the event loss is the
        # normalised one), Adam, torch's fill of d_poses, and the NCCL all-reduce(s) of the gradient buffer / batch norms
        normalised = self.args.event_threshold <= 0
        self.launches_per_step = eng.launch_count() - before + (2 if normalised else 1) + 1 + 1 + (world() > 1) * (2 + normalised)
        self._cg = (cg, sig, static, out, {k: None for k in static})
        return self._replay(inputs)

    def _replay(self, inputs):
        cg, _, static, out, seen = self._cg
        for k, v in inputs.items():
            if v is None:
                continue
            # small host tensors (the two timestamp vectors) are compared by value, device tensors by identity + version
            key = ("v", tuple(v.tolist())) if (not v.is_cuda and v.numel() <= 64) else (v.data_ptr(), v._version, v.device)
            if seen[k] != key:                       # the caller passed new data: refresh the captured iteration's input
                static[k].copy_(v, non_blocking=True)
                seen[k] = key
        cg.replay()
        return out

    # ---------------------------------------------------------------------------------------------------------------------------
    def _step_autograd(self, events_accu, idx_evt, idx_rgb, blur_target, ts_evt, ts_rgb, H, W, K, K_event, H_ev, W_ev, remap_evt, remap_rgb):
        g, a = self.graph, self.args
        mark = self._mark
        mark(None)
        poses_evt = g.get_pose_evt(a, ts_evt)
        poses_rgb = g.get_pose_rgb(a, ts_rgb)
        # same Philox draws as the direct step, which renders both pose sets as one batch: stream offset 64 * global_step + 1, the
        # blur rays keyed behind the event rays
        g._render_calls = 64 * self.global_step
        ret_evt = g.render(self.global_step, poses_evt, idx_evt, H_ev or H, W_ev or W, K_event, a, enable_crf=True, sensor_type="event",
                           remap=remap_evt, training=True)
        g._render_calls = 64 * self.global_step
        ret_rgb = g.render(self.global_step, poses_rgb, idx_rgb, H, W, K, a, enable_crf=True, sensor_type="rgb", remap=remap_rgb, training=True,
                           ray_base=ret_evt["rgb_map"].shape[0])
        mark("forward")
        if getattr(a, "optimize_event_crf", False):                      # train.py:176-185
            ret_evt = {k: g.event_crf.forward(ret_evt[k]) for k in ("rgb_map", "rgb0")}
        if getattr(a, "optimize_rgb_crf", False):                        # train.py:186-192
            ret_rgb = {k: g.rgb_crf.forward(ret_rgb[k]) for k in ("rgb_map", "rgb0")}
        loss, parts = IF.training_loss(ret_evt, ret_rgb, events_accu, idx_evt, blur_target, a)
        for o in self.optims[3:]:
            o.zero_grad()
        if not self.fused:
            self.flat.zero()
        mark("loss")
        loss.backward()
        mark("backward")
        flags = [getattr(a, "optimize_nerf", True), getattr(a, "optimize_pose", True), getattr(a, "optimize_trans", False),
                 getattr(a, "optimize_rgb_crf", False), getattr(a, "optimize_event_crf", False)]
        if self.fused:
            self.flat.all_reduce_sum()
            mark("all_reduce")
            rates = self._rates()
            groups = [(b, e, self.base_lr[k][0], rates[k], flags[k]) for k, (b, e) in enumerate(self.group_ranges)]
            adam_step_sched(self.flat_params, self.flat.flat, self.exp_avg, self.exp_avg_sq, groups, self.step_dev,
                            getattr(a, "lrate_decay", 200) * 1000, grad_scale=1.0 / world(), zero_grads=True,
                            advance_scratch=self.step_scratch)
            g.engine(a).invalidate_weights()                             # in-place update torch's version counters do not see
        else:
            self.flat.all_reduce_mean()
            mark("all_reduce")
            for o, f in zip(self.optims[:3], flags[:3]):
                if f:
                    o.step()
        for o, f in zip(self.optims[3:], flags[3:]):                     # tone-mapper optimisers (train.py:349-352); their few hundred
            if f:                                                        # parameters are averaged over the ranks like the rest
                if world() > 1:
                    import torch.distributed as dist
                    for grp in o.param_groups:
                        for p in grp["params"]:
                            if p.grad is not None:
                                dist.all_reduce(p.grad)
                                p.grad.div_(world())
                o.step()
        mark("optimizer")
        self._host_bookkeeping()
        return loss.detach(), parts

    # ---------------------------------------------------------------------------------------------------------------------------
    # Reference-format optimiser state (train.py:443-455 saves optimizer.state_dict() x5, test.py:102-106 reloads them): the fused
    # tail keeps the Adam moments in flat buffers, these two calls mirror them into / out of the torch.optim.Adam objects.
    def export_optimizer_state(self):
        """Flat moments + step count -> optimizer.state of the three Adam objects (so optimizer.state_dict() is a checkpoint)."""
        if not self.fused:
            return
        step = float(self.global_step)
        off = 0
        owners = [self.optims[0]] * (len(self.params) - 2) + [self.optims[1], self.optims[2]]
        for p, o in zip(self.params, owners):
            n = p.numel()
            o.state[p] = {"step": torch.tensor(step), "exp_avg": self.exp_avg[off:off + n].view_as(p).clone(),
                          "exp_avg_sq": self.exp_avg_sq[off:off + n].view_as(p).clone()}
            off += n

    def import_optimizer_state(self):
        """optimizer.state (e.g. after optimizer.load_state_dict of a reference checkpoint) -> flat moments + step count."""
        if not self.fused:
            return
        off, step = 0, None
        owners = [self.optims[0]] * (len(self.params) - 2) + [self.optims[1], self.optims[2]]
        for p, o in zip(self.params, owners):
            n = p.numel()
            st = o.state.get(p)
            if st:
                self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
                self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                step = int(st["step"]) if step is None else step
            off += n
        if step is not None:
            self.global_step = step
            self.step_dev.fill_(step)

    def state_dict(self):
        """The optimiser part of the reference's checkpoint (train.py:447-452) + the step counter."""
        self.export_optimizer_state()
        keys = ("optimizer_nerf", "optimizer_pose", "optimizer_trans", "optimizer_rgb_crf", "optimizer_event_crf")
        sd = {k: o.state_dict() for k, o in zip(keys, self.optims)}
        sd["global_step"] = self.global_step
        return sd

    def load_state_dict(self, sd):
        keys = ("optimizer_nerf", "optimizer_pose", "optimizer_trans", "optimizer_rgb_crf", "optimizer_event_crf")
        for k, o in zip(keys, self.optims):
            if k in sd:
                o.load_state_dict(sd[k])
        self.import_optimizer_state()
        if "global_step" in sd:
            self.global_step = int(sd["global_step"])
            if self.fused:
                self.step_dev.fill_(self.global_step)

    def _mark(self, name):
        if self.phase_ms is None:
            return
        import time
        torch.cuda.synchronize()
        now = time.perf_counter()
        if name is not None:
            self.phase_ms.append((self.global_step, name, round((now - self._t_prev) * 1e3, 2)))
        self._t_prev = now


def _dev(v, dev):
    if v is None:
        return None
    if not isinstance(v, torch.Tensor):
        v = torch.as_tensor(v)
    return v.to(device=dev).contiguous()


def _as_tuple(K):
    import numpy as np
    return tuple(float(x) for x in np.asarray(K.cpu() if isinstance(K, torch.Tensor) else K, dtype=np.float32).reshape(-1))
