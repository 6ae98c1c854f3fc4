"""Mirror of model/nerf.py: NeRF (parameter holder), Graph (forward / render / render_video).

The modules own the same parameters under the same names as the reference, so state dicts,
optimisers and init_nerf work unchanged; the arithmetic of NeRF.forward / raw2output /
Graph.render is not here but in libbenerf_b200.so (bnrf_render_forward).
"""
import abc
from datetime import datetime

import numpy as np
import torch
from torch import nn

from .engine import Engine
from . import image_formation


from ._lib import LINEAR_NAMES

_OUT_KEYS = ("rgb_map", "disp_map", "acc_map", "rgb0", "disp0", "acc0", "sigma")


def barf_c2f_channel_weights(iter_step, args):
    """model/nerf.py:16-26 as per-channel factors of the 63-channel point and 27-channel direction encodings fed to the network
    (cat([x, barf_c2f_weight(PE(x))]), model/nerf.py:75-88): ((63 floats), (27 floats)).  Upstream multiplies the [M, 6L] encoding
    viewed as (-1, L) by the L ramp values, so channel e of the sin/cos part takes weight[e % L], and the raw input 1."""
    out = []
    for L in (args.multires, args.multires_views):
        progress = iter_step / args.max_iter
        alpha = (progress - args.barf_c2f_start) / (args.barf_c2f_end - args.barf_c2f_start) * L
        k = torch.arange(L)
        weight = (1 - (alpha - k).clamp_(min=0, max=1).mul_(np.pi).cos_()) / 2          # same float32 ops as upstream
        out.append(tuple([1.0, 1.0, 1.0] + [float(weight[e % L]) for e in range(6 * L)]))
    return tuple(out)


class _RenderFn(torch.autograd.Function):
    """Graph.render under autograd: forward = bnrf_render_forward_train, backward = bnrf_render_backward.

    Differentiable inputs: poses [P,3,4] and the 24 (+24) NeRF parameters; differentiable outputs: rgb_map and rgb0
    (the only outputs train.py:205-331 puts into a loss).  Everything else is marked non-differentiable.
    """

    @staticmethod
    def forward(ctx, call, poses, *params):
        eng, ray_idx, H, W, K, remap, rng, seed, offset, n_fine, ray_base = call
        n = poses.shape[0] * ray_idx.numel()
        saved = torch.empty(eng.saved_bytes(n), device=eng.device, dtype=torch.uint8)
        ret = eng.render(poses, ray_idx, H, W, K, remap=remap, rng=rng, seed=seed, offset=offset, saved=saved, ray_base=ray_base)
        keys = [k for k in _OUT_KEYS if k in ret]
        ctx.call, ctx.keys, ctx.saved_buf, ctx.poses = call, keys, saved, poses
        ctx.params = params
        outs = tuple(ret[k] for k in keys)
        ctx.mark_non_differentiable(*[o for k, o in zip(keys, outs) if k not in ("rgb_map", "rgb0")])
        return outs

    @staticmethod
    def backward(ctx, *gouts):
        eng, ray_idx, H, W, K, remap, rng, seed, offset, n_fine, ray_base = ctx.call
        g = {k: (go.contiguous() if go is not None else None) for k, go in zip(ctx.keys, gouts)}
        # bnrf_render_backward ACCUMULATES into the gradient tables.  When every parameter already owns a gradient buffer
        # (benerf_b200.parallel.FlatGrads: p.grad are views of one flat fp32 buffer) the kernels add straight into p.grad
        # and autograd is told there is nothing left to accumulate: no temporaries, no 48 add kernels per render.
        if ctx.saved_buf is None:
            raise RuntimeError("Graph.render's saved activations were released by the first backward pass (they are ~10 KB per "
                               "sample); render again instead of backward(retain_graph=True)")
        # direct accumulation only on request (parallel.FlatGrads sets _bnrf_direct_grad on the parameters it lays out): otherwise
        # torch.autograd.grad(loss, params) would silently get no gradient while .grad is mutated
        direct = all(getattr(p, "_bnrf_direct_grad", False) and p.grad is not None and p.grad.is_contiguous()
                     and p.grad.dtype == torch.float32 and p.grad.device == eng.device for p in ctx.params)
        if direct:
            grads = [p.grad for p in ctx.params]
        else:
            flat = torch.zeros(sum(p.numel() for p in ctx.params), device=eng.device, dtype=torch.float32)
            grads, off = [], 0
            for p in ctx.params:
                grads.append(flat[off:off + p.numel()].view(p.shape))
                off += p.numel()
        names = [n + sfx for n in LINEAR_NAMES for sfx in (".weight", ".bias")]
        gc = dict(zip(names, grads[:24]))
        gf = dict(zip(names, grads[24:48])) if n_fine else None
        d_poses = torch.zeros_like(ctx.poses)
        eng.render_backward(ctx.poses, ray_idx, H, W, K, ctx.saved_buf, g.get("rgb_map"), g.get("rgb0"), gc, gf, d_poses, remap=remap)
        ctx.saved_buf = None
        return (None, d_poses) + (tuple(None for _ in grads) if direct else tuple(grads))


class Model:
    @abc.abstractmethod
    def build_network(self, args, poses=None, event_poses=None):
        pass

    @abc.abstractmethod
    def setup_optimizer(self, args):
        pass

    def after_train(self):
        print(f"Successfully finished model on {datetime.now()}")


class NeRF(nn.Module):
    """model/nerf.py:40-64: the 12 linears.  forward()/raw2output() live in the CUDA library."""

    def __init__(self, D=8, W=256, input_ch=63, input_ch_views=27, output_ch=4, skips=[4], use_viewdirs=False, channels=3):
        super().__init__()
        if (D, W, input_ch, input_ch_views, list(skips), use_viewdirs) != (8, 256, 63, 27, [4], True):
            raise ValueError("benerf_b200 implements the architecture Graph hard-codes (model/optimize.py:9): "
                             "D=8, W=256, input_ch=63, input_ch_views=27, skips=[4], use_viewdirs=True")
        self.D, self.W, self.input_ch, self.input_ch_views = D, W, input_ch, input_ch_views
        self.skips, self.use_viewdirs, self.channels = skips, use_viewdirs, channels
        self.pts_linears = nn.ModuleList(
            [nn.Linear(input_ch, W)] + [nn.Linear(W, W) if i not in skips else nn.Linear(W + input_ch, W) for i in range(D - 1)])
        self.views_linears = nn.ModuleList([nn.Linear(input_ch_views + W, W // 2)])
        self.feature_linear = nn.Linear(W, W)
        self.alpha_linear = nn.Linear(W, 1)
        self.rgb_linear = nn.Linear(W // 2, channels)

    def forward(self, *a, **k):
        raise RuntimeError("NeRF.forward is fused into bnrf_render_forward; call Graph.render")


class Graph(nn.Module):
    def __init__(self, args, D=8, W=256, input_ch=63, input_ch_views=27, output_ch=4, skips=[4], use_viewdirs=False):
        super().__init__()
        self.nerf = NeRF(D, W, input_ch, input_ch_views, output_ch, skips, use_viewdirs, args.channels)
        self.channels = args.channels
        if args.N_importance > 0:
            self.nerf_fine = NeRF(D, W, input_ch, input_ch_views, output_ch, skips, use_viewdirs, args.channels)
        self.pose_eye = torch.eye(3, 4)
        self._engine = None
        self._render_calls = 0
        self._seed = int(getattr(args, "seed", 0) or 0)
        self._events_dev = None

    # ------------------------------------------------------------------------------------
    def engine(self, args):
        if self._engine is None:
            if not getattr(args, "ndc", True) or not args.use_viewdirs:
                raise ValueError("benerf_b200 covers the configuration every shipped config uses: ndc=True, use_viewdirs=True")
            self._engine = Engine(n_samples=args.N_samples, n_importance=args.N_importance, channels=args.channels,
                                  mlp_mode=getattr(args, "mlp_mode", "tc"), gemm_mode=getattr(args, "gemm_mode", "tc"))
        return self._engine

    def seed(self):
        """Philox key of this process: args.seed with the data-parallel rank folded into the upper half, so that pixel shards
        draw independent jitter / density noise / inverse-CDF samples, as the reference's draws over a full batch are."""
        from .parallel import rank
        return self._seed ^ (rank() << 32)

    def _render_params(self, nets):
        """The 24 (+24) parameters in bnrf_set_weights order; the Parameter objects live as long as the modules, so the
        list is built once (named_parameters() walks the module tree: ~60 % of a training step's host time if redone)."""
        if getattr(self, "_param_cache", None) is None or self._param_cache[0] != len(nets):
            ps = []
            for m in nets:
                table = dict(m.named_parameters())
                ps += [table[n + sfx] for n in LINEAR_NAMES for sfx in (".weight", ".bias")]
            self._param_cache = (len(nets), ps)
        return self._param_cache[1]

    def _sync_barf(self, eng, iter_step, args):
        w = barf_c2f_channel_weights(iter_step, args)
        if w != getattr(self, "_barf_weights", None):
            eng.set_encoding_weights(*w)
            self._barf_weights = w

    def _sync(self, eng, iter_step=0, args=None):
        # BARF coarse-to-fine (args.use_barf_c2f, model/nerf.py:16-26,75-88): the per-channel encoding weights of this iter_step
        # are folded into the packed weight matrices, so a change of the weights forces a repack
        w = barf_c2f_channel_weights(iter_step, args) if args is not None and getattr(args, "use_barf_c2f", False) else None
        if w != getattr(self, "_barf_weights", None):
            eng.set_encoding_weights(*(w if w is not None else (None, None)))
            self._barf_weights = w
        eng.sync_weights(0, self.nerf)
        if hasattr(self, "nerf_fine"):
            eng.sync_weights(1, self.nerf_fine)

    # ------------------------------------------------------------------------------------
    def forward(self, iter_step, events, rgb_exp_ts, H, W, K, K_event, args, img_xy_remap, evt_xy_remap):
        """model/nerf.py:160-234: pick an event window, accumulate it, interpolate the two pose sets,
        render the event pair and the N-pose blur batch."""
        dev = self.engine(args).device
        ev = self._device_events(events, dev, args)
        ts_all = events["ts"]
        if args.event_time_window:
            window_t = args.accumulate_time_length
            if args.random_sampling_window:
                low_t = np.random.rand(1) * (1 - window_t)
                upper_t = low_t + window_t
            else:
                low_t = np.random.randint((1 - window_t) // window_t) * window_t
                upper_t = np.min((low_t + window_t, 1.0))
            # the reference masks low <= ts <= up over all events (order-independent, model/nerf.py:170-178); on time-sorted
            # events -- checked once at upload, a sorted device copy is made otherwise -- that closed window (Q16) is one slice
            lo = int(np.searchsorted(ev["ts_sorted"], float(np.asarray(low_t).reshape(-1)[0]), side="left"))
            hi = int(np.searchsorted(ev["ts_sorted"], float(np.asarray(upper_t).reshape(-1)[0]), side="right"))
            events_ts = np.stack((low_t, upper_t)).reshape(2)
            win = ev["sorted"]
        else:
            num = len(events["pol"])
            N_window = round(num * args.accumulate_time_length)
            if args.random_sampling_window:
                lo = np.random.randint(num - N_window)
            else:
                lo = np.random.randint((num - N_window) // N_window) * N_window
            hi = int(lo + N_window)
            events_ts = ts_all[np.array([lo, hi - 1])]
            win = ev["given"]                           # an index window is a slice of the events in the order given (model/nerf.py:190-193)
        events_accu = image_formation.accumulate_events(win["x"][lo:hi], win["y"][lo:hi], win["pol"][lo:hi],
                                                        args.event_height, args.event_width)
        spline_evt_poses = self.get_pose_evt(args, torch.tensor(events_ts, dtype=torch.float32))
        spline_rgb_poses = self.get_pose_rgb(args, torch.tensor(rgb_exp_ts, dtype=torch.float32))
        ray_idx_event = torch.randperm(args.event_height * args.event_width, device=dev)[:args.sampling_event_rays]
        ret_event = self.render(iter_step, spline_evt_poses, ray_idx_event, args.event_height, args.event_width,
                                K_event, args, enable_crf=True, sensor_type="event", remap=evt_xy_remap, training=True)
        ray_idx_rgb = torch.randperm(H * W, device=dev)[:args.sampling_rgb_rays // args.num_interpolated_pose]
        ret_rgb = self.render(iter_step, spline_rgb_poses, ray_idx_rgb, H, W, K, args, enable_crf=True,
                              sensor_type="rgb", remap=img_xy_remap, training=True)
        return ret_event, ret_rgb, ray_idx_event, ray_idx_rgb, events_accu

    def _device_events(self, events, dev, args):
        """Keep the event arrays resident on the device instead of masking + uploading the window every iteration
        (model/nerf.py:162-199).  Returns {"given": arrays in the caller's order, "sorted": arrays in time order, "ts_sorted"}; for
        time-sorted input (verified here, once per events object) the two are the same tensors."""
        key = id(events["ts"])
        if self._events_dev is None or self._events_dev[0] != key:
            pol = np.asarray(events["pol"], dtype=np.float32).copy()
            if args.dataset == "TUM_VIE":
                pol[pol == 0] = -1                      # 0 = negative polarity in TUM-VIE (model/nerf.py:194-196)
            ts = np.asarray(events["ts"])

            def upload(order=None):
                pick = (lambda a: a) if order is None else (lambda a: a[order])
                return {"x": torch.as_tensor(pick(np.asarray(events["x"])), dtype=torch.int32, device=dev),
                        "y": torch.as_tensor(pick(np.asarray(events["y"])), dtype=torch.int32, device=dev),
                        "pol": torch.as_tensor(pick(pol), device=dev)}
            given = upload()
            if ts.size < 2 or bool(np.all(ts[1:] >= ts[:-1])):
                state = {"given": given, "sorted": given, "ts_sorted": ts}
            else:
                order = np.argsort(ts, kind="stable")
                state = {"given": given, "sorted": upload(order), "ts_sorted": ts[order]}
            self._events_dev = (key, state)
        return self._events_dev[1]

    # ------------------------------------------------------------------------------------
    def render(self, iter_step, poses, ray_idx, H, W, K, args, enable_crf: bool, sensor_type: str, remap,
               near=0., far=1., training=False, rng=None, ray_base=0):
        """model/nerf.py:236-343.  Returns {'rgb_map','disp_map','acc_map'} (+ 'rgb0','disp0','acc0','sigma'
        when N_importance > 0), pose-major.  training/eval produce identical rays upstream and
        share one path here.  enable_crf / sensor_type are accepted and ignored exactly as upstream
        (model/nerf.py:127-131).  Extensions (keyword-only): rng = dict of the four draws for parity runs; ray_base = index of this
        render's first ray in a larger batch (its Philox draws then equal those of Engine.render_multi for the same rows)."""
        eng = self.engine(args)
        if (near, far) != (0., 1.):
            raise ValueError("Graph.render is only ever called with near=0, far=1 upstream (model/nerf.py:239)")
        self._sync(eng, iter_step, args)
        dev = eng.device
        poses = poses[:, :3, :4].to(device=dev, dtype=torch.float32).contiguous()
        ray_idx = torch.as_tensor(ray_idx).reshape(-1).to(device=dev, dtype=torch.int64).contiguous()
        use_remap = args.dataset == "TUM_VIE" and remap is not None
        remap_t = torch.as_tensor(remap, dtype=torch.float32, device=dev).contiguous() if use_remap else None
        self._render_calls += 1
        K_np = np.asarray(K.cpu() if isinstance(K, torch.Tensor) else K, dtype=np.float32)
        nets = [self.nerf] + ([self.nerf_fine] if hasattr(self, "nerf_fine") else [])
        params = self._render_params(nets)
        if torch.is_grad_enabled() and (poses.requires_grad or any(p.requires_grad for p in params)):
            call = (eng, ray_idx, H, W, K_np, remap_t, rng, self.seed(), self._render_calls, len(nets) > 1, int(ray_base))
            outs = _RenderFn.apply(call, poses, *params)
            return dict(zip([k for k in _OUT_KEYS if len(nets) > 1 or k in ("rgb_map", "disp_map", "acc_map")], outs))
        return eng.render(poses.detach(), ray_idx, H, W, K_np, remap=remap_t, rng=rng, seed=self.seed(), offset=self._render_calls,
                          ray_base=ray_base)

    @torch.no_grad()
    def render_video(self, iter_step, poses, H, W, K, args, remap, type):
        """model/nerf.py:353-390: full-image render, chunked by args.chunk rays; returns [H,W,...] maps."""
        all_ret = {}
        ray_idx = torch.arange(0, H * W)
        if str(type) not in ("radience", "rgb"):
            raise ValueError(type)
        # the reference walks the image in args.chunk-ray pieces to bound its activation memory (model/nerf.py:360); here a
        # ray costs 4 KB of scratch, so the whole image is one launch sequence (chunks of 4 M rays for very large frames)
        chunk = max(int(args.chunk), int(getattr(args, "render_chunk_rays", 1 << 22)))
        for i in range(0, ray_idx.shape[0], chunk):
            ret = self.render(iter_step, poses, ray_idx[i:i + chunk], H, W, K, args, enable_crf=(str(type) == "rgb"),
                              sensor_type=("rgb" if str(type) == "rgb" else None), remap=remap, training=False)
            for k in ret:
                all_ret.setdefault(k, []).append(ret[k])
        for k in all_ret:
            all_ret[k] = torch.cat(all_ret[k], 0).reshape([H, W] + list(all_ret[k][0].shape[1:]))
        return all_ret

    @abc.abstractmethod
    def get_pose_evt(self, args, events_ts, seg_num=None):
        pass

    @abc.abstractmethod
    def get_pose_rgb(self, args, exposure_ts, seg_num=None):
        pass
