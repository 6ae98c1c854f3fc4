"""Thin host wrapper over the C ABI: owns one bnrf_ctx, passes raw device pointers.

PyTorch is used for device memory and streams only; every number on the render path is
produced by libbenerf_b200.so.  All calls are enqueued on torch's current CUDA stream.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import BnrfError

TRAJ = {"spline": 0, "linear": 1}
LOG_MODE = {"BeNeRF_Blender": 0, "BeNeRF_Unreal": 0, "E2NeRF_Synthetic": 1, "E2NeRF_Real": 1, "safelog": 0, "linlog": 1}
MLP_MODES = {"tc": _lib.MLP_TC_FP16X2, "simt": _lib.MLP_SIMT_FP32, "tc1": _lib.MLP_TC_1CTA, "tc2": _lib.MLP_TC_PAIR_SS}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t, dtype=torch.float32, device=None, name="tensor"):
    if t is None:
        return None
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise BnrfError(f"{name}: expected a CUDA tensor")
    if t.dtype != dtype or not t.is_contiguous():
        raise BnrfError(f"{name}: expected contiguous {dtype}, got {t.dtype} contiguous={t.is_contiguous()}")
    if device is not None and t.device != device:
        raise BnrfError(f"{name}: tensor on {t.device}, engine on {device}")
    return C.c_void_p(t.data_ptr())


class Engine:
    """One context per (process, device) -- SURVEY 8-b conventions."""

    def __init__(self, n_samples=64, n_importance=64, channels=3, ndc=True, near=0.0, far=1.0, mlp_mode="tc", device=None, gemm_mode="tc"):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise BnrfError("benerf_b200 needs a CUDA (sm_100a) device; there is no CPU path")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else torch.device(device).index or 0)
        self.cfg = _lib.Cfg(n_samples, n_importance, channels, int(bool(ndc)), near, far, MLP_MODES[mlp_mode], {"tc": 0, "simt": 1, "tc_linear": 2, "tc_chain1": 3}[gemm_mode])
        self.n_samples, self.n_importance, self.channels = n_samples, n_importance, channels
        self.mlp_mode, self.gemm_mode = mlp_mode, gemm_mode
        self._ctx = C.c_void_p()
        rc = self.lib.bnrf_create(C.byref(self._ctx), self.device.index, C.byref(self.cfg))
        if rc != _lib.OK:
            raise BnrfError(f"bnrf_create failed ({rc}): {self.lib.bnrf_last_error(None).decode()}")
        self._workspace = None
        self._weights_version = [None, None]

    # -- lifetime -----------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self.lib.bnrf_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != _lib.OK:
            raise BnrfError(f"{what} failed ({rc}): {self.lib.bnrf_last_error(self._ctx).decode()}")

    # -- parameters ---------------------------------------------------------------------
    def set_weights(self, net, params):
        """params: {'pts_linears.0.weight': tensor, ...} (reference state-dict names) or an nn.Module."""
        if hasattr(params, "state_dict"):
            params = dict(params.named_parameters())
        ws, bs, keep = (C.c_void_p * 12)(), (C.c_void_p * 12)(), []
        for i, name in enumerate(_lib.LINEAR_NAMES):
            w = params[name + ".weight"].detach()
            b = params[name + ".bias"].detach()
            keep += [w, b]
            ws[i] = _ptr(w, device=self.device, name=name + ".weight")
            bs[i] = _ptr(b, device=self.device, name=name + ".bias")
        self._check(self.lib.bnrf_set_weights(self._ctx, int(net), ws, bs, _stream()), "bnrf_set_weights")

    def _weight_tables(self, params):
        ws, bs, keep = (C.c_void_p * 12)(), (C.c_void_p * 12)(), []
        for i, name in enumerate(_lib.LINEAR_NAMES):
            w, b = params[name + ".weight"].detach(), params[name + ".bias"].detach()
            keep += [w, b]
            ws[i] = _ptr(w, device=self.device, name=name + ".weight")
            bs[i] = _ptr(b, device=self.device, name=name + ".bias")
        return ws, bs, keep

    def set_weights_pair(self, params_coarse, params_fine):
        """Both networks in one repack (bnrf_set_weights_pair)."""
        wc, bc, k0 = self._weight_tables(params_coarse)
        wf, bf, k1 = self._weight_tables(params_fine)
        self._check(self.lib.bnrf_set_weights_pair(self._ctx, wc, bc, wf, bf, _stream()), "bnrf_set_weights_pair")

    def sync_weights(self, net, module):
        """Repack only when an optimiser step (or load_state_dict) touched the module's parameters."""
        cache = self.__dict__.setdefault("_sync_params", {})
        if cache.get(net, (None,))[0] is not module:
            table = dict(module.named_parameters())          # walked once per module, not once per render
            cache[net] = (module, {n + sfx: table[n + sfx] for n in _lib.LINEAR_NAMES for sfx in (".weight", ".bias")})
        params = cache[net][1]
        version = tuple((p.data_ptr(), p._version) for p in params.values())
        if version != self._weights_version[net]:
            self.set_weights(net, params)
            self._weights_version[net] = version

    def invalidate_weights(self):
        """Force a repack at the next render: for parameter updates torch cannot see (bnrf_adam_step writes in place)."""
        self._weights_version = [None, None]

    def set_encoding_weights(self, w_pts=None, w_dir=None):
        """BARF c2f channel weights (bnrf_set_encoding_weights): w_pts [63], w_dir [27] host sequences, or None / None to switch the
        weighting off.  The packed weights depend on them: the next sync_weights / set_weights repacks."""
        if w_pts is None:
            self._check(self.lib.bnrf_set_encoding_weights(self._ctx, None, None, _stream()), "bnrf_set_encoding_weights")
        else:
            a = (C.c_float * 63)(*[float(v) for v in w_pts])
            b = (C.c_float * 27)(*[float(v) for v in w_dir])
            self._check(self.lib.bnrf_set_encoding_weights(self._ctx, a, b, _stream()), "bnrf_set_encoding_weights")
        self.invalidate_weights()

    def set_sample_grid(self, t_vals):
        t = torch.as_tensor(t_vals, dtype=torch.float32).cpu().contiguous()
        arr = (C.c_float * t.numel())(*t.tolist())
        self._check(self.lib.bnrf_set_sample_grid(self._ctx, arr, t.numel(), _stream()), "bnrf_set_sample_grid")

    # -- measurement hooks --------------------------------------------------------------
    def profile(self, enable=True):
        self._check(self.lib.bnrf_profile(self._ctx, int(enable)), "bnrf_profile")

    def profile_read(self):
        ms, timed, flops, launches = C.c_double(), C.c_int64(), C.c_double(), C.c_int64()
        self._check(self.lib.bnrf_profile_read(self._ctx, C.byref(ms), C.byref(timed), C.byref(flops), C.byref(launches)), "bnrf_profile_read")
        return {"mlp_ms": ms.value, "mlp_timed": timed.value, "mlp_flops": flops.value, "launches": launches.value}

    def launch_count(self):
        """Kernel launches issued through this context since the last profile(True) (or since it was created)."""
        return self.profile_read()["launches"]

    def mlp_trace(self, enable=True):
        """Debug: per-CTA stall counters of the tensor-core MLP kernel (bnrf_debug_mlp_trace)."""
        self._trace = torch.zeros(148 * 2, 16, device=self.device, dtype=torch.int64) if enable else None
        self._check(self.lib.bnrf_debug_mlp_trace(self._ctx, C.c_void_p(self._trace.data_ptr()) if enable else None), "bnrf_debug_mlp_trace")
        return self._trace

    # -- a1/a2 ----------------------------------------------------------------------------
    def spline_poses(self, knots, transform, ts, traj="spline"):
        P = ts.numel()
        out = torch.empty(P, 3, 4, device=self.device, dtype=torch.float32)
        self._check(self.lib.bnrf_spline_poses(self._ctx, _ptr(knots, name="knots"), _ptr(transform, name="transform"),
                                               _ptr(ts, name="ts"), P, TRAJ[traj], _ptr(out), _stream()), "bnrf_spline_poses")
        return out

    def spline_poses_pair(self, knots, transform, ts, n_plain, traj="spline", out=None):
        """Event poses (first n_plain timestamps, knots only) and RGB poses (the rest, knots + transform) in one launch."""
        P = ts.numel()
        if out is None:
            out = torch.empty(P, 3, 4, device=self.device, dtype=torch.float32)
        self._check(self.lib.bnrf_spline_poses_pair(self._ctx, _ptr(knots, name="knots"), _ptr(transform, name="transform"), _ptr(ts, name="ts"),
                                                    P, int(n_plain), TRAJ[traj], _ptr(out), _stream()), "bnrf_spline_poses_pair")
        return out

    def spline_poses_pair_backward(self, knots, transform, ts, n_plain, d_poses, d_knots, d_transform, traj="spline"):
        self._check(self.lib.bnrf_spline_poses_pair_backward(
            self._ctx, _ptr(knots, name="knots"), _ptr(transform, name="transform"), _ptr(ts, name="ts"), ts.numel(), int(n_plain), TRAJ[traj],
            _ptr(d_poses, name="d_poses"), _ptr(d_knots), _ptr(d_transform), _stream()), "bnrf_spline_poses_pair_backward")

    # -- a3-a10 -----------------------------------------------------------------------------
    def _K(self, K):
        flat = [float(v) for v in torch.as_tensor(K, dtype=torch.float32).reshape(-1).tolist()]
        return (C.c_float * 9)(*flat)

    def render(self, poses, ray_idx, H, W, K, remap=None, rng=None, seed=0, offset=0, want_sigma=True, want_depth=False, want_z=False,
               saved=None, offset_dev=None, ray_base=0):
        """Graph.render (model/nerf.py:236-343).  rng: dict of the four draws (parity mode) or None (Philox).
        saved: uint8 device tensor of saved_bytes(N) bytes -> training mode (bnrf_render_forward_train)."""
        P, R = poses.shape[0], ray_idx.numel()
        n = P * R
        Sf = self.n_samples + self.n_importance
        C_ = self.channels
        new = lambda *s: torch.empty(*s, device=self.device, dtype=torch.float32)
        ret = {"rgb_map": new(n, C_), "disp_map": new(n), "acc_map": new(n)}
        if self.n_importance > 0:
            ret.update({"rgb0": new(n, C_), "disp0": new(n), "acc0": new(n)})
            if want_sigma:
                ret["sigma"] = new(n, Sf)
        depth = new(n) if want_depth else None
        z_vals = new(n, Sf) if want_z else None
        outs = _lib.Outputs(*[_ptr(ret.get(k)) for k in ("rgb_map", "disp_map", "acc_map", "rgb0", "disp0", "acc0", "sigma")],
                            _ptr(depth), _ptr(z_vals))
        r = _lib.Rng(None, None, None, None, None, int(seed), int(offset), int(ray_base), _ptr(offset_dev, torch.int64, name="offset_dev"))
        if rng is not None:
            r.t_rand, r.noise_c = _ptr(rng["t_rand"], name="t_rand"), _ptr(rng["noise_c"], name="noise_c")
            if self.n_importance > 0:
                r.u, r.noise_f = _ptr(rng["u"], name="u"), _ptr(rng["noise_f"], name="noise_f")
                r.z_fine = _ptr(rng.get("z_fine"), name="z_fine")        # parity-only override, see benerf_b200.h
        need = self.lib.bnrf_workspace_bytes(self._ctx, n)
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = torch.empty(need, device=self.device, dtype=torch.uint8)
        common = (self._ctx, _ptr(poses, name="poses"), _ptr(ray_idx, torch.int64, name="ray_idx"), P, R, int(H), int(W),
                  self._K(K), _ptr(remap, name="remap"), C.byref(r), C.byref(outs), C.c_void_p(self._workspace.data_ptr()),
                  self._workspace.numel())
        if saved is None:
            self._check(self.lib.bnrf_render_forward(*common, _stream()), "bnrf_render_forward")
        else:
            self._check(self.lib.bnrf_render_forward_train(*common, _ptr(saved, torch.uint8, name="saved"), saved.numel(), _stream()),
                        "bnrf_render_forward_train")
        if want_depth:
            ret["depth_map"] = depth
        if want_z:
            ret["z_vals"] = z_vals
        return ret

    def _segs(self, segs):
        """segs: [(poses [P,3,4], ray_idx [R], H, W, K, remap or None), ...] -> (bnrf_render_seg array, total rays)."""
        arr = (_lib.RenderSeg * len(segs))()
        n = 0
        for a, (poses, ray_idx, H, W, K, remap) in zip(arr, segs):
            a.poses, a.ray_idx = _ptr(poses, name="poses").value, _ptr(ray_idx, torch.int64, name="ray_idx").value
            a.P, a.R, a.H, a.W = poses.shape[0], ray_idx.numel(), int(H), int(W)
            a.K = self._K(K)
            a.remap = _ptr(remap, name="remap").value if remap is not None else None
            n += a.P * a.R
        return arr, n

    def render_multi(self, segs, seed=0, offset=0, saved=None, offset_dev=None, want_sigma=False):
        """Several Graph.render calls as ONE ray batch (bnrf_render_forward_multi): outputs cover the concatenated rays of the
        segments (segment 0 first, each pose-major).  saved: uint8 device tensor of saved_bytes(N) bytes -> training mode."""
        arr, n = self._segs(segs)
        Sf = self.n_samples + self.n_importance
        new = lambda *s: torch.empty(*s, device=self.device, dtype=torch.float32)
        ret = {"rgb_map": new(n, self.channels), "disp_map": new(n), "acc_map": new(n)}
        if self.n_importance > 0:
            ret.update({"rgb0": new(n, self.channels), "disp0": new(n), "acc0": new(n)})
            if want_sigma:
                ret["sigma"] = new(n, Sf)
        outs = _lib.Outputs(*[_ptr(ret.get(k)) for k in ("rgb_map", "disp_map", "acc_map", "rgb0", "disp0", "acc0", "sigma")], None, None)
        r = _lib.Rng(None, None, None, None, None, int(seed), int(offset), 0, _ptr(offset_dev, torch.int64, name="offset_dev"))
        need = self.lib.bnrf_workspace_bytes(self._ctx, n)
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = torch.empty(need, device=self.device, dtype=torch.uint8)
        self._check(self.lib.bnrf_render_forward_multi(
            self._ctx, arr, len(segs), C.byref(r), C.byref(outs), C.c_void_p(self._workspace.data_ptr()), self._workspace.numel(),
            _ptr(saved, torch.uint8, name="saved"), saved.numel() if saved is not None else 0, _stream()), "bnrf_render_forward_multi")
        return ret

    def wait_fine_gradients(self, stream):
        """bnrf_wait_fine_gradients: `stream` waits until the fine network's parameter gradients of the last render backward pass
        enqueued on this engine are complete (the coarse network's backward pass may still be running)."""
        self._check(self.lib.bnrf_wait_fine_gradients(self._ctx, C.c_void_p(stream.cuda_stream)), "bnrf_wait_fine_gradients")

    def render_backward_multi(self, segs, saved, d_rgb_map, d_rgb0, grads_coarse, grads_fine, d_poses):
        """bnrf_render_backward_multi: d_rgb_map / d_rgb0 over the concatenated rays; d_poses: one [P_i,3,4] tensor per segment."""
        arr, n = self._segs(segs)
        need = self.lib.bnrf_backward_workspace_bytes(self._ctx, n)
        if getattr(self, "_bwd_workspace", None) is None or self._bwd_workspace.numel() < need:
            self._bwd_workspace = torch.empty(need, device=self.device, dtype=torch.uint8)
        as_table = lambda g: g if g is None or isinstance(g, _lib.ParamGrads) else self._grad_table(g)
        gc, gf = as_table(grads_coarse), as_table(grads_fine)
        dp = (C.c_void_p * len(segs))(*[_ptr(d, name="d_poses").value for d in d_poses])
        self._check(self.lib.bnrf_render_backward_multi(
            self._ctx, arr, len(segs), _ptr(d_rgb_map, name="d_rgb_map"), _ptr(d_rgb0, name="d_rgb0"), _ptr(saved, torch.uint8, name="saved"),
            saved.numel(), C.byref(gc) if gc is not None else None, C.byref(gf) if gf is not None else None, dp,
            C.c_void_p(self._bwd_workspace.data_ptr()), self._bwd_workspace.numel(), _stream()), "bnrf_render_backward_multi")

    # -- a16: backward ------------------------------------------------------------------------
    def saved_bytes(self, n_rays):
        return int(self.lib.bnrf_saved_bytes(self._ctx, int(n_rays)))

    def _grad_table(self, grads):
        """grads: {'pts_linears.0.weight': tensor, ...} fp32 device tensors in the parameters' own shapes."""
        t, keep = _lib.ParamGrads(), []
        for i, name in enumerate(_lib.LINEAR_NAMES):
            t.weights[i] = _ptr(grads[name + ".weight"], device=self.device, name="grad " + name + ".weight").value
            t.biases[i] = _ptr(grads[name + ".bias"], device=self.device, name="grad " + name + ".bias").value
        return t

    def render_backward(self, poses, ray_idx, H, W, K, saved, d_rgb_map, d_rgb0, grads_coarse, grads_fine, d_poses, remap=None):
        """Adds d loss / d parameters into grads_* and d loss / d poses into d_poses (bnrf_render_backward)."""
        P, R = poses.shape[0], ray_idx.numel()
        need = self.lib.bnrf_backward_workspace_bytes(self._ctx, P * R)
        if getattr(self, "_bwd_workspace", None) is None or self._bwd_workspace.numel() < need:
            self._bwd_workspace = torch.empty(need, device=self.device, dtype=torch.uint8)
        # gradient tables: {name: tensor} dicts, or bnrf_param_grads structs built once by the caller (_grad_table)
        as_table = lambda g: g if g is None or isinstance(g, _lib.ParamGrads) else self._grad_table(g)
        gc, gf = as_table(grads_coarse), as_table(grads_fine)
        self._check(self.lib.bnrf_render_backward(
            self._ctx, _ptr(poses, name="poses"), _ptr(ray_idx, torch.int64, name="ray_idx"), P, R, int(H), int(W), self._K(K),
            _ptr(remap, name="remap"), _ptr(d_rgb_map, name="d_rgb_map"), _ptr(d_rgb0, name="d_rgb0"),
            _ptr(saved, torch.uint8, name="saved"), saved.numel(), C.byref(gc) if gc is not None else None,
            C.byref(gf) if gf is not None else None, _ptr(d_poses, name="d_poses"),
            C.c_void_p(self._bwd_workspace.data_ptr()), self._bwd_workspace.numel(), _stream()), "bnrf_render_backward")

    def spline_poses_backward(self, knots, transform, ts, d_poses, traj="spline", d_knots=None, d_transform=None):
        """d_knots [4,6] / d_transform [6]: ADDED into when given (the kernel accumulates), else fresh zero tensors."""
        if d_knots is None:
            d_knots = torch.zeros(4, 6, device=self.device, dtype=torch.float32)
        if d_transform is None and transform is not None:
            d_transform = torch.zeros(6, device=self.device, dtype=torch.float32)
        self._check(self.lib.bnrf_spline_poses_backward(
            self._ctx, _ptr(knots, name="knots"), _ptr(transform, name="transform"), _ptr(ts, name="ts"), ts.numel(), TRAJ[traj],
            _ptr(d_poses, name="d_poses"), _ptr(d_knots), _ptr(d_transform), _stream()), "bnrf_spline_poses_backward")
        return d_knots, d_transform

    def debug_sgemm(self, A, B, ta, tb, M, N, K, epi=0, C_out=None, mask=None, r_row=None, r_col=None):
        out = C_out if C_out is not None else torch.zeros(M, N, device=self.device, dtype=torch.float32)
        self._check(self.lib.bnrf_debug_sgemm(
            self._ctx, int(ta), int(tb), M, N, K, _ptr(A), A.stride(0), _ptr(B), B.stride(0), _ptr(out), out.stride(0), int(epi),
            _ptr(mask), mask.stride(0) if mask is not None else 0, _ptr(r_row), 1, _ptr(r_col), _stream()), "bnrf_debug_sgemm")
        return out

    def debug_tile_dgrad(self, A, B, mask=None, r_row=None, r_col=None, out=None):
        """out[rows,N] = A[rows,K] @ B[N,K].T (+ outer(r_row, r_col)) (* (mask > 0)) through bwd_tiles.cu (tests)."""
        rows, K = A.shape
        N = B.shape[0]
        acc = out is not None
        if out is None:
            out = torch.zeros(rows, N, device=self.device, dtype=torch.float32)
        self._check(self.lib.bnrf_debug_tile_dgrad(self._ctx, rows, K, N, _ptr(A), _ptr(B), _ptr(mask), _ptr(r_row), _ptr(r_col),
                                                   int(acc), _ptr(out), _stream()), "bnrf_debug_tile_dgrad")
        return out

    def debug_tile_wgrad(self, dz, h, dW, col0=0, n_valid=None, dB=None, wrow=None, dWv=None, dBv=None):
        """dW[:, col0:col0+n_valid] += dz.T @ h[:, :n_valid]; dB += dz.sum(0); dWv += wrow @ h; dBv += wrow.sum() (tests)."""
        rows, M = dz.shape
        N = h.shape[1]
        self._check(self.lib.bnrf_debug_tile_wgrad(self._ctx, rows, M, N, _ptr(dz), _ptr(h), _ptr(wrow), _ptr(dW), dW.stride(0), int(col0),
                                                   int(N if n_valid is None else n_valid), _ptr(dB), _ptr(dWv), _ptr(dBv), _stream()),
                    "bnrf_debug_tile_wgrad")
        return dW

    # -- stage operators ------------------------------------------------------------------
    def op_rays(self, poses, ray_idx, H, W, K, remap=None):
        n = poses.shape[0] * ray_idx.numel()
        o, d, v = (torch.empty(n, 3, device=self.device) for _ in range(3))
        self._check(self.lib.bnrf_op_rays(self._ctx, _ptr(poses), _ptr(ray_idx, torch.int64), poses.shape[0], ray_idx.numel(),
                                          int(H), int(W), self._K(K), _ptr(remap), _ptr(o), _ptr(d), _ptr(v), _stream()), "bnrf_op_rays")
        return o, d, v

    def op_stratified(self, t_rand):
        z = torch.empty_like(t_rand)
        self._check(self.lib.bnrf_op_stratified(self._ctx, _ptr(t_rand), t_rand.shape[0], t_rand.shape[1], _ptr(z), _stream()), "bnrf_op_stratified")
        return z

    def op_mlp(self, net, rays_o, rays_d, viewdirs, z):
        n, S = z.shape
        raw = torch.empty(n, S, self.channels + 1, device=self.device)
        self._check(self.lib.bnrf_op_mlp(self._ctx, int(net), _ptr(rays_o), _ptr(rays_d), _ptr(viewdirs), _ptr(z), n, S, _ptr(raw), _stream()), "bnrf_op_mlp")
        return raw

    def op_composite(self, raw, z, rays_d, noise):
        n, S = z.shape
        new = lambda *s: torch.empty(*s, device=self.device)
        out = {"rgb_map": new(n, self.channels), "disp_map": new(n), "acc_map": new(n), "weights": new(n, S),
               "depth_map": new(n), "sigma": new(n, S)}
        self._check(self.lib.bnrf_op_composite(self._ctx, _ptr(raw), _ptr(z), _ptr(rays_d), _ptr(noise), n, S,
                                               *[_ptr(out[k]) for k in ("rgb_map", "disp_map", "acc_map", "weights", "depth_map", "sigma")],
                                               _stream()), "bnrf_op_composite")
        return out

    def op_resample(self, z_coarse, weights, u):
        n, S = z_coarse.shape
        K = u.shape[1]
        zf = torch.empty(n, S + K, device=self.device)
        self._check(self.lib.bnrf_op_resample(self._ctx, _ptr(z_coarse), _ptr(weights), _ptr(u), n, S, K, _ptr(zf), _stream()), "bnrf_op_resample")
        return zf


# -- image formation: context-free entry points ---------------------------------------------
def adam_step(params, grads, exp_avg, exp_avg_sq, groups, step, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0, zero_grads=True):
    """bnrf_adam_step over flat fp32 device buffers; groups: [(begin, end, lr, active), ...] (see benerf_b200.h)."""
    lib = _lib.load()
    n = params.numel()
    arr = (_lib.AdamGroup * len(groups))(*[_lib.AdamGroup(int(b), int(e), float(lr), int(bool(a))) for b, e, lr, a in groups])
    _rc(lib.bnrf_adam_step(_ptr(params, name="params"), _ptr(grads, name="grads"), _ptr(exp_avg, name="exp_avg"),
                           _ptr(exp_avg_sq, name="exp_avg_sq"), n, arr, len(groups), int(step), float(betas[0]), float(betas[1]),
                           float(eps), float(grad_scale), int(bool(zero_grads)), _stream()), "bnrf_adam_step")


def adam_step_sched(params, grads, exp_avg, exp_avg_sq, groups, step_dev, decay_steps, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0,
                    zero_grads=True, advance_scratch=None):
    """bnrf_adam_step_sched: groups [(begin, end, lr0, decay_rate, active), ...]; step_dev: int64 device scalar = global_step.
    advance_scratch: an int64 device scalar (zero-initialised, owned by the caller): the launch also increments step_dev."""
    lib = _lib.load()
    arr = (_lib.AdamSchedGroup * len(groups))(*[_lib.AdamSchedGroup(int(b), int(e), float(lr0), float(rate), int(bool(a)))
                                                for b, e, lr0, rate, a in groups])
    _rc(lib.bnrf_adam_step_sched(_ptr(params, name="params"), _ptr(grads, name="grads"), _ptr(exp_avg, name="exp_avg"),
                                 _ptr(exp_avg_sq, name="exp_avg_sq"), params.numel(), arr, len(groups),
                                 _ptr(step_dev, torch.int64, name="step_dev"), float(decay_steps), float(betas[0]), float(betas[1]),
                                 float(eps), float(grad_scale), int(bool(zero_grads)),
                                 _ptr(advance_scratch, torch.int64, name="advance_scratch") if advance_scratch is not None else None,
                                 _stream()), "bnrf_adam_step_sched")


def step_advance(step_dev):
    _rc(_lib.load().bnrf_step_advance(_ptr(step_dev, torch.int64, name="step_dev"), _stream()), "bnrf_step_advance")


def loss_cfg(args):
    """The flags of train.py:205-331 as a bnrf_loss_cfg.  event_loss / rgb_loss are store_true flags upstream (config.py:215-218)."""
    return _lib.LossCfg(int(args.channels), int(args.num_interpolated_pose), LOG_MODE[args.dataset],
                        int(bool(getattr(args, "event_loss", True))), int(bool(getattr(args, "rgb_loss", True))),
                        float(args.event_threshold), float(getattr(args, "event_coeff_syn", 0.1)),
                        float(getattr(args, "event_coeff_real", 2.0)), float(getattr(args, "rgb_coeff", 1.0)))


def training_loss_fused(cfg, evt_fine, evt_coarse, events_accu, idx_evt, blur_fine, blur_coarse, blur_target, all_reduce=None,
                        out_like=None, split=None):
    """bnrf_training_loss (+ bnrf_training_loss_finish): the loss block of train.py:163-331 and its gradients w.r.t. the four
    renders in one or two launches.  all_reduce: callable summing a float64 device tensor over the ranks in place (the five batch
    sums of the normalised event loss), or None.  Returns (loss_out float64 [5], (d_evt_fine, d_evt_coarse, d_blur_fine, d_blur_coarse)) -- or, with out_like = (fine, coarse) [n, C]
    batch tensors whose rows [0, split) are the event render and [split, n) the blur render, the two [n, C] gradient buffers."""
    lib = _lib.load()
    dev = evt_fine.device
    R_e, R_b = idx_evt.numel(), blur_target.shape[0]
    ws = torch.empty(lib.bnrf_training_loss_workspace_bytes(R_e) // 8 + 1, device=dev, dtype=torch.float64)
    if out_like is not None:
        # the event and blur renders are the two row ranges [0, split) / [split, n) of one batch (Engine.render_multi): their
        # gradients are written into the matching halves of ONE [n, C] buffer per level, ready for render_backward_multi
        full = (torch.empty_like(out_like[0]), torch.empty_like(out_like[1]))
        grads = (full[0][:split], full[1][:split], full[0][split:], full[1][split:])
    else:
        full = None
        grads = tuple(torch.empty_like(t) for t in (evt_fine, evt_coarse, blur_fine, blur_coarse))
    loss_out = torch.empty(5, device=dev, dtype=torch.float64)
    ev = events_accu.reshape(-1)
    _rc(lib.bnrf_training_loss(C.byref(cfg), _ptr(evt_fine, name="evt_fine"), _ptr(evt_coarse, name="evt_coarse"),
                               _ptr(ev, torch.float64, name="events_accu"), _ptr(idx_evt, torch.int64, name="idx_evt"), R_e,
                               _ptr(blur_fine, name="blur_fine"), _ptr(blur_coarse, name="blur_coarse"),
                               _ptr(blur_target, name="blur_target"), R_b, _ptr(ws, torch.float64), *[_ptr(g) for g in grads],
                               _ptr(loss_out, torch.float64), _stream()), "bnrf_training_loss")
    if cfg.event_loss and cfg.event_threshold <= 0:
        if all_reduce is not None:
            all_reduce(ws[:5])
        _rc(lib.bnrf_training_loss_finish(C.byref(cfg), _ptr(evt_fine), _ptr(evt_coarse), _ptr(ev, torch.float64),
                                          _ptr(idx_evt, torch.int64), R_e, R_b, _ptr(ws, torch.float64), _ptr(grads[0]), _ptr(grads[1]),
                                          _ptr(loss_out, torch.float64), _stream()), "bnrf_training_loss_finish")
    return loss_out, (grads if full is None else full)


def _rc(rc, what):
    if rc != _lib.OK:
        raise BnrfError(f"{what} failed ({rc})")


class _BlurMeanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgb, n_poses):
        ctx.n_poses, ctx.shape = n_poses, rgb.shape
        return _blur_mean_raw(rgb.contiguous(), n_poses)

    @staticmethod
    def backward(ctx, g):
        lib = _lib.load()
        d = torch.empty(ctx.shape, device=g.device, dtype=torch.float32)
        C_ = ctx.shape[-1]
        _rc(lib.bnrf_blur_mean_backward(_ptr(g.contiguous(), name="g"), int(ctx.n_poses), g.numel() // C_, C_, _ptr(d), _stream()),
            "bnrf_blur_mean_backward")
        return d, None


class _EventLogDiffFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rgb, n_bins, mode):
        rgb = rgb.contiguous()
        ctx.save_for_backward(rgb)
        ctx.n_bins, ctx.mode = n_bins, mode
        return _event_logdiff_raw(rgb, n_bins, mode)

    @staticmethod
    def backward(ctx, g):
        (rgb,) = ctx.saved_tensors
        lib = _lib.load()
        C_ = rgb.shape[-1]
        R = rgb.numel() // ((ctx.n_bins + 1) * C_)
        d = torch.empty_like(rgb)
        _rc(lib.bnrf_event_logdiff_backward(_ptr(rgb, name="rgb"), _ptr(g.contiguous(), name="g"), int(ctx.n_bins), R, C_,
                                            LOG_MODE[ctx.mode], _ptr(d), _stream()), "bnrf_event_logdiff_backward")
        return d, None, None


def blur_mean(rgb, n_poses):
    """[P*R, C] pose-major (or [P,R,C]) -> [R, C]   (train.py:299-318).  Differentiable (bnrf_blur_mean_backward)."""
    if torch.is_grad_enabled() and rgb.requires_grad:
        return _BlurMeanFn.apply(rgb, n_poses)
    return _blur_mean_raw(rgb, n_poses)


def event_logdiff(rgb, n_bins, dataset_or_mode):
    """[(B+1)*R, C] pose-major -> [B, R] log-brightness differences of consecutive poses (train.py:205-292).
    Differentiable (bnrf_event_logdiff_backward)."""
    if torch.is_grad_enabled() and rgb.requires_grad:
        return _EventLogDiffFn.apply(rgb, n_bins, dataset_or_mode)
    return _event_logdiff_raw(rgb, n_bins, dataset_or_mode)


def _blur_mean_raw(rgb, n_poses):
    lib = _lib.load()
    C_ = rgb.shape[-1]
    R = rgb.numel() // (n_poses * C_)
    out = torch.empty(R, C_, device=rgb.device, dtype=torch.float32)
    _rc(lib.bnrf_blur_mean(_ptr(rgb, name="rgb"), int(n_poses), R, C_, _ptr(out), _stream()), "bnrf_blur_mean")
    return out


def _event_logdiff_raw(rgb, n_bins, dataset_or_mode):
    lib = _lib.load()
    C_ = rgb.shape[-1]
    R = rgb.numel() // ((n_bins + 1) * C_)
    out = torch.empty(n_bins, R, device=rgb.device, dtype=torch.float32)
    _rc(lib.bnrf_event_logdiff(_ptr(rgb, name="rgb"), int(n_bins), R, C_, LOG_MODE[dataset_or_mode], _ptr(out), _stream()), "bnrf_event_logdiff")
    return out


def accumulate_events(x, y, pol, H, W, out=None):
    """int32 x, y, float pol (device) -> float64 [H, W] polarity image (utils/event_utils.py:247-259)."""
    lib = _lib.load()
    if out is None:
        out = torch.zeros(H, W, device=x.device, dtype=torch.float64)
    _rc(lib.bnrf_accumulate_events(_ptr(x, torch.int32, name="x"), _ptr(y, torch.int32, name="y"), _ptr(pol, name="pol"),
                                   x.numel(), int(H), int(W), _ptr(out, torch.float64), _stream()), "bnrf_accumulate_events")
    return out


def accumulate_events_binned(x, y, pol, bounds, H, W, out=None):
    """`len(bounds) - 1` consecutive windows of a time-sorted event array in one launch: events [bounds[b], bounds[b+1]) ->
    float64 image b of out [bins, H, W] (each window as utils/event_utils.py:247-259; the bins of get_pose_evt(..., seg_num),
    model/optimize.py:58-71).  bounds: int64, non-decreasing, within [0, len(x)] (e.g. torch.searchsorted(ts, bin_edges))."""
    lib = _lib.load()
    bounds = torch.as_tensor(bounds).to(device=x.device, dtype=torch.int64).contiguous()
    bins = bounds.numel() - 1
    if bins < 0:
        raise ValueError("bounds needs at least one entry")
    if bins > 0 and (bool((bounds[1:] < bounds[:-1]).any()) or int(bounds[0]) < 0 or int(bounds[-1]) > x.numel()):
        raise ValueError("bounds must be non-decreasing and within [0, number of events]")
    if out is None:
        out = torch.zeros(bins, H, W, device=x.device, dtype=torch.float64)
    elif tuple(out.shape) != (bins, int(H), int(W)):
        raise ValueError(f"out must have shape ({bins}, {H}, {W})")
    _rc(lib.bnrf_accumulate_events_binned(_ptr(x, torch.int32, name="x"), _ptr(y, torch.int32, name="y"), _ptr(pol, name="pol"),
                                          x.numel(), _ptr(bounds, torch.int64, name="bounds"), bins, int(H), int(W),
                                          _ptr(out, torch.float64), _stream()), "bnrf_accumulate_events_binned")
    return out
