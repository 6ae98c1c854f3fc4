#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference; the GPU box does not
have it):      python tools/make_golden.py

The reference is imported read-only with three absent third-party modules
stubbed (h5py, hdf5plugin, imageio[.v3]; none is on the arithmetic path,
SURVEY 8-c).  The four random draws of every Graph.render call are replaced by
the seeded tensors of tests/cases.py by patching torch.rand/torch.randn for the
duration of the call, so the CUDA engine can be fed the very same numbers.
The image-formation / loss block is inline code in the reference's train()
(train.py:163-337) and cannot be imported; it is replayed here line by line
with the reference's own RGB2Gray, rgb2brightlog and MSELoss callables.
"""
import os
import sys
import types
import contextlib
from argparse import Namespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("BENERF_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
from tests.cases import CASES, make_inputs, golden_path, make_rng_draws  # noqa: E402


def import_reference():
    """Stub the missing imports, then import the reference modules from REF."""
    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules.setdefault(name, m)
        return sys.modules[name]
    stub("h5py", File=object)
    stub("hdf5plugin")
    iio = stub("imageio")
    v3 = stub("imageio.v3", imwrite=lambda *a, **k: None, imread=lambda *a, **k: None)
    iio.v3 = v3
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import spline, run_nerf_helpers                       # noqa: E401
    from model import nerf, optimize, embedder
    from utils import math_utils, img_utils, event_utils
    from loss import imgloss
    return Namespace(spline=spline, helpers=run_nerf_helpers, nerf=nerf, optimize=optimize,
                     embedder=embedder, math_utils=math_utils, img_utils=img_utils,
                     event_utils=event_utils, imgloss=imgloss)


def ref_args(case):
    from tests.cases import BARF_MAX_ITER, BARF_START, BARF_END
    return Namespace(dataset=case.dataset, channels=case.channels, N_samples=case.n_samples,
                     N_importance=case.n_importance, multires=10, multires_views=4, i_embed=0,
                     use_viewdirs=True, use_barf_c2f=case.barf_iter >= 0, ndc=True, traj=case.traj,
                     num_interpolated_pose=case.n_poses, rgb_crf_net_hidden=0, rgb_crf_net_width=128,
                     event_crf_net_hidden=0, event_crf_net_width=128, chunk=4096, max_iter=BARF_MAX_ITER,
                     barf_c2f_start=BARF_START, barf_c2f_end=BARF_END)


@contextlib.contextmanager
def injected_rng(draws):
    """Serve torch.rand / torch.randn from a queue: t_rand, noise_c, u, noise_f."""
    queue = [("rand", draws["t_rand"]), ("randn", draws["noise_c"])]
    if "u" in draws:
        queue += [("rand", draws["u"]), ("randn", draws["noise_f"])]
    real_rand, real_randn = torch.rand, torch.randn

    def serve(kind):
        def f(*shape, **kw):
            want = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape)
            k, t = queue.pop(0)
            assert k == kind and tuple(t.shape) == want, (k, kind, t.shape, want)
            return t.clone()
        return f
    torch.rand, torch.randn = serve("rand"), serve("randn")
    try:
        yield
        assert not queue, "reference consumed fewer draws than expected"
    finally:
        torch.rand, torch.randn = real_rand, real_randn


@contextlib.contextmanager
def record_calls(graph, store, tag):
    """Record z_vals / raw / weights flowing through raw2output of both networks."""
    nets = [("c", graph.nerf)] + ([("f", graph.nerf_fine)] if hasattr(graph, "nerf_fine") else [])
    originals = []
    for lvl, net in nets:
        orig = net.raw2output

        def wrapped(crf, en, st, raw, z, d, _orig=orig, _lvl=lvl):
            out = _orig(crf, en, st, raw, z, d)
            store[f"{tag}_raw_{_lvl}"] = raw.detach().clone()
            store[f"{tag}_z_{_lvl}"] = z.detach().clone()
            store[f"{tag}_weights_{_lvl}"] = out[3].detach().clone()
            store[f"{tag}_depth_{_lvl}"] = out[4].detach().clone()
            store[f"{tag}_rays_d"] = d.detach().clone()
            return out
        net.raw2output = wrapped
        originals.append((net, orig))
    try:
        yield
    finally:
        for net, orig in originals:
            net.raw2output = orig


def build_graph(ref, case, inp):
    args = ref_args(case)
    graph = ref.optimize.Model(args).build_network(args)
    graph.nerf.load_state_dict(inp["coarse"])
    if case.n_importance > 0:
        graph.nerf_fine.load_state_dict(inp["fine"])
    graph.evt_knot_pose_se3.params.weight.data = torch.nn.Parameter(inp["knots"].clone())
    graph.transform.params.weight.data = torch.nn.Parameter(inp["transform"].clone())
    return graph, args


def run_case(ref, case):
    torch.manual_seed(0)
    inp = make_inputs(case)
    graph, args = build_graph(ref, case, inp)
    out = {}
    K = torch.Tensor(case.K)
    poses_evt = graph.get_pose_evt(args, torch.tensor(case.window, dtype=torch.float32))
    poses_rgb = graph.get_pose_rgb(args, torch.tensor(case.exposure, dtype=torch.float32))
    out["poses_evt"], out["poses_rgb"] = poses_evt.detach(), poses_rgb.detach()
    rets = {}
    for tag, poses, idx, draws, sensor in (("evt", poses_evt, inp["idx_evt"], inp["rng_evt"], "event"),
                                           ("rgb", poses_rgb, inp["idx_rgb"], inp["rng_rgb"], "rgb")):
        with injected_rng(draws), record_calls(graph, out, tag):
            ret = graph.render(max(case.barf_iter, 0), poses, idx, case.H, case.W, K, args, enable_crf=True,
                               sensor_type=sensor, remap=None, training=case.training)
        rets[tag] = ret
        for k, v in ret.items():
            out[f"{tag}_{k}"] = v.detach()

    # events: run the reference's accumulate function with .to('cuda') redirected to the CPU
    win = inp["events"]
    keep = np.where((case.window[0] <= win["ts"]) * (win["ts"] <= case.window[1]))   # model/nerf.py:170-178
    real_to = torch.Tensor.to
    torch.Tensor.to = lambda self, *a, **k: real_to(self, *[("cpu" if x == "cuda" else x) for x in a], **k)
    try:
        accu = ref.event_utils.accumulate_events_on_gpu(np.zeros((case.H, case.W)), win["x"][keep],
                                                        win["y"][keep], win["pol"][keep])
    finally:
        torch.Tensor.to = real_to
    out["events_accu"] = accu

    if case.n_importance > 0:
        # ---- replay of train.py:163-337 with the reference's own callables ----
        rgb2gray, mse_loss = ref.img_utils.RGB2Gray(), ref.imgloss.MSELoss()
        rgb2brightlog = ref.math_utils.rgb2brightlog
        ret_event, ret_rgb = rets["evt"], rets["rgb"]
        pixels_num = inp["idx_evt"].shape[0]
        g1 = {k: ret_event[k][:pixels_num] for k in ("rgb_map", "rgb0")}
        g2 = {k: ret_event[k][pixels_num:] for k in ("rgb_map", "rgb0")}
        target_s = accu.reshape(-1, 1)[inp["idx_evt"]]
        loss = 0
        ev_losses = {}
        if case.event_threshold > 0:
            target_s *= torch.tensor(case.event_threshold)
        for level in ("rgb_map", "rgb0"):
            if case.channels == 3:
                b2 = rgb2brightlog(rgb2gray(g2[level]), args.dataset)
                b1 = rgb2brightlog(rgb2gray(g1[level]), args.dataset)
            else:
                b2 = rgb2brightlog(g2[level], args.dataset)
                b1 = rgb2brightlog(g1[level], args.dataset)
            diff = b2 - b1
            out[f"event_diff_{level}"] = diff.detach()
            if case.event_threshold > 0:
                l = mse_loss(diff, target_s) * 0.1                      # event_coeff_syn
            else:
                rn = diff / (torch.linalg.norm(diff, dim=0, keepdim=True) + 1e-9)
                tn = target_s / (torch.linalg.norm(target_s, dim=0, keepdim=True) + 1e-9)
                l = mse_loss(rn, tn) * 2.0                              # event_coeff_real
            ev_losses[level] = l
        loss = loss + (ev_losses["rgb0"] + ev_losses["rgb_map"])
        interval = inp["blur_target"].shape[0]
        blur, blur0 = 0, 0
        for j in range(case.n_poses):
            blur = blur + ret_rgb["rgb_map"][j * interval:(j + 1) * interval]
            blur0 = blur0 + ret_rgb["rgb0"][j * interval:(j + 1) * interval]
            if (j + 1) % case.n_poses == 0:
                blur, blur0 = blur / case.n_poses, blur0 / case.n_poses
        out["blur_rgb_map"], out["blur_rgb0"] = blur.detach(), blur0.detach()
        rgb_loss = mse_loss(blur, inp["blur_target"]) * 1.0 + mse_loss(blur0, inp["blur_target"]) * 1.0
        loss = loss + rgb_loss
        out["loss"] = loss.detach().reshape(1)
        out["loss_parts"] = torch.stack([ev_losses["rgb_map"], ev_losses["rgb0"],
                                         mse_loss(blur, inp["blur_target"]),
                                         mse_loss(blur0, inp["blur_target"])]).detach()
        loss.backward()
        out["grad_knots"] = graph.evt_knot_pose_se3.params.weight.grad.clone()
        out["grad_transform"] = graph.transform.params.weight.grad.clone()
        for lvl, net in (("c", graph.nerf), ("f", graph.nerf_fine)):
            norms, samples = [], []
            for name, p in net.named_parameters():
                g = p.grad.reshape(-1)
                norms.append(g.norm())
                samples.append(g[::97])
            out[f"grad_norms_{lvl}"] = torch.stack(norms)
            out[f"grad_samples_{lvl}"] = torch.cat(samples)
    return {k: v.detach().cpu().numpy() for k, v in out.items()}


def run_functions(ref):
    """Function-level vectors for a2, a4, a6, a8, a9 on their own seeded inputs."""
    rng = np.random.default_rng(99)
    out = {}
    # a2: spline / linear at several motion scales, including the u==0 / u==1 nudges
    ts = torch.tensor([0.0, 1e-3, 0.25, 0.5, 0.731, 1.0], dtype=torch.float32)
    for s_i, scale in enumerate((0.01, 0.3, 1.5)):
        knots = torch.from_numpy(((rng.random((4, 6)) - 0.3) * scale).astype(np.float32))
        out[f"spline_knots_{s_i}"] = knots
        ks = [knots[i].reshape(1, 1, 6) for i in range(4)]
        out[f"spline_cubic_{s_i}"] = ref.spline.cubic_spline_pose_unit_time(*ks, ts.clone())
        out[f"spline_linear_{s_i}"] = ref.spline.linear_pose_unit_time(ks[0], ks[3], ts.clone())
    out["spline_ts"] = ts
    # a6: positional encoding of NDC-range points
    x = torch.from_numpy((rng.random((64, 3)) * 2.1 - 1.05).astype(np.float32))
    args = Namespace(use_barf_c2f=False)
    out["pe_x"] = x
    out["pe_pts"] = ref.embedder.get_embedder(args, 10, 0)[0](x)
    out["pe_dirs"] = ref.embedder.get_embedder(args, 4, 0)[0](x)
    # a9: sample_pdf on peaked / flat / all-zero weight rows
    n, b = 48, 63
    bins = torch.sort(torch.from_numpy(rng.random((n, b)).astype(np.float32)), -1)[0]
    w = torch.from_numpy(rng.random((n, b - 1)).astype(np.float32)) ** 8
    w[:4] = 0.0
    w[4:8, 10:] = 0.0
    u = torch.from_numpy(rng.random((n, 64), dtype=np.float32))
    out["pdf_bins"], out["pdf_weights"], out["pdf_u"] = bins, w, u
    real_rand = torch.rand
    torch.rand = lambda *a, **k: u.clone()
    try:
        out["pdf_samples"] = ref.helpers.sample_pdf(bins, w, 64)
    finally:
        torch.rand = real_rand
    # a8: raw2output for C = 3 and C = 1 on random raw
    for C in (3, 1):
        net = ref.nerf.NeRF(use_viewdirs=True, channels=C)
        raw = torch.from_numpy((rng.standard_normal((32, 64, C + 1)) * 2).astype(np.float32))
        z = torch.sort(torch.from_numpy(rng.random((32, 64), dtype=np.float32)), -1)[0]
        d = torch.from_numpy(rng.standard_normal((32, 3)).astype(np.float32))
        noise = torch.from_numpy(rng.standard_normal((32, 64), dtype=np.float32))
        real_randn = torch.randn
        torch.randn = lambda *a, **k: noise.clone()
        try:
            r = net.raw2output(None, True, "rgb", raw, z, d)
        finally:
            torch.randn = real_randn
        for key, val in zip(("rgb_map", "disp_map", "acc_map", "weights", "depth_map", "sigma"), r):
            out[f"r2o{C}_{key}"] = val
        out[f"r2o{C}_raw"], out[f"r2o{C}_z"], out[f"r2o{C}_d"], out[f"r2o{C}_noise"] = raw, z, d, noise
    # a3/a4: rays + ndc for one pose, eval-branch get_rays over a small image versus chosen pixels
    H, W, f = 12, 20, 15.0
    K = torch.tensor([[f, 0, W / 2], [0, f, H / 2], [0, 0, 1]], dtype=torch.float32)
    pose = ref.spline.cubic_spline_pose_unit_time(
        *[out["spline_knots_1"][i].reshape(1, 1, 6) for i in range(4)], torch.tensor([0.4]))[0]
    o, d = ref.helpers.get_rays(H, W, K, pose, Namespace(dataset="x"), None)
    on, dn = ref.helpers.ndc_rays(H, W, K[0][0], 1.0, o, d)
    out["rays_pose"], out["rays_o"], out["rays_d"], out["rays_o_ndc"], out["rays_d_ndc"] = pose, o, d, on, dn
    # f4: the TUM-VIE undistortion-LUT branch of Graph.render's training path (model/nerf.py:241-252): every integer pixel is
    # replaced by remap[j, i] before get_specific_rays; two poses, pose-major ray order
    Hr, Wr, fr = 10, 14, 11.0
    Kr = torch.tensor([[fr, 0, Wr / 2], [0, fr, Hr / 2], [0, 0, 1]], dtype=torch.float32)
    jj, ii = torch.meshgrid(torch.arange(Hr, dtype=torch.float32), torch.arange(Wr, dtype=torch.float32), indexing="ij")
    remap = torch.stack([ii + torch.from_numpy((rng.random((Hr, Wr)) - 0.5).astype(np.float32)) * 2.0,
                         jj + torch.from_numpy((rng.random((Hr, Wr)) - 0.5).astype(np.float32)) * 2.0], -1)
    idx = torch.from_numpy(rng.permutation(Hr * Wr)[:40].astype(np.int64))
    poses2 = ref.spline.cubic_spline_pose_unit_time(
        *[out["spline_knots_1"][i].reshape(1, 1, 6) for i in range(4)], torch.tensor([0.2, 0.7]))
    ray_idx_ = idx.repeat(poses2.shape[0])
    poses_rep = poses2.unsqueeze(1).repeat(1, idx.shape[0], 1, 1).reshape(-1, 3, 4)
    j = ray_idx_.reshape(-1, 1).squeeze() // Wr
    i = ray_idx_.reshape(-1, 1).squeeze() % Wr
    rect = remap[j, i]
    i, j = rect[..., 0], rect[..., 1]
    ro, rd = ref.helpers.get_specific_rays(i, j, Kr, poses_rep)
    ron, rdn = ref.helpers.ndc_rays(Hr, Wr, Kr[0][0], 1.0, ro, rd)
    out["remap_lut"], out["remap_idx"], out["remap_poses"] = remap, idx, poses2
    out["remap_rays_o"], out["remap_rays_d"], out["remap_rays_o_ndc"], out["remap_rays_d_ndc"] = ro, rd, ron, rdn
    # f4: the tone mappers (model/component.py:38-149), forward and gradients; parameters perturbed away from the init so that
    # every layer carries signal.  crf{i}: (class, hidden, width)
    from model import component
    for i, (cls, hidden, width) in enumerate(((component.ColorToneMapper, 0, 128), (component.LuminanceToneMapper, 0, 128),
                                              (component.ColorToneMapper, 2, 32))):
        torch.manual_seed(100 + i)
        m = cls(hidden=hidden, width=width, input_type="Gray")
        m.weights_biases_init()
        seq = m.mlp_gray if hasattr(m, "mlp_gray") else m.mlp_luminance
        with torch.no_grad():
            for p in seq.parameters():
                p.add_(torch.from_numpy((rng.standard_normal(tuple(p.shape)) * 0.3).astype(np.float32)))
        x = torch.from_numpy(rng.random((96, 1)).astype(np.float32)).requires_grad_(True)
        gy = torch.from_numpy(rng.standard_normal((96, 1)).astype(np.float32))
        y = m.forward(x)
        y.backward(gy)
        out[f"crf{i}_x"], out[f"crf{i}_y"], out[f"crf{i}_gy"], out[f"crf{i}_dx"] = x.detach(), y.detach(), gy, x.grad
        for j, p in enumerate(seq.parameters()):
            out[f"crf{i}_p{j}"], out[f"crf{i}_dp{j}"] = p.detach().clone(), p.grad.clone()
    return {k: v.detach().cpu().numpy() for k, v in out.items()}


def main():
    ref = import_reference()
    os.makedirs(os.path.dirname(golden_path("x")), exist_ok=True)
    only = set(sys.argv[1:])              # e.g. `python tools/make_golden.py functions` regenerates that fixture alone
    if not only or "functions" in only:
        np.savez_compressed(golden_path("functions"), **run_functions(ref))
        print("functions", os.path.getsize(golden_path("functions")))
    for name, case in CASES.items():
        if only and name not in only:
            continue
        data = run_case(ref, case)
        np.savez_compressed(golden_path(name), **data)
        print(name, os.path.getsize(golden_path(name)), "bytes;", len(data), "arrays")


if __name__ == "__main__":
    main()
