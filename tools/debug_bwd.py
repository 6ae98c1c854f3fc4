#!/usr/bin/env python
"""Debug aid: per-ray comparison of d_raw (composite backward) and parameter gradients with the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import pose, render as orender
from tests.cases import CASES, make_inputs
from tests.test_gpu_backward import _build_graph, _oracle_leaves, rel_err
from tests.gpu_util import to_dev, DEV

name = sys.argv[1] if len(sys.argv) > 1 else "blender_gray_coarse"
case = CASES[name]; inp = make_inputs(case)
graph, args = _build_graph(case, inp)
knots, transform, coarse, fine_p = _oracle_leaves(case, inp)
poses_o = pose.poses_from_knots(knots, transform, *case.exposure, case.n_poses, case.traj)
draws = dict(inp["rng_rgb"])
want = orender.render(coarse, fine_p, poses_o, inp["idx_rgb"], case.H, case.W, case.K, draws, n_samples=case.n_samples,
                      n_importance=case.n_importance, channels=case.channels, return_intermediates=True)
ex = want["_extra"]
print("extra keys", list(ex.keys()))
raw_c = ex["raw_coarse"]; raw_c.retain_grad()
g = torch.Generator().manual_seed(3)
cot = torch.randn(want["rgb_map"].shape, generator=g)
(want["rgb_map"] * cot).sum().backward()
poses = graph.get_pose_rgb(args, torch.tensor(case.exposure, dtype=torch.float32))
got = graph.render(0, poses, inp["idx_rgb"], case.H, case.W, case.K, args, enable_crf=True, sensor_type="rgb", remap=None, training=True, rng=to_dev(draws))
(got["rgb_map"] * cot.to(DEV)).sum().backward()
torch.cuda.synchronize()
eng = graph.engine(args)
n = case.n_poses * case.r_rgb; S = case.n_samples; C = case.channels
d_raw = eng._bwd_workspace[: n * S * (C + 1) * 4].view(torch.float32).reshape(n, S, C + 1).cpu()
w = raw_c.grad.reshape(n, S, C + 1)
err = (d_raw - w).abs()
print("d_raw rel err", rel_err(d_raw, w), "max abs", float(err.max()), "at", divmod(int(err.reshape(n, -1).max(-1)[0].argmax()), 1))
per_ray = err.reshape(n, -1).max(-1)[0]
top = per_ray.topk(5)
for v, i in zip(top.values.tolist(), top.indices.tolist()):
    s = int(err[i].max(-1)[0].argmax())
    print(f" ray {i} sample {s}: err {v:.3e} ours {d_raw[i, s].tolist()} want {w[i, s].tolist()} |w|max {float(w[i].abs().max()):.3e}")
for pname, p in graph.nerf.named_parameters():
    print(pname, f"{rel_err(p.grad, coarse[pname].grad):.3e}")
