#!/usr/bin/env python
"""Per-CTA busy time of the pair weight-gradient kernel inside one training-shaped backward pass (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from argparse import Namespace
from benerf_b200 import optimize, run_nerf_helpers
from tests.cases import CASES
from tests.test_gpu_backward import case_args


def main():
    case = CASES["e2nerf_syn"]
    args = case_args(case)
    torch.manual_seed(0)
    model = optimize.Model(args)
    g = model.build_network(args)
    run_nerf_helpers.init_nerf(g.nerf); run_nerf_helpers.init_nerf(g.nerf_fine)
    g.to("cuda")
    eng = g.engine(args)
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 107
    idx = torch.randint(0, case.H * case.W, (R,), device="cuda")
    poses = g.get_pose_rgb(args, torch.tensor(case.exposure))
    for it in range(3):
        ret = g.render(0, poses, idx, case.H, case.W, case.K, args, enable_crf=True, sensor_type="rgb", remap=None, training=True)
        if it == 2:
            tr = eng.mlp_trace(True)
        (ret["rgb_map"].sum()).backward()            # fine network only: one pair-wgrad launch (the last one overwrites the trace)
        torch.cuda.synchronize()
    t = tr.reshape(-1)[:148 * 4].reshape(148, 4).cpu()
    cyc = t[:, 0].double()
    print(f"rays {19 * R}  busy cycles: min {cyc.min():.0f} mean {cyc.mean():.0f} max {cyc.max():.0f}")
    order = torch.argsort(cyc, descending=True)
    for b in order[:12].tolist() + order[-4:].tolist():
        print(f"  cta {b:3d} cycles {int(t[b,0]):9d} stages {int(t[b,1]):5d} segs {int(t[b,2])} first job {int(t[b,3])}  cycles/stage {float(t[b,0]) / max(int(t[b,1]),1):7.0f}")
    eng.mlp_trace(False)


if __name__ == "__main__":
    main()
