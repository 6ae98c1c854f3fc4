#!/usr/bin/env python
"""Hand-off timeline of the dgrad chain kernel (dgrad_chain2.cu; debug aid): one training-shaped backward pass."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
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
        (ret["rgb_map"].sum() + ret["rgb0"].sum()).backward()
        torch.cuda.synchronize()
    t = tr.reshape(-1)[3072:3072 + 9 * 16].cpu().tolist()
    t0 = t[10]
    print(" step | mma: reach  commit h0  commit h1 | epi(w8): acc h0  chunks 0..3 -> | acc h1  chunks 4..7 ->")
    for s in range(9):
        r = [x - t0 if x else 0 for x in t[s * 16: s * 16 + 16]]
        print(f"  {s}   | {r[10]:7d} {r[11]:8d} {r[12]:8d} | {r[0]:7d} {r[1]:6d} {r[2]:6d} {r[3]:6d} {r[4]:6d} | {r[9]:7d} {r[5]:6d} {r[6]:6d} {r[7]:6d} {r[8]:6d}")
    eng.mlp_trace(False)


if __name__ == "__main__":
    main()
