#!/usr/bin/env python
"""Quick device timing of the render path (not the bench contract; bring-up aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.cases import CASES, make_inputs
from benerf_b200.engine import Engine

def main():
    modes = sys.argv[1].split(",") if len(sys.argv) > 1 else ["tc", "simt"]
    R = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    case = CASES["unreal_rgb"]
    inp = make_inputs(case)
    dev = "cuda"
    from oracle import pose
    poses = pose.poses_from_knots(inp["knots"], None, 0.2, 0.8, 19).to(dev).contiguous()
    idx = torch.randperm(case.H * case.W)[:R].to(dev)
    for mode in modes:
        eng = Engine(mlp_mode=mode)
        eng.set_weights(0, {k: v.to(dev) for k, v in inp["coarse"].items()})
        eng.set_weights(1, {k: v.to(dev) for k, v in inp["fine"].items()})
        for _ in range(2):
            ret = eng.render(poses, idx, case.H, case.W, case.K, seed=1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 5 if mode.startswith("tc") else 2
        e0.record()
        for i in range(iters):
            ret = eng.render(poses, idx, case.H, case.W, case.K, seed=2 + i)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        n = 19 * R
        flops = n * 192 * 2 * 593408
        print(f"mode={mode} rays={n} ms={ms:.3f} rays/s={n/ms*1e3:.4g} alg TFLOP/s={flops/ms/1e9:.1f} "
              f"rgb_mean={ret['rgb_map'].mean().item():.4f} nan={int(torch.isnan(ret['rgb_map']).sum())}", flush=True)

if __name__ == "__main__":
    main()
