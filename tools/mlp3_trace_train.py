#!/usr/bin/env python
"""Stall accounting of the in-TMEM MLP kernel in TRAINING mode (activation spill through the staging ring): the fine pass of a
training-shaped render (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.cases import CASES, make_inputs
from benerf_b200.engine import Engine
from tools.mlp3_trace import NAMES


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 107
    case = CASES["unreal_rgb"]
    inp = make_inputs(case)
    dev = "cuda"
    eng = Engine(mlp_mode="tc")
    eng.set_weights(0, {k: v.to(dev) for k, v in inp["coarse"].items()})
    eng.set_weights(1, {k: v.to(dev) for k, v in inp["fine"].items()})
    from oracle import pose
    poses = pose.poses_from_knots(inp["knots"], None, 0.2, 0.8, 19).to(dev).contiguous()
    idx = torch.randint(0, case.H * case.W, (R,), device=dev)
    n = 19 * R
    saved = torch.empty(eng.saved_bytes(n), device=dev, dtype=torch.uint8)
    for _ in range(2):
        eng.render(poses, idx, case.H, case.W, case.K, seed=1, saved=saved)
    tr = eng.mlp_trace(True)
    eng.render(poses, idx, case.H, case.W, case.K, seed=2, saved=saved)
    torch.cuda.synchronize()
    t = tr[:148].double().cpu()
    tiles_per_cta = n * 128 / 128 / 148
    print(f"fine pass: rows={n * 128} tiles/CTA={tiles_per_cta:.1f}")
    for i, name in enumerate(NAMES):
        col = t[:, i]
        print(f"  {name:<22s} mean {col.mean():12.0f} cyc  per tile {col.mean() / tiles_per_cta:9.0f}  min {col.min():12.0f} max {col.max():12.0f}")
    tl = tr.reshape(-1)[148 * 16: 148 * 16 + 256].cpu().tolist()
    t0 = tl[0]
    print(" step | mma: first  a_rdy kb0  kb1   kb2   kb3  commit h0  h1 | epi(w8): acc h0  chunk a  chunk b  acc h1  chunk c  chunk d")
    for s in range(9):
        m = [x - t0 if x else 0 for x in tl[s * 8: s * 8 + 8]]
        e = [x - t0 if x else 0 for x in tl[128 + s * 8: 128 + s * 8 + 8]]
        print(f"  {s:2d}  | {m[0]:7d} {m[1]:7d} {m[2]:6d} {m[3]:6d} {m[4]:6d} {m[5]:8d} {m[6]:7d} | {e[0]:8d} {e[1]:8d} {e[2]:8d} {e[3]:8d} {e[4]:8d} {e[5]:8d}")
    eng.mlp_trace(False)


if __name__ == "__main__":
    main()
