#!/usr/bin/env python
"""Stall accounting of the in-TMEM MLP kernel (mlp_tc3.cu; debug aid): one fine-pass-sized launch, per-role wait cycles and
the hand-off timeline of one tile."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.cases import CASES, make_inputs
from benerf_b200.engine import Engine

NAMES = ["kernel total", "mma wait PE_FULL", "mma wait A_READY", "mma wait W_FULL", "mma loop total", "tma wait W_EMPTY",
         "epi wait ACC_FULL", "epi loop total", "front wait PE_EMPTY", "front loop total", "epi wait ST_EMPTY"]


def main():
    n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 64
    S = 128
    case = CASES["unreal_rgb"]
    inp = make_inputs(case)
    dev = "cuda"
    eng = Engine(mlp_mode="tc")
    eng.set_weights(1, {k: v.to(dev) for k, v in inp["fine"].items()})
    g = torch.Generator().manual_seed(0)
    o = (torch.rand(n_rays, 3, generator=g) * 2 - 1).to(dev)
    d = (torch.rand(n_rays, 3, generator=g) * 2 - 1).to(dev)
    v = torch.nn.functional.normalize(d, dim=-1).contiguous()
    z = torch.sort(torch.rand(n_rays, S, generator=g), -1)[0].to(dev)
    for _ in range(2):
        eng.op_mlp(1, o, d, v, z)
    tr = eng.mlp_trace(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.op_mlp(1, o, d, v, z); e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    t = tr[:148].double().cpu()
    tiles_per_cta = n_rays * S / 128 / 148
    print(f"rows={n_rays * S} tiles/CTA={tiles_per_cta:.1f} op_mlp ms={ms:.3f} (incl. viewbias) -> {n_rays * S * 2 * 593408 / ms / 1e9:.1f} alg TFLOP/s")
    for i, name in enumerate(NAMES):
        col = t[:, i]
        print(f"  {name:<22s} mean {col.mean():12.0f} cyc  per tile {col.mean() / tiles_per_cta:9.0f}  min {col.min():12.0f} max {col.max():12.0f}")
    tl = tr.reshape(-1)[148 * 16: 148 * 16 + 256].cpu().tolist()
    t0 = tl[0]
    print("timeline of CTA 0, tile 3 (cycles relative to the first MMA of step 0)")
    print(" step | mma: first  a_rdy kb0  kb1   kb2   kb3  commit h0  h1 | epi(w8): acc h0  chunk a  chunk b  acc h1  chunk c  chunk d")
    for s in range(9):
        m = [x - t0 if x else 0 for x in tl[s * 8: s * 8 + 8]]
        e = [x - t0 if x else 0 for x in tl[128 + s * 8: 128 + s * 8 + 8]]
        print(f"  {s:2d}  | {m[0]:7d} {m[1]:7d} {m[2]:6d} {m[3]:6d} {m[4]:6d} {m[5]:8d} {m[6]:7d} | {e[0]:8d} {e[1]:8d} {e[2]:8d} {e[3]:8d} {e[4]:8d} {e[5]:8d}")
    eng.mlp_trace(False)


if __name__ == "__main__":
    main()
