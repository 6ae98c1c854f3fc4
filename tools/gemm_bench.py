#!/usr/bin/env python
"""Device timing of the backward pass's GEMM shapes through bnrf_debug_sgemm (bring-up aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from benerf_b200.engine import Engine

def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
    modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["tc", "simt"]
    dev = "cuda"
    for mode in modes:
        eng = Engine(gemm_mode=mode)
        dz = torch.randn(rows, 256, device=dev); h = torch.randn(rows, 256, device=dev); wt = torch.randn(256, 256, device=dev)
        out = torch.empty(rows, 256, device=dev); dw = torch.zeros(256, 256, device=dev)
        cases = {
            "dgrad masked [rows,256]x[256,256]": lambda: eng.debug_sgemm(dz, wt, 0, 1, rows, 256, 256, epi=3, C_out=out, mask=h),
            "dgrad store  [rows,256]x[256,256]": lambda: eng.debug_sgemm(dz, wt, 0, 1, rows, 256, 256, epi=0, C_out=out),
            "wgrad atomic [256,rows]x[rows,256]": lambda: eng.debug_sgemm(dz, h, 1, 0, 256, 256, rows, epi=2, C_out=dw),
        }
        for name, fn in cases.items():
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            flops = 2.0 * rows * 256 * 256
            gb = rows * 256 * 4 * (3 if "masked" in name else 2) / 1e9
            print(f"{mode:5s} {name}: {ms*1e3:8.1f} us  {flops/ms/1e9:7.1f} TFLOP/s  {gb/ms*1e3:7.0f} GB/s (algorithmic)", flush=True)

if __name__ == "__main__":
    main()
