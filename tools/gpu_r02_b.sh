#!/bin/bash
# round 2, call B: full GPU suite + training-step bench (eager vs graph) + render quick bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 15
python bench.py --mode train --steps 20 --warmup 5 2>&1 | tail -n 1 | cut -c1-700
python bench.py --mode train --train-config e2nerf_real --steps 20 --warmup 5 2>&1 | tail -n 1 | cut -c1-500
