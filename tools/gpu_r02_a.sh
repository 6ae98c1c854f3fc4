#!/bin/bash
# round 2, call A: parity report + ncu evidence of the backward kernels as the training step launches them
mkdir -p gpurun_out
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
timeout 600 python tools/parity_report.py gpurun_out/parity_report.json > gpurun_out/parity_report.log 2>&1
echo "parity report exit $?"; tail -n 40 gpurun_out/parity_report.log
timeout 600 python -m pytest tests/test_gpu_render.py -m gpu -x -q 2>&1 | tail -n 15
timeout 900 $NCU --set full --clock-control none --import-source on \
    -k regex:"dgrad_chain_pair|tile_wgrad|rgb_head_wgrad|heads_backward|mlp_tc2" -s 20 -c 8 -f -o gpurun_out/prof_r02a_bwd \
    python bench.py --mode train --steps 1 --warmup 1 > gpurun_out/prof_r02a_bwd.log 2>&1
echo "ncu bwd exit $?"
python tools/ncu_key_metrics.py gpurun_out/prof_r02a_bwd.ncu-rep > gpurun_out/prof_r02a_bwd.csv 2>/dev/null
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train_r02a.csv \
    python bench.py --mode train --steps 2 --warmup 1 > gpurun_out/launches_train_r02a.log 2>&1
echo "launch list exit $?"
python bench.py --mode train --steps 20 --warmup 5 2>&1 | tail -n 3
ls -la gpurun_out
