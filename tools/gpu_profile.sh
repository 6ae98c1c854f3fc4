#!/bin/bash
# ncu evidence (1 GPU). Outputs under gpurun_out/; summaries are copied to profiles/ by hand.
mkdir -p gpurun_out
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
PIX=${1:-8192}
# 1) every launch of one bench step with its device time (cold-cache, serialised: compare shares)
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --pixels $PIX --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
echo "launch list exit $?"
# 2) full capture of the dominant kernel (fine-pass-sized MLP launch, second of three)
timeout 900 $NCU --set full --clock-control none --import-source on -k regex:mlp_tc -s 1 -c 1 -f -o gpurun_out/prof_mlp_tc \
    python tools/mlp_trace.py ${2:-37888} > gpurun_out/prof_mlp.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out | head -30
