#!/bin/bash
# ncu evidence (1 GPU). Outputs under gpurun_out/; tools/ncu_key_metrics.py turns the reports into profiles/*.csv.
mkdir -p gpurun_out
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
PIX=${1:-65536}
# 1) every launch of one bench step with its device time (cold-cache, serialised: compare shares)
timeout 900 $NCU --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --pixels $PIX --no-cpu-baseline --no-train-step > gpurun_out/launches_bench.log 2>&1
echo "launch list exit $?"
# 2) full capture of the dominant kernel AS THE BENCH LAUNCHES IT: 4th MLP launch = fine pass of the 19-pose blur render
timeout 1200 $NCU --set full --clock-control none --import-source on -k regex:mlp_tc2_kernel -s 3 -c 1 -f -o gpurun_out/prof_mlp_tc2_bench \
    python bench.py --steps 1 --warmup 1 --pixels $PIX --no-cpu-baseline --no-train-step > gpurun_out/prof_mlp_bench.log 2>&1
echo "mlp full capture exit $?"
# 3) the backward pass's tensor-core kernels inside one training step (chain + wgrad of the event render's fine network)
timeout 900 $NCU --set full --clock-control none --import-source on -k regex:"dgrad_chain|tile_wgrad|mlp_tc2" -s 4 -c 4 -f -o gpurun_out/prof_bwd_tiles \
    python bench.py --mode train --steps 1 --warmup 1 > gpurun_out/prof_bwd.log 2>&1
# 4) launch list of one training step
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv \
    python bench.py --mode train --steps 2 --warmup 1 > gpurun_out/launches_train.log 2>&1
echo "backward full capture exit $?"
for r in prof_mlp_tc2_bench prof_bwd_tiles; do
  python tools/ncu_key_metrics.py gpurun_out/$r.ncu-rep > gpurun_out/$r.csv 2>/dev/null
done
ls -la gpurun_out | grep "prof_\|launches_bench"
