#!/bin/bash
# Round-2 ncu evidence (1 GPU); outputs under gpurun_out/, summarised into profiles/ by tools/r02_profiles.py.
mkdir -p gpurun_out
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
PIX=${1:-65536}
B="--pixels $PIX --no-cpu-baseline --no-train-step"
# 1) every launch of one bench step with its device time (cold-cache, serialised: compare shares)
timeout 900 $NCU --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 $B > gpurun_out/r02_launches_bench.log 2>&1
echo "bench launch list exit $?"
# 2) full capture of the dominant kernel AS THE BENCH LAUNCHES IT: 4th MLP launch = fine pass of the 19-pose blur render
timeout 1200 $NCU --set full --clock-control none --import-source on -k regex:mlp_tc3_kernel -s 3 -c 1 -f -o gpurun_out/r02_mlp_tc3_bench \
    python bench.py --steps 1 --warmup 1 $B > gpurun_out/r02_mlp_tc3_bench.log 2>&1
echo "mlp full capture exit $?"
# 3) launch list of the training step
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_train.csv \
    python bench.py --mode train --steps 2 --warmup 3 > gpurun_out/r02_launches_train.log 2>&1
echo "train launch list exit $?"
# 4) the tensor-core kernels of one training step (both networks: forward, dgrad chain, weight gradients)
timeout 900 $NCU --set full --clock-control none --import-source on \
    -k regex:"dgrad_chain_pair|tile_wgrad|mlp_tc3" -s 24 -c 8 -f -o gpurun_out/r02_train_kernels \
    python bench.py --mode train --steps 1 --warmup 3 > gpurun_out/r02_train_kernels.log 2>&1
echo "train kernels exit $?"
for r in r02_mlp_tc3_bench r02_train_kernels; do
  python tools/ncu_key_metrics.py gpurun_out/$r.ncu-rep > gpurun_out/$r.csv 2>/dev/null
  $NCU -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null
done
ls -la gpurun_out | grep "r02_"
