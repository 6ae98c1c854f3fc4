#!/bin/bash
# launch list of the training step (per-kernel device time, cold-cache / serialised: compare shares)
mkdir -p gpurun_out
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
TAG=${1:-x}
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train_$TAG.csv \
    python bench.py --mode train --steps 2 --warmup 3 > gpurun_out/launches_train_$TAG.log 2>&1
echo "launch list exit $?"
python tools/launch_shares.py gpurun_out/launches_train_$TAG.csv | head -${2:-16}
