#!/bin/bash
# launch list of the training step + ncu full of the big three (one launch each, fine network of the blur render)
mkdir -p gpurun_out
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
TAG=${1:-r02c}
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train_$TAG.csv \
    python bench.py --mode train --steps 2 --warmup 3 > gpurun_out/launches_train_$TAG.log 2>&1
echo "launch list exit $?"
python tools/launch_shares.py gpurun_out/launches_train_$TAG.csv | head -40
timeout 900 $NCU --set full --clock-control none --import-source on \
    -k regex:"dgrad_chain_pair|tile_wgrad|mlp_tc3" -s 36 -c 8 -f -o gpurun_out/prof_${TAG}_bwd \
    python bench.py --mode train --steps 1 --warmup 3 > gpurun_out/prof_${TAG}_bwd.log 2>&1
echo "ncu bwd exit $?"
python tools/ncu_key_metrics.py gpurun_out/prof_${TAG}_bwd.ncu-rep > gpurun_out/prof_${TAG}_bwd.csv 2>/dev/null
