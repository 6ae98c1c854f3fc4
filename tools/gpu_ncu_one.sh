#!/bin/bash
# ncu --set full of ONE launch of a kernel inside the training step: $1 = kernel regex, $2 = launches to skip, $3 = tag
mkdir -p gpurun_out
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
timeout 900 $NCU --set full --clock-control none --import-source on -k regex:"$1" -s ${2:-4} -c 1 -f -o gpurun_out/prof_$3 \
    python bench.py --mode train --steps 1 --warmup 3 > gpurun_out/prof_$3.log 2>&1
echo "ncu exit $?"
python tools/ncu_key_metrics.py gpurun_out/prof_$3.ncu-rep > gpurun_out/prof_$3.csv 2>/dev/null
cat gpurun_out/prof_$3.csv
