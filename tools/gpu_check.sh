#!/bin/bash
# Run the GPU parity suite in separate processes (a trapped kernel poisons its CUDA context).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for t in $(ls tests/test_gpu_*.py | xargs -n1 basename | sed 's/\.py$//'); do
  timeout 600 python -m pytest tests/$t.py -m gpu -q -s --timeout 180 > gpurun_out/$t.log 2>&1
  echo "$t exit $?" | tee -a gpurun_out/summary.txt
  tail -n 25 gpurun_out/$t.log
done
timeout 300 python tools/quick_bench.py ${1:-tc,simt} ${2:-1024} > gpurun_out/quick_bench.log 2>&1
echo "quick_bench exit $?" | tee -a gpurun_out/summary.txt
cat gpurun_out/quick_bench.log
