#!/bin/bash
# Quick loop for MLP kernel work: operator-level parity (own process), stall trace, fine-pass timing.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -s --timeout 120 -k "mlp" > gpurun_out/mlp_ops.log 2>&1
echo "mlp ops exit $?"; tail -n 15 gpurun_out/mlp_ops.log
for m in ${1:-tc tc1}; do
  MLP_MODE=$m timeout 120 python tools/mlp_trace.py ${2:-9472} > gpurun_out/trace_$m.log 2>&1
  echo "trace $m exit $?"; cat gpurun_out/trace_$m.log
done
