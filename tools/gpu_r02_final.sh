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
# 5) per-tile stall accounting of the three tensor-core kernels (clock64 counters written by the kernels themselves)
{ echo "== forward kernel, inference mode (tools/mlp3_trace.py)"; timeout 200 python tools/mlp3_trace.py 2>&1 | tail -24;
  echo; echo "== forward kernel, training mode (tools/mlp3_trace_train.py 107)"; timeout 200 python tools/mlp3_trace_train.py 107 2>&1 | tail -24; } > gpurun_out/r02_mlp_tc3_stall_trace.txt
{ echo "== dgrad chain on CTA pairs (tools/chain_trace.py 107)"; timeout 200 python tools/chain_trace.py 107 2>&1 | tail -30;
  echo; echo "== weight gradients on CTA pairs (tools/wgrad_trace.py 107)"; timeout 200 python tools/wgrad_trace.py 107 2>&1 | tail -30; } > gpurun_out/r02_bwd_stall_trace.txt
# 6) measured parity numbers of every golden case
timeout 600 python tools/parity_report.py gpurun_out/r02_parity_report.json > gpurun_out/r02_parity_report.log 2>&1
echo "parity report exit $?"
# 7) the bench lines themselves (never under ncu): our arm, then the CPU reference arm
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
echo "reference arm exit $?"
for r in r02_mlp_tc3_bench r02_train_kernels; do
  python tools/ncu_key_metrics.py gpurun_out/$r.ncu-rep > gpurun_out/$r.csv 2>/dev/null
  $NCU -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null
done
ls -la gpurun_out | grep "r02_"
