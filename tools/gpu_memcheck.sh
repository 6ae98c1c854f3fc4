#!/bin/bash
# compute-sanitizer memcheck over the round-2 kernels: training iterations of every golden case (forward in training mode, fused loss,
# heads, dgrad chain, both wgrad kernels, tail, rays backward, spline backward), the Trainer (graph capture off and on), the
# operator-level MLP tests of all modes, tone mappers, binned event accumulation.
mkdir -p gpurun_out
CS=$(command -v compute-sanitizer || echo /usr/local/cuda/bin/compute-sanitizer)
LOG=gpurun_out/r02_memcheck.log
: > $LOG
run() {
  echo "### $*" >> $LOG
  timeout 1200 $CS --tool memcheck --error-exitcode 9 python -m pytest "$@" -x -q >> $LOG 2>&1
  echo "exit $?" >> $LOG
}
run tests/test_gpu_backward.py -k "training_iteration_gradients_match_reference or render_backward_matches_oracle"
run tests/test_gpu_backward.py -k "trainer or fused_training_loss"
run tests/test_gpu_ops.py -k "mlp or spline"
run tests/test_gpu_render.py -k "tone_mappers or binned or identical_samples or free_running or fuses_compositing or graph_forward"
grep -n "^###\|^exit\|ERROR SUMMARY\|passed\|failed" $LOG
