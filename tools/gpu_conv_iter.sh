#!/bin/bash
# Conv-kernel iteration loop: operator tests for the tensor-core kernels, then the per-layer timing.
mkdir -p gpurun_out
rm -f gpurun_out/ops_summary.log
bash tools/gpu_ops_tests.sh "conv3x3_fwd conv3x3_dgrad conv3x3_wgrad convt2x2" > /dev/null 2>&1
cat gpurun_out/ops_summary.log
if grep -q "failed\|exit 1\|exit 124\|rror" gpurun_out/ops_summary.log; then
  grep -h "^E " gpurun_out/ops_conv*.log gpurun_out/ops_convt2x2.log | head -30
  echo "SKIPPING bench: tests failed"
  exit 1
fi
timeout 900 python tools/layer_bench.py > gpurun_out/layer_bench.log 2>&1; echo "layer_bench exit $?"
grep -E "sum conv|train step|eval forward|enqueue" gpurun_out/layer_bench.log
