#!/bin/bash
# Runs the operator-level GPU tests group by group, each in its own process with a timeout, so a
# trap in one kernel neither hangs the box nor poisons the CUDA context of the other groups.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/ops_gpuinfo.log 2>&1
for k in "${@:-conv3x3_fwd conv3x3_dgrad conv3x3_wgrad convt2x2 first_conv bn_relu_pool head adam}"; do
  for g in $k; do
    echo "=== $g" | tee -a gpurun_out/ops_summary.log
    timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "$g" --tb=short -p no:cacheprovider > gpurun_out/ops_$g.log 2>&1
    echo "exit $?" | tee -a gpurun_out/ops_summary.log
    tail -n 3 gpurun_out/ops_$g.log | tee -a gpurun_out/ops_summary.log
  done
done
