#!/bin/bash
# Network-level GPU tests (each group in its own process) and the per-layer timing tool.
mkdir -p gpurun_out
for g in eval_forward train_step loss_kernels loss_curve checkpoint full_size; do
  echo "=== $g" | tee -a gpurun_out/model_summary.log
  timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -k "$g" --tb=short -s -p no:cacheprovider > gpurun_out/model_$g.log 2>&1
  echo "exit $?" | tee -a gpurun_out/model_summary.log
  grep -E "parity|passed|failed|Error|error|assert" gpurun_out/model_$g.log | tail -n 12 | tee -a gpurun_out/model_summary.log
done
timeout 900 python tools/layer_bench.py > gpurun_out/layer_bench.log 2>&1
echo "layer_bench exit $?" | tee -a gpurun_out/model_summary.log
tail -n 5 gpurun_out/layer_bench.log
