#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ops_summary.log
bash tools/gpu_ops_tests.sh "conv3x3_dgrad convt2x2 bn_relu_pool head" > /dev/null 2>&1
grep -E "===|passed|failed" gpurun_out/ops_summary.log
grep -h "^E " gpurun_out/ops_*.log | head -20
for g in train_step loss_curve; do
  timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -k "$g" --tb=short -s -p no:cacheprovider > gpurun_out/model_$g.log 2>&1
  echo "model $g exit $?"; grep -E "parity|passed|failed|^E " gpurun_out/model_$g.log | tail -n 8
done
timeout 900 python tools/layer_bench.py > gpurun_out/layer_bench.log 2>&1; echo "layer_bench exit $?"
grep -E "sum conv|train step|eval forward|enqueue" gpurun_out/layer_bench.log
