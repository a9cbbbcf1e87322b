#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ops_summary.log
bash tools/gpu_ops_tests.sh "convt2x2 bn_relu_pool conv3x3_fwd" > /dev/null 2>&1
grep -E "===|passed|failed" gpurun_out/ops_summary.log
grep -h "^E " gpurun_out/ops_*.log | head -20
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q --tb=short -s -p no:cacheprovider > gpurun_out/model_all.log 2>&1
echo "model tests exit $?"; grep -E "passed|failed|^E " gpurun_out/model_all.log | tail -n 8
timeout 900 python tools/layer_bench.py > gpurun_out/layer_bench.log 2>&1; echo "layer_bench exit $?"
grep -E "sum conv|up..up|bn_bwd|bnrelu|maxpool|train step|eval forward|enqueue" gpurun_out/layer_bench.log
