#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ops_summary.log
bash tools/gpu_ops_tests.sh "first_conv conv3x3_fwd" > /dev/null 2>&1
grep -E "===|passed|failed" gpurun_out/ops_summary.log
grep -h "^E " gpurun_out/ops_*.log | head -20
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q --tb=short -s -p no:cacheprovider > gpurun_out/model_all.log 2>&1
echo "model tests exit $?"; grep -E "parity|passed|failed|^E " gpurun_out/model_all.log | tail -n 20
for cfg in "1" "0"; do
  UNETK_FIRST_FWD_TC=$cfg ONLY_STEP=1 timeout 600 python tools/layer_bench.py > gpurun_out/lb_fc$cfg.log 2>&1
  echo "first_fwd_tc=$cfg: $(grep -E 'train step|eval forward|enqueue' gpurun_out/lb_fc$cfg.log | tr '\n' ' ')"
done
