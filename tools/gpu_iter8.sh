#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ops_summary.log
bash tools/gpu_ops_tests.sh > /dev/null 2>&1
grep -E "===|passed|failed" gpurun_out/ops_summary.log
grep -h "^E " gpurun_out/ops_*.log | head -20
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q --tb=short -s -p no:cacheprovider > gpurun_out/model_all.log 2>&1
echo "model tests exit $?"; grep -E "passed|failed|^E " gpurun_out/model_all.log | tail -n 8
for cfg in "1" "0" "1" "0"; do
  UNETK_PDL=$cfg ONLY_STEP=1 timeout 600 python tools/layer_bench.py > gpurun_out/lb_pdl$cfg.log 2>&1
  echo "pdl=$cfg: $(grep -E 'train step|eval forward|enqueue' gpurun_out/lb_pdl$cfg.log | tr '\n' ' ')"
done
UNETK_PDL=1 UNETK_PARALLEL_DECODERS=0 UNETK_OVERLAP_WGRAD=0 ONLY_STEP=1 timeout 600 python tools/layer_bench.py > gpurun_out/lb_pdl_single.log 2>&1
echo "pdl=1 single stream: $(grep -E 'train step' gpurun_out/lb_pdl_single.log)"
UNETK_PDL=0 UNETK_PARALLEL_DECODERS=0 UNETK_OVERLAP_WGRAD=0 ONLY_STEP=1 timeout 600 python tools/layer_bench.py > gpurun_out/lb_nopdl_single.log 2>&1
echo "pdl=0 single stream: $(grep -E 'train step' gpurun_out/lb_nopdl_single.log)"
