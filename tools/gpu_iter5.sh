#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ops_summary.log
bash tools/gpu_ops_tests.sh "bn_relu_pool first_conv head" > /dev/null 2>&1
grep -E "===|passed|failed" gpurun_out/ops_summary.log
grep -h "^E " gpurun_out/ops_*.log | head -20
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q --tb=short -s -p no:cacheprovider > gpurun_out/model_all.log 2>&1
echo "model tests exit $?"; grep -E "parity|passed|failed|^E " gpurun_out/model_all.log | tail -n 16
for cfg in "1 1" "1 0" "0 1" "0 0"; do
  set -- $cfg
  UNETK_PARALLEL_DECODERS=$1 UNETK_OVERLAP_WGRAD=$2 ONLY_STEP=1 timeout 600 python tools/layer_bench.py > gpurun_out/lb_$1$2.log 2>&1
  echo "parallel_decoders=$1 overlap_wgrad=$2: $(grep -E 'train step|eval forward|enqueue' gpurun_out/lb_$1$2.log | tr '\n' ' ')"
done
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
