#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ops_summary.log
bash tools/gpu_ops_tests.sh > /dev/null 2>&1
grep -E "===|passed|failed" gpurun_out/ops_summary.log
grep -h "^E " gpurun_out/ops_*.log | head -20
timeout 600 python tests/checks/fuse_check.py > gpurun_out/fuse_ab.log 2>&1; echo "fuse check exit $?"; grep -A40 "fuse check" gpurun_out/fuse_ab.log | head -n 60 || tail -n 20 gpurun_out/fuse_ab.log
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q --tb=short -s -p no:cacheprovider > gpurun_out/model_all.log 2>&1
echo "model tests exit $?"; grep -E "parity|passed|failed|^E " gpurun_out/model_all.log | tail -n 16
timeout 900 python tools/layer_bench.py > gpurun_out/layer_bench.log 2>&1; echo "layer_bench exit $?"
cat gpurun_out/layer_bench.log | tail -n 45
timeout 600 python tools/wgrad_sweep.py > gpurun_out/wgrad_sweep.txt 2>&1; echo "wgrad_sweep exit $?"
