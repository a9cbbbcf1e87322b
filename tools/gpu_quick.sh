#!/bin/bash
# quick regression: network-level GPU tests + whole-step timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -n 6 gpurun_out/pytest_gpu.log
for i in 1 2; do
ONLY_STEP=1 timeout 600 python tools/layer_bench.py > gpurun_out/lb.log 2>&1
echo "$(grep -E 'train step|eval forward|enqueue' gpurun_out/lb.log | tr '\n' ' ' | cut -c1-330)"
done
