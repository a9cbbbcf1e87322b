#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -n 12 gpurun_out/pytest_gpu.log
for cfg in "1" "0"; do
  UNETK_GRAPHS=$cfg ONLY_STEP=1 timeout 600 python tools/layer_bench.py > gpurun_out/lb_graph$cfg.log 2>&1
  echo "graphs=$cfg: $(grep -E 'train step|eval forward|enqueue|segmentation train' gpurun_out/lb_graph$cfg.log | tr '\n' ' ')"
  grep -iE "warn|error|Traceback" gpurun_out/lb_graph$cfg.log | head -5
done
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cat gpurun_out/bench.json | cut -c1-900; grep -iE "warn|error" gpurun_out/bench.err | head -5
