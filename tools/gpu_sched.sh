#!/bin/bash
# scheduling experiments: whole-step time under different stream policies (each its own process)
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" ONLY_STEP=1 timeout 300 python tools/layer_bench.py > gpurun_out/lb_$name.log 2>&1
  echo "$name: $(grep -E 'train step' gpurun_out/lb_$name.log | head -1 | cut -c1-60)"
}
run base UNETK_PRIORITY_STREAMS=0 UNETK_GRAD_SLOTS=2
run prio_s2 UNETK_PRIORITY_STREAMS=1 UNETK_GRAD_SLOTS=2
run prio_s4 UNETK_PRIORITY_STREAMS=1 UNETK_GRAD_SLOTS=4
run prio_s4_early UNETK_PRIORITY_STREAMS=1 UNETK_GRAD_SLOTS=4 UNETK_WGRAD_EARLY=1
run base_s4 UNETK_PRIORITY_STREAMS=0 UNETK_GRAD_SLOTS=4
run prio_s6 UNETK_PRIORITY_STREAMS=1 UNETK_GRAD_SLOTS=6
run base_b UNETK_PRIORITY_STREAMS=0 UNETK_GRAD_SLOTS=2
run prio_s4_b UNETK_PRIORITY_STREAMS=1 UNETK_GRAD_SLOTS=4
timeout 300 python tools/timeline.py --out gpurun_out/timeline_r02c.txt > gpurun_out/timeline.log 2>&1; echo "timeline exit $?"
head -n 2 gpurun_out/timeline.log | cut -c1-250
timeout 600 python -m pytest tests/test_model_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -n 4 gpurun_out/pytest_gpu.log
