#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_seg_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -n 6 gpurun_out/pytest_gpu.log
run() {  # name, env...
  name=$1; shift
  env "$@" ONLY_STEP=1 timeout 300 python tools/layer_bench.py > gpurun_out/lb_$name.log 2>&1
  echo "$name: $(grep -E 'train step' gpurun_out/lb_$name.log | head -1 | cut -c1-60)"
}
run a
run b
run noprio UNETK_PRIORITY_STREAMS=0 UNETK_GRAD_SLOTS=2
timeout 300 python tools/timeline.py --out gpurun_out/timeline_r02e.txt > gpurun_out/timeline.log 2>&1; echo "timeline exit $?"
head -n 2 gpurun_out/timeline.log | cut -c1-250
export UNETK_GRAPHS=0
timeout 300 python tools/ncu_step.py > gpurun_out/ncu_step_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:head_ --launch-skip 4 -c 4 -o gpurun_out/head_r02 -f python tools/ncu_step.py > gpurun_out/ncu_head.log 2>&1
echo "ncu exit $?"; tail -n 3 gpurun_out/ncu_head.log
