#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_seg_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -n 6 gpurun_out/pytest_gpu.log
run() {  # name, env...
  name=$1; shift
  env "$@" ONLY_STEP=1 timeout 300 python tools/layer_bench.py > gpurun_out/lb_$name.log 2>&1
  echo "$name: $(grep -E 'train step' gpurun_out/lb_$name.log | head -1 | cut -c1-60)"
}
run nofuse UNETK_HEAD_FUSE_BN=0
run fuse UNETK_HEAD_FUSE_BN=1
run nofuse_b UNETK_HEAD_FUSE_BN=0
run fuse_b UNETK_HEAD_FUSE_BN=1
run fuse_noprio UNETK_HEAD_FUSE_BN=1 UNETK_PRIORITY_STREAMS=0 UNETK_GRAD_SLOTS=2
timeout 300 python tools/timeline.py --out gpurun_out/timeline_r02d.txt > gpurun_out/timeline.log 2>&1; echo "timeline exit $?"
head -n 2 gpurun_out/timeline.log | cut -c1-250
