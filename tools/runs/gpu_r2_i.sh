#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/timeline.py --out gpurun_out/timeline_r02.txt > gpurun_out/timeline.log 2>&1; echo "timeline exit $?"
head -n 2 gpurun_out/timeline.log | cut -c1-250
UNETK_FUSED_STEP=0 timeout 300 python tools/timeline.py --out gpurun_out/timeline_r02_unfused.txt > gpurun_out/timeline_unfused.log 2>&1; echo "timeline (unfused) exit $?"
head -n 2 gpurun_out/timeline_unfused.log | cut -c1-250
timeout 900 python tools/layer_bench.py > gpurun_out/layer_bench_r02.log 2>&1; echo "layer_bench exit $?"
tail -n 12 gpurun_out/layer_bench_r02.log
