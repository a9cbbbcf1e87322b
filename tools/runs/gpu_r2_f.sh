#!/bin/bash
mkdir -p gpurun_out
for mode in conv all; do
UNETK_FUSE_BN_REDUCE=$mode timeout 600 python tests/checks/fuse_check.py > gpurun_out/fuse_check_$mode.log 2>&1; echo "fuse_check $mode exit $?"
grep "fuse check" gpurun_out/fuse_check_$mode.log | cut -c1-150 || tail -5 gpurun_out/fuse_check_$mode.log
done
for mode in head conv all head conv; do
UNETK_FUSE_BN_REDUCE=$mode timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-inference > gpurun_out/bench_$mode.json 2> gpurun_out/bench_$mode.err; echo "bench $mode exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$mode.json"))
print("$mode", d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["loss"])
PY
done
UNETK_FUSE_BN_REDUCE=conv timeout 300 python tools/timeline.py --out gpurun_out/timeline_r02g.txt > gpurun_out/timeline.log 2>&1; echo "timeline exit $?"
head -n 2 gpurun_out/timeline.log | cut -c1-250
UNETK_FUSE_BN_REDUCE=conv timeout 900 python tools/layer_bench.py > gpurun_out/layer_bench_conv.log 2>&1; echo "layer_bench exit $?"
