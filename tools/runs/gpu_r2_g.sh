#!/bin/bash
mkdir -p gpurun_out
for mode in head deep head deep; do
UNETK_FUSE_BN_REDUCE=$mode timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-inference > gpurun_out/bench_$mode.json 2> gpurun_out/bench_$mode.err; echo "bench $mode exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$mode.json"))
print("$mode", d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["loss"])
PY
done
UNETK_FUSE_BN_REDUCE=deep UNETK_FUSE_DEEP_PIXELS=16384 timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-inference > gpurun_out/bench_deep16k.json 2> gpurun_out/bench_deep16k.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_deep16k.json"))
print("deep16k", d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["loss"])
PY
