#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_ops_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -n 8 gpurun_out/pytest_gpu.log
for i in 1 2; do
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-inference > gpurun_out/bench_$i.json 2> gpurun_out/bench_$i.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$i.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["loss"])
PY
done
UNETK_FUSED_STEP=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-inference > gpurun_out/bench_unfused.json 2> gpurun_out/bench_unfused.err; echo "bench unfused exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_unfused.json"))
print("unfused", d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["loss"])
PY
timeout 300 python tools/timeline.py --out gpurun_out/timeline_r02f.txt > gpurun_out/timeline.log 2>&1; echo "timeline exit $?"
head -n 2 gpurun_out/timeline.log | cut -c1-250
