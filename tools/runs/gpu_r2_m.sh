#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -x -q -m gpu -p no:cacheprovider -k "fused_update" > gpurun_out/pytest_gpu_m.log 2>&1; echo "pytest exit $?"
tail -n 3 gpurun_out/pytest_gpu_m.log
timeout 300 python tools/update_bench.py 2>&1 | tee gpurun_out/update_bench.log
for mode in 1 0; do
UNETK_FUSED_UPDATE=$mode timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-inference > gpurun_out/bench_upd$mode.json 2> gpurun_out/bench_upd$mode.err; echo "bench (update=$mode) exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_upd$mode.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["loss"], d["gpu_launches"])
PY
done
timeout 300 python tools/timeline.py --out gpurun_out/timeline_r02_update.txt > gpurun_out/timeline_update.log 2>&1; echo "timeline exit $?"
head -n 3 gpurun_out/timeline_update.log | cut -c1-220
grep -n "fused_update" gpurun_out/timeline_r02_update.txt
