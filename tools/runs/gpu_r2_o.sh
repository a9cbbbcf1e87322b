#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_ops.log 2>&1; echo "pytest ops exit $?"
tail -n 3 gpurun_out/pytest_ops.log
ONLY_STEP=0 timeout 900 python tools/layer_bench.py > gpurun_out/layer_bench_o.log 2>&1; echo "layer_bench exit $?"
grep -E "^(inc|down|up)[0-9.a-z]* |sum conv|train step|eval forward" gpurun_out/layer_bench_o.log | cut -c1-210
for i in 1 2; do
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-inference > gpurun_out/bench_o$i.json 2> gpurun_out/bench_o$i.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_o$i.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["loss"], d["clocks"])
PY
done
