#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -n 6 gpurun_out/pytest_gpu.log
for i in 1 2; do
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-inference > gpurun_out/bench_$i.json 2> gpurun_out/bench_$i.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$i.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
PY
done
