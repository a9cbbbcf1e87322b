#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -p no:cacheprovider -k "head" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -n 3 gpurun_out/pytest_gpu.log
timeout 900 python tools/layer_bench.py > gpurun_out/layer_bench_r02.log 2>&1; echo "layer_bench exit $?"
grep -E "head_|train step" gpurun_out/layer_bench_r02.log
for i in 1 2; do
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-inference > gpurun_out/bench_$i.json 2> gpurun_out/bench_$i.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$i.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["loss"])
PY
done
