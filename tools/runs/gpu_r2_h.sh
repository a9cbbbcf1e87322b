#!/bin/bash
mkdir -p gpurun_out
for mode in 0 1 2 0 1 2; do
UNETK_TMA_STORE=$mode timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-inference > gpurun_out/bench_tma$mode.json 2> gpurun_out/bench_tma$mode.err; echo "bench tma=$mode exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_tma$mode.json"))
print("tma=$mode", d["value"], d["ms_per_step"], d["e2e"]["value"], d["loss"])
PY
done
timeout 900 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -n 3 gpurun_out/pytest_gpu.log
