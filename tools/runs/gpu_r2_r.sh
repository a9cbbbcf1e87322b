#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -x -q -m gpu -p no:cacheprovider -k "predict_batches or config4" > gpurun_out/pytest_gpu_r.log 2>&1; echo "pytest exit $?"
tail -n 5 gpurun_out/pytest_gpu_r.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r.json 2> gpurun_out/bench_r.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_r.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
print(d["inference"])
PY
