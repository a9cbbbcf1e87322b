#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu -p no:cacheprovider -k "prefetch or loss_curve" > gpurun_out/pytest_e2e.log 2>&1; echo "pytest exit $?"
tail -n 5 gpurun_out/pytest_e2e.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-inference --no-cpu-baseline > gpurun_out/bench_e2e.json 2> gpurun_out/bench_e2e.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_e2e.json'))
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks')})
PY
tail -n 3 gpurun_out/bench_e2e.err
