#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -n 2 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["clocks"])
print(d["inference"]["value"], d["inference"]["e2e"], d["inference"]["e2e_one_batch_serial"])
print(d["cpu_baseline"])
PY
