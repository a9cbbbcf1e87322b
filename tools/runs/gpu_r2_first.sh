#!/bin/bash
# round 2, first GPU call: the changed paths (advisor fixes), smoke, the step timeline and a short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_ops_gpu.py -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -n 8 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
tail -n 3 gpurun_out/smoke.log
timeout 300 python tools/timeline.py --out gpurun_out/timeline_r02a.txt > gpurun_out/timeline.log 2>&1; echo "timeline exit $?"
head -n 30 gpurun_out/timeline.log
timeout 300 python tools/timeline.py --eager --out gpurun_out/timeline_r02a_eager.txt > gpurun_out/timeline_eager.log 2>&1; echo "timeline eager exit $?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
cut -c1-600 gpurun_out/bench.json
