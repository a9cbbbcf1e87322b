#!/bin/bash
# Full round-end check + launch list of the bench command (ncu pass runs after the plain run exits 0).
bash tools/gpu_full_check.sh
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/launches_cur.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu launch list exit $?"
timeout 900 python tools/layer_bench.py > gpurun_out/layer_bench.log 2>&1; echo "layer_bench exit $?"
grep -E "sum conv|train step|eval forward|enqueue" gpurun_out/layer_bench.log
