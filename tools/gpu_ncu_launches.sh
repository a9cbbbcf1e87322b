#!/bin/bash
# Launch list of the bench command only (one metric, no replay).  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
export UNETK_GRAPHS=0     # ncu cannot profile launches recorded during stream capture: eager launches for profiling
R=${1:-r01c}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-inference"
timeout 600 $CMD > gpurun_out/ncu_plain.log 2>&1 || { echo "plain bench run failed"; tail -n 20 gpurun_out/ncu_plain.log; exit 1; }
echo "plain bench ok"
SECONDS=0
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/launches_$R.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $? in ${SECONDS}s ($(wc -l < gpurun_out/launches_$R.csv) lines)"
