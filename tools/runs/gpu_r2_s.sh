#!/bin/bash
mkdir -p gpurun_out
run() {
  env $1 timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-inference > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; rc=$?
  python - "$1" $rc <<PY
import json, sys
d=json.load(open("gpurun_out/bench_ab.json"))
print(f"{sys.argv[1]:40s} rc={sys.argv[2]} {d['ms_per_step']:.3f} ms  {d['value']:.1f} img/s  e2e {d['e2e']['value']:.1f}  sm {d['clocks']['sm_mhz']} MHz  {d['clocks']['power_w_max']} W")
PY
}
for rep in 1 2; do
run UNETK_EW_BLOCKS_PER_SM=3
run UNETK_EW_BLOCKS_PER_SM=5
run UNETK_EW_BLOCKS_PER_SM=6
run UNETK_EW_BLOCKS_PER_SM=4
done
