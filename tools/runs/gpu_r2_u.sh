#!/bin/bash
# A/B under the power cap (100 steps): TMA-store epilogues, BatchNorm-backward sums fused into dgrad epilogues
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
run UNETK_TMA_STORE=0
run UNETK_TMA_STORE=1
run UNETK_TMA_STORE=2
run UNETK_FUSE_BN_REDUCE=deep
done
