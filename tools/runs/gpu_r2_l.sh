#!/bin/bash
mkdir -p gpurun_out
python tools/update_prof.py > gpurun_out/update_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fused_update -s 2 -c 2 -f -o gpurun_out/fused_update_r02 python tools/update_prof.py > gpurun_out/ncu_update.log 2>&1
echo "ncu exit $?"; tail -n 3 gpurun_out/ncu_update.log
