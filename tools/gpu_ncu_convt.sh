#!/bin/bash
mkdir -p gpurun_out
W=${1:-fwd}
WHICH=$W python tools/convt_prof.py > gpurun_out/convt_prof_plain.log 2>&1 &&
WHICH=$W ncu --set full --clock-control none --import-source on -k regex:"igemm|wgrad" -s 3 -c 3 -f -o gpurun_out/convt_${W}_r02 python tools/convt_prof.py > gpurun_out/ncu_convt.log 2>&1
echo "ncu exit $?"; tail -n 2 gpurun_out/ncu_convt.log
