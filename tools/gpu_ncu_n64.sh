#!/bin/bash
# ncu --set full with source correlation of the narrow (N = 64) conv launches: where do the epilogue warps stall?
mkdir -p gpurun_out
timeout 300 python tools/ncu_conv.py n64 > gpurun_out/ncu_conv_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:igemm_kernel --launch-skip 2 -c 2 -f -o gpurun_out/igemm64_r02 python tools/ncu_conv.py n64 > gpurun_out/ncu_n64.log 2>&1
echo "ncu exit $?"; tail -n 2 gpurun_out/ncu_n64.log; ls -la gpurun_out/igemm64_r02.ncu-rep
