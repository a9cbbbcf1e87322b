#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -x -q -m gpu -p no:cacheprovider -k "fused_update" > gpurun_out/pytest_gpu_m.log 2>&1; echo "pytest exit $?"
tail -n 3 gpurun_out/pytest_gpu_m.log
timeout 300 python tools/update_bench.py 2>&1 | tee gpurun_out/update_bench.log | grep -E "max_blocks    0|adam"
