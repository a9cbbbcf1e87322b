#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q --tb=short -s -p no:cacheprovider -k "bilinear" > gpurun_out/bilinear.log 2>&1
echo "bilinear test exit $?"; grep -E "parity|passed|failed|^E |Error" gpurun_out/bilinear.log | tail -n 20
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -n 8 gpurun_out/pytest_gpu.log
ONLY_STEP=1 timeout 600 python tools/layer_bench.py > gpurun_out/lb.log 2>&1
echo "$(grep -E 'train step|eval forward|enqueue' gpurun_out/lb.log | tr '\n' ' ')"
