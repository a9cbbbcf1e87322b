#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_preprocess_gpu.py tests/test_eval_gpu.py -q -s -m gpu -p no:cacheprovider > gpurun_out/pytest_pre.log 2>&1; echo "pytest exit $?"
grep -E "eval_image_transform 16|pixel_confusion [0-9]|passed|failed" gpurun_out/pytest_pre.log | head
grep -vE "^Validation|it/s" gpurun_out/pytest_pre.log | tail -n 40
