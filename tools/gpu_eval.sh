#!/bin/bash
# device-side validate_epoch reductions: parity tests + their measured bandwidth line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_eval_gpu.py -q -s -m gpu -p no:cacheprovider > gpurun_out/pytest_eval.log 2>&1; echo "pytest exit $?"
grep -E "pixel_confusion [0-9]|passed|failed|Error|error" gpurun_out/pytest_eval.log | head -40
tail -n 30 gpurun_out/pytest_eval.log
