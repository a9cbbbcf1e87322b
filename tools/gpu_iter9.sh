#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_seg_gpu.py -m gpu -q --tb=short -s -p no:cacheprovider > gpurun_out/seg_tests.log 2>&1
echo "seg tests exit $?"; grep -E "parity|passed|failed|^E |Error" gpurun_out/seg_tests.log | tail -n 30
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q --tb=short -s -p no:cacheprovider -k "loss_curve or train_step" > gpurun_out/model_some.log 2>&1
echo "model tests exit $?"; grep -E "parity|passed|failed|^E " gpurun_out/model_some.log | tail -n 8
