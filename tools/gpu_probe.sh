#!/bin/bash
mkdir -p gpurun_out
timeout 300 ./tools/umma_probe > gpurun_out/umma_probe.txt 2>&1; echo "probe exit $?"
tail -n 3 gpurun_out/umma_probe.txt
timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu -k "checkpoint" -s -p no:cacheprovider 2>&1 | grep -E "parity|passed|failed"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
