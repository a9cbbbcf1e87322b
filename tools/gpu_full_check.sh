#!/bin/bash
# What the driver runs at round end: GPU tests, smoke, bench (both arms).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee gpurun_out/full_summary.log
tail -n 15 gpurun_out/pytest_gpu.log | tee -a gpurun_out/full_summary.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/full_summary.log
tail -n 3 gpurun_out/smoke.log | tee -a gpurun_out/full_summary.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "reference arm exit $?" | tee -a gpurun_out/full_summary.log
cat gpurun_out/bench_ref.json | tee -a gpurun_out/full_summary.log; tail -n 3 gpurun_out/bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" | tee -a gpurun_out/full_summary.log
cat gpurun_out/bench.json | tee -a gpurun_out/full_summary.log; tail -n 5 gpurun_out/bench.err
timeout 900 python tools/layer_bench.py > gpurun_out/layer_bench.log 2>&1; echo "layer_bench exit $?"
tail -n 12 gpurun_out/layer_bench.log
