#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py -x -q -m gpu -p no:cacheprovider -k "fused" > gpurun_out/pytest_gpu_m.log 2>&1; echo "pytest exit $?"
tail -n 3 gpurun_out/pytest_gpu_m.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tests/checks/ddp_check.py --update > gpurun_out/ddp_check_update_$N.log 2>&1; echo "ddp_check --update exit $?"
grep "DDP CHECK" gpurun_out/ddp_check_update_$N.log || tail -n 20 gpurun_out/ddp_check_update_$N.log | cut -c1-300
for mode in 1 0 1 0; do
UNETK_FUSED_UPDATE=$mode timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$mode bench.py --gpus $N --steps 60 --warmup 5 --no-inference > gpurun_out/bench_${N}_upd$mode.json 2> gpurun_out/bench_${N}_upd$mode.err; echo "bench (update=$mode) exit $?"
cut -c1-200 gpurun_out/bench_${N}_upd$mode.json
done
for mode in 1 0 1 0; do
UNETK_FUSED_UPDATE=$mode timeout 400 python bench.py --steps 100 --warmup 10 --no-inference --no-cpu-baseline > gpurun_out/bench_1_upd$mode.json 2> gpurun_out/bench_1_upd$mode.err; echo "bench 1 GPU (update=$mode) exit $?"
cut -c1-200 gpurun_out/bench_1_upd$mode.json
done
