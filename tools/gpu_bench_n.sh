#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err; echo "bench exit $?"
cut -c1-330 gpurun_out/bench_$N.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench_$N.json | head -1; grep -o '"inference": {[^}]*}' gpurun_out/bench_$N.json | cut -c1-200; grep -o '"clocks": {[^}]*}' gpurun_out/bench_$N.json; tail -n 3 gpurun_out/bench_$N.err
