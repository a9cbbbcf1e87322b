#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py > gpurun_out/ddp_check_$N.log 2>&1; echo "ddp_check exit $?"
grep "DDP CHECK" gpurun_out/ddp_check_$N.log || tail -n 20 gpurun_out/ddp_check_$N.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err; echo "bench exit $?"
cat gpurun_out/bench_$N.json; tail -n 5 gpurun_out/bench_$N.err
