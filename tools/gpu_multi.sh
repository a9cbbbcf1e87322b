#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/checks/ddp_check.py > gpurun_out/ddp_check_$N.log 2>&1; echo "ddp_check exit $?"
grep "DDP CHECK" gpurun_out/ddp_check_$N.log || tail -n 20 gpurun_out/ddp_check_$N.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 tests/checks/ddp_check.py --fused > gpurun_out/ddp_check_fused_$N.log 2>&1; echo "ddp_check --fused exit $?"
grep "DDP CHECK" gpurun_out/ddp_check_fused_$N.log || tail -n 20 gpurun_out/ddp_check_fused_$N.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err; echo "bench exit $?"
cut -c1-700 gpurun_out/bench_$N.json; grep -o '"inference": {[^}]*}' gpurun_out/bench_$N.json; tail -n 3 gpurun_out/bench_$N.err
if [ "${GRAPH_DDP_TRIAL:-0}" = "1" ]; then
  UNETK_GRAPH_DDP=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --no-inference > gpurun_out/bench_${N}_graphddp.json 2> gpurun_out/bench_${N}_graphddp.err; echo "bench (backward graph incl. NCCL) exit $?"
  cut -c1-400 gpurun_out/bench_${N}_graphddp.json; grep -iE "warn|error" gpurun_out/bench_${N}_graphddp.err | head -5
  UNETK_GRAPH_DDP=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tests/checks/ddp_check.py > gpurun_out/ddp_check_${N}_graphddp.log 2>&1; echo "ddp_check (graph ddp) exit $?"
  grep "DDP CHECK" gpurun_out/ddp_check_${N}_graphddp.log || tail -n 8 gpurun_out/ddp_check_${N}_graphddp.log
fi
