#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "bn_relu_pool" -p no:cacheprovider 2>&1 | tail -n 2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tests/checks/ddp_check.py --sync-bn > gpurun_out/ddp_check_syncbn_$N.log 2>&1; echo "ddp_check --sync-bn exit $?"
grep "DDP CHECK" gpurun_out/ddp_check_syncbn_$N.log || tail -n 25 gpurun_out/ddp_check_syncbn_$N.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 tests/checks/ddp_check.py > gpurun_out/ddp_check_$N.log 2>&1; echo "ddp_check exit $?"
grep "DDP CHECK" gpurun_out/ddp_check_$N.log || tail -n 25 gpurun_out/ddp_check_$N.log
