#!/bin/bash
# A/B: NCCL channel count (SM footprint of the all-reduce kernels beside the backward pass)
N=${1:-4}
mkdir -p gpurun_out
run() {
  env $1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $N --steps 60 --warmup 5 --no-inference --no-cpu-baseline > gpurun_out/bench_nccl.json 2> gpurun_out/bench_nccl.err; rc=$?
  python - "$1" $rc <<PY
import json, sys
d=json.load(open("gpurun_out/bench_nccl.json"))
print(f"{sys.argv[1]:32s} rc={sys.argv[2]} {d['ms_per_step']:.3f} ms  {d['value']:.1f} img/s  e2e {d['e2e']['value']:.1f}")
PY
}
run NCCL_DEBUG=WARN 29541
run NCCL_MAX_NCHANNELS=4 29542
run NCCL_MAX_NCHANNELS=2 29543
run NCCL_DEBUG=WARN 29544
run NCCL_MAX_NCHANNELS=4 29545
run NCCL_MAX_NCHANNELS=2 29546
