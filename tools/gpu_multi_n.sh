#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tests/checks/ddp_check.py --update > gpurun_out/ddp_check_update_$N.log 2>&1; echo "ddp_check --update exit $?"
grep "DDP CHECK" gpurun_out/ddp_check_update_$N.log || tail -n 20 gpurun_out/ddp_check_update_$N.log | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 40 --warmup 5 > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$N.json"))
print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"], d["inference"]["value"], d["inference"]["e2e"])
PY
timeout 400 python bench.py --gpus 1 --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_1_same_box.json 2> gpurun_out/bench_1_same_box.err; echo "bench N=1 exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_1_same_box.json"))
print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"], d["inference"]["value"], d["inference"]["e2e"])
PY
