#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py -x -q -m gpu -p no:cacheprovider -k "fused_update" > gpurun_out/pytest_gpu_m.log 2>&1; echo "pytest exit $?"
tail -n 2 gpurun_out/pytest_gpu_m.log
for flag in "" "--fused" "--update"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tests/checks/ddp_check.py $flag > gpurun_out/ddp_check_x_$N.log 2>&1; echo "ddp_check $flag exit $?"
grep "DDP CHECK" gpurun_out/ddp_check_x_$N.log | cut -c1-200 || tail -n 20 gpurun_out/ddp_check_x_$N.log | cut -c1-300
done
for rep in 1 2; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$rep bench.py --gpus $N --steps 60 --warmup 5 --no-inference --no-cpu-baseline > gpurun_out/bench_$N.json 2> gpurun_out/bench_$N.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$N.json"))
print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"])
PY
done
timeout 400 python bench.py --gpus 1 --steps 60 --warmup 5 --no-cpu-baseline --no-inference > gpurun_out/bench_1_same_box.json 2> gpurun_out/bench_1_same_box.err; echo "bench N=1 exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_1_same_box.json"))
print(d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"])
PY
