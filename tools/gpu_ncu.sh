#!/bin/bash
# Round profile: launch list of the bench command + `ncu --set full` of the dominant tcgen05 kernels over
# ONE training step's launches (tools/ncu_step.py, second step).  Every ncu pass runs only after the same
# command exited 0 without ncu.  Reports are exported to CSV on the box (gpurun_out/ is capped at 64 MiB);
# one single-launch report with source is kept.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
export UNETK_GRAPHS=0     # ncu cannot profile launches recorded during stream capture: eager launches for profiling
R=${1:-r01b}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-inference"
timeout 600 $CMD > gpurun_out/ncu_plain.log 2>&1 || { echo "plain bench run failed"; tail -n 20 gpurun_out/ncu_plain.log; exit 1; }
echo "plain bench ok"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/launches_$R.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $?"
STEP="python tools/ncu_step.py"
timeout 300 $STEP > gpurun_out/ncu_step_plain.log 2>&1 || { echo "plain step run failed"; tail -n 20 gpurun_out/ncu_step_plain.log; exit 1; }
full() {  # name, mangled-name regex, launches per step
  timeout 900 ncu --set full --clock-control none --kernel-name-base mangled -k "regex:$2" --launch-skip $3 -c $3 \
    -f -o /tmp/$1_$R $STEP > gpurun_out/ncu_$1.log 2>&1
  echo "ncu full $1 exit $?"
  ncu -i /tmp/$1_$R.ncu-rep --page raw --csv > gpurun_out/$1_${R}_raw.csv 2>/dev/null
  echo "  $(wc -c < gpurun_out/$1_${R}_raw.csv) bytes of csv"
}
full igemm256 'igemm_kernelILi256ELi1ELi2ELb0' 43
full wgrad128 'wgrad_kernelILi0ELi128ELb0ELi2' 20
full igemm64 'igemm_kernelILi64ELi2ELi2ELb0' 11
# one launch with source correlation, kept as a report: the widest mid layer (first igemm<256> launch of the step)
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k 'regex:igemm_kernelILi256ELi1ELi2ELb0' \
  --launch-skip 45 -c 1 -f -o gpurun_out/igemm256_one_$R $STEP > gpurun_out/ncu_one.log 2>&1
echo "single-launch report exit $? ($(ls -la gpurun_out/igemm256_one_$R.ncu-rep 2>/dev/null | awk '{print $5}') bytes)"
du -sh gpurun_out
