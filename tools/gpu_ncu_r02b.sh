#!/bin/bash
# round 2: launch list of the training bench (no inference passes) + DRAM bytes / time of the bandwidth-bound launches of one step
mkdir -p gpurun_out
export UNETK_GRAPHS=0
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-inference"
timeout 600 $CMD > gpurun_out/ncu_plain.log 2>&1 || { echo "plain bench run failed"; tail -n 20 gpurun_out/ncu_plain.log; exit 1; }
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $?"
STEP="python tools/ncu_step.py"
timeout 300 $STEP > gpurun_out/ncu_step_plain.log 2>&1 || { echo "plain step run failed"; exit 1; }
KREG='regex:bnrelu|bn_bwd|maxpool|head_|combined_loss|adam_step|pack_conv|unpack_conv|permute_batched|image_to_nhwc64'
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sectors.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed \
  --clock-control none -k "$KREG" --launch-skip 91 -c 91 --csv --log-file gpurun_out/bw_kernels_r02.csv $STEP > gpurun_out/ncu_bw.log 2>&1
echo "ncu bandwidth kernels exit $? ($(wc -l < gpurun_out/bw_kernels_r02.csv) csv lines)"
