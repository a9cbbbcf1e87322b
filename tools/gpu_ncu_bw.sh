#!/bin/bash
# ncu evidence for the bandwidth-bound kernels: DRAM bytes and time of every elementwise / head / loss /
# optimizer launch of ONE training step (tools/ncu_step.py, second step), exported to CSV on the box.
mkdir -p gpurun_out
R=${1:-r01b}
export UNETK_GRAPHS=0
STEP="python tools/ncu_step.py"
timeout 300 $STEP > gpurun_out/ncu_step_plain.log 2>&1 || { echo "plain step run failed"; tail -n 20 gpurun_out/ncu_step_plain.log; exit 1; }
KREG='regex:bnrelu|bn_bwd|maxpool|head_|combined_loss|adam_step|pack_conv|unpack_conv|permute_batched|image_to_nhwc64'
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sectors.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed \
  --clock-control none -k "$KREG" --launch-skip 97 -c 97 --csv --log-file gpurun_out/bw_kernels_$R.csv $STEP > gpurun_out/ncu_bw.log 2>&1
echo "ncu bandwidth kernels exit $? ($(wc -l < gpurun_out/bw_kernels_$R.csv) csv lines)"
