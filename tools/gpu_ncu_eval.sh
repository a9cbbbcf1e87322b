#!/bin/bash
# ncu evidence for the evaluation-side / input-side kernels: DRAM bytes and time of every launch of the second pass
mkdir -p gpurun_out
R=${1:-r02}
STEP="python tools/ncu_eval.py"
timeout 300 $STEP > gpurun_out/ncu_eval_plain.log 2>&1 || { echo "plain run failed"; tail -n 20 gpurun_out/ncu_eval_plain.log; exit 1; }
tail -n 1 gpurun_out/ncu_eval_plain.log
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sectors.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed \
  --clock-control none -k "regex:anomaly_score|class_confusion|flip_rotate|gray_sum|jitter_kernel|pixel_confusion|resize_|seg_loss|select_|ssim_|u8_to_tensor" --csv --log-file gpurun_out/eval_kernels_$R.csv $STEP > gpurun_out/ncu_eval.log 2>&1
echo "ncu eval kernels exit $? ($(wc -l < gpurun_out/eval_kernels_$R.csv) csv lines)"
