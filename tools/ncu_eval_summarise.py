"""Per-kernel time and DRAM traffic of the evaluation-side / input-side kernels (second pass of tools/ncu_eval.py),
from the ncu metric CSV written by tools/gpu_ncu_eval.sh  ->  profiles/ncu_eval_<tag>_summary.txt"""
import collections
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
rows = list(csv.reader(open(os.path.join(ROOT, "gpurun_out", f"eval_kernels_{tag}.csv"), errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ki, mi, ui, vi, idi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value"), h.index("ID")
launch = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    d = launch.setdefault(r[idi], {"name": r[ki].replace("void ", "").replace("unetk::", "").replace("<unnamed>::", "").replace("unnamed>::", "").replace("(anonymous namespace)::", "").split("(")[0].split("<")[0]})
    v = float(r[vi].replace(",", ""))
    v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[ui], 1)   # bytes / ns
    d[r[mi]] = v
ls = list(launch.values())
ls = ls[len(ls) // 2:]           # the second pass
agg = collections.OrderedDict()
for d in ls:
    a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    a[3] = max(a[3], d.get("lts__t_sectors.avg.pct_of_peak_sustained_elapsed", 0.0))
    a[4] = max(a[4], d.get("sm__throughput.avg.pct_of_peak_sustained_elapsed", 0.0))
lines = ["# ncu (gpu__time_duration, dram__bytes_read/write, --clock-control none; cold-cache, serialised launches) of the evaluation-side and",
         "# input-side kernels, second pass of tools/ncu_eval.py: validate_epoch's reductions on 64 maps of 512^2 (pixel_confusion; select_*:",
         "# the percentile), SegmentationMetrics.update + the segmentation losses on 8 x 4 x 512^2 logits, SSIM / anomaly score on 16 x 3 x 256^2,",
         "# the input transforms on 16 images of 1024^2 -> 256^2 (train: resize, flip/rotate, four jitter stages, to-tensor).",
         "# launches   time_us   dram_MB   dram_GB/s   L2 %peak  SM %peak   kernel"]
for k, (n, t, b, l2, smp) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{n:9d} {t/1e3:9.1f} {b/1e6:9.1f} {b/t if t else 0:10.0f} {l2:9.1f} {smp:9.1f}   {k}")
out = os.path.join(ROOT, "profiles", f"ncu_eval_{tag}_summary.txt")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
