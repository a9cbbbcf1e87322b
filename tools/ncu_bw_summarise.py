"""Per-kernel DRAM traffic and achieved GB/s of the bandwidth-bound kernels of one training step, from
the ncu metric CSV written by tools/gpu_ncu_bw.sh  ->  profiles/ncu_bandwidth_<tag>_summary.txt"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01b"
rows = list(csv.reader(open(os.path.join(ROOT, "gpurun_out", f"bw_kernels_{tag}.csv"), errors="ignore")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ki, mi, ui, vi, idi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value"), h.index("ID")
launch = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    d = launch.setdefault(r[idi], {"name": r[ki].replace("void ", "").replace("unetk::", "").split("(")[0]})
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)   # bytes / ns
    d[r[mi]] = v
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6543.1) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6543.1
agg = collections.OrderedDict()
for d in launch.values():
    a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    a[3] = max(a[3], (d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)) / max(d.get("gpu__time_duration.sum", 1.0), 1.0))
lines = [f"# ncu (gpu__time_duration, dram__bytes_read/write, --clock-control none) of the bandwidth-bound launches of ONE AnomalyUNet training",
         f"# step (B=16 @256^2, tools/ncu_step.py step 2, graphs off).  GB/s = DRAM bytes / kernel time; peak = measured copy bandwidth {peak:.0f} GB/s.",
         f"# Deep layers' tensors (<= 33 MB) stay in the 126 MB L2, so the per-kernel AVERAGE understates the full-resolution launches: see 'best launch'.",
         "# launches   time_us   dram_MB   avg_GB/s  frac   best_launch_GB/s  frac   kernel"]
for k, (n, t, b, best) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    gbs = b / t if t else 0.0
    lines.append(f"{n:9d} {t/1e3:9.1f} {b/1e6:9.1f} {gbs:9.0f} {gbs/peak:6.2f} {best:12.0f} {best/peak:11.2f}   {k}")
out = os.path.join(ROOT, "profiles", f"ncu_bandwidth_{tag}_summary.txt")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
