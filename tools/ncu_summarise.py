"""Summarise `ncu --set full` reports (read with `ncu -i ... --page raw --csv`) into a text table and
profiles/ncu_traffic.json (DRAM bytes per launch per kernel, used by bench.py's roofline.traffic).

    python tools/ncu_summarise.py r01b igemm256 wgrad128 igemm64
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_tensor.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.avg.pct_of_peak_sustained_elapsed",
           "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active"]
KERNEL_KEY = {"igemm256": "igemm_kernel<256,1,2>", "wgrad128": "wgrad_kernel<0,128,0,2>", "igemm64": "igemm_kernel<64,2,2>"}


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    tag = sys.argv[1]
    out_lines = []
    traffic = {}
    for name in sys.argv[2:]:
        rep = os.path.join(ROOT, "gpurun_out", f"{name}_{tag}_raw.csv")     # exported on the GPU box (tools/gpu_ncu.sh)
        if not os.path.exists(rep):
            out_lines.append(f"# {name}: no report")
            continue
        raw = open(rep, errors="ignore").read()
        rows = list(csv.reader(io.StringIO(raw)))
        hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
        names, units = rows[hdr], rows[hdr + 1]
        col = {n: i for i, n in enumerate(names)}
        out_lines.append(f"# {name} ({rep.split('/')[-1]}): ncu --set full --clock-control none, one training step's launches "
                         f"of {KERNEL_KEY.get(name, name)} (B=16 @256^2), in launch order")
        out_lines.append("# dur_us tensor_pipe% dram_rd_MB dram_wr_MB dram% l2% sm% grid regs  kernel")
        tot_b = tot_t = 0.0
        n = 0
        tp = []
        for r in rows[hdr + 2:]:
            if len(r) < len(names):
                continue
            def g(m):
                return r[col[m]] if m in col else "nan"
            dur = float(g("gpu__time_duration.sum").replace(",", ""))
            du = units[col["gpu__time_duration.sum"]]
            dur_us = dur * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(du, 1)
            rd = to_bytes(g("dram__bytes_read.sum"), units[col["dram__bytes_read.sum"]])
            wr = to_bytes(g("dram__bytes_write.sum"), units[col["dram__bytes_write.sum"]])
            t = float(g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed").replace(",", ""))
            out_lines.append(f"{dur_us:8.1f} {t:6.1f} {rd/1e6:9.2f} {wr/1e6:9.2f} "
                             f"{float(g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):6.1f} "
                             f"{float(g('lts__t_sectors.avg.pct_of_peak_sustained_elapsed')):6.1f} "
                             f"{float(g('sm__throughput.avg.pct_of_peak_sustained_elapsed')):6.1f} "
                             f"{g('launch__grid_size'):>5s} {g('launch__registers_per_thread'):>4s}  {r[col['Kernel Name']][:60]}")
            tot_b += rd + wr
            tot_t += dur_us
            tp.append((t, dur_us))
            n += 1
        if n:
            wtp = sum(a * b for a, b in tp) / tot_t
            out_lines.append(f"# {n} launches, {tot_t:.1f} us, time-weighted tensor pipe active {wtp:.1f} %, "
                             f"DRAM {tot_b/1e6:.1f} MB total = {tot_b/n/1e6:.2f} MB per launch")
            traffic[KERNEL_KEY.get(name, name)] = {"dram_bytes_per_launch": tot_b / n, "launches": n, "us_total_under_ncu": tot_t,
                                                   "tensor_pipe_active_pct_time_weighted": wtp, "source": f"ncu --set full, raw page csv {name}_{tag}_raw.csv"}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", f"ncu_full_{tag}_summary.txt"), "w") as f:
        f.write("\n".join(out_lines) + "\n")
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    print("\n".join(out_lines))


if __name__ == "__main__":
    main()
