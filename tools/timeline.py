"""Kernel-by-kernel, stream-by-stream layout of one graph-replayed training step, without a profiler.

    python tools/timeline.py [--batch 16] [--size 256] [--out gpurun_out/timeline.txt] [--eager]

Every C-ABI call is bracketed by two one-thread stamp kernels (`unetk_debug_stamp`: %globaltimer) on the call's
own stream; the stamps are captured into the engine's CUDA graphs with everything else, so the third step (a
replay) rewrites them with the times of a real replayed step.  "start" is the moment the stream was ready to
launch the kernel (its predecessors on the stream and its event waits done), "end" when it had finished: a
tensor-core kernel that waited for the SMs held by another stream's persistent kernel shows that wait inside
its duration.  The stamps add two tiny launches per call, so absolute times run a few percent long; what the
table is for is WHERE the tensor pipe has nothing to do.
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (UNETK_FUSED_STEP=0 in the environment lays out the unfused model(x) / criterion / backward sequence instead)
TENSOR = ("unetk_conv3x3_fwd", "unetk_conv3x3_dgrad", "unetk_conv3x3_wgrad", "unetk_convt2x2_fwd", "unetk_convt2x2_dgrad",
          "unetk_convt2x2_wgrad", "unetk_first_conv_wgrad_tc")


def union_length(iv):
    iv = sorted(iv)
    total, cur_s, cur_e = 0, None, None
    for s, e in iv:
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                total += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    if cur_e is not None:
        total += cur_e - cur_s
    return total


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "timeline.txt"))
    ap.add_argument("--eager", action="store_true", help="no CUDA graphs (UNETK_GRAPHS=0)")
    args = ap.parse_args()
    if args.eager:
        os.environ["UNETK_GRAPHS"] = "0"
    import torch
    from tiaozhanbei_unet_b200 import _lib, train_utils
    from tiaozhanbei_unet_b200.model import AnomalyUNet

    dev = torch.device("cuda", 0)
    torch.manual_seed(42)
    model = AnomalyUNet(3).to(dev).train()
    crit = train_utils.CombinedLoss()
    opt = train_utils.get_optimizer(model, "adam", 1e-3, 1e-4)
    g = torch.Generator().manual_seed(0)
    image = torch.randn(args.batch, 3, args.size, args.size, generator=g).to(dev)
    mask = (torch.rand(args.batch, 1, args.size, args.size, generator=g) > 0.95).float().to(dev)

    def step():
        train_utils.train_step(model, crit, opt, image, mask)

    step()                                   # eager warm-up (plans, streams)
    torch.cuda.synchronize()
    tl = _lib.timeline = _lib.Timeline(dev)
    step()                                   # captures the graphs, stamps included
    torch.cuda.synchronize()
    mark = torch.zeros(2, dtype=torch.int64, device=dev)
    lib = _lib.get()
    s = torch.cuda.current_stream().cuda_stream
    # steady state: the host runs ahead of the device, so the last of several back-to-back steps shows the
    # device-side schedule without the host's enqueue latency in it
    for _ in range(4):
        step()
    lib.unetk_debug_stamp(mark.data_ptr(), s)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()                                   # replay
    e1.record()
    lib.unetk_debug_stamp(mark.data_ptr() + 8, s)
    torch.cuda.synchronize()
    t0, t1 = mark.tolist()
    rows = [r for r in tl.intervals() if r[3] >= t0 and r[4] <= t1 and r[4] >= r[3]]
    _lib.timeline = None

    streams = {}
    for r in rows:
        streams.setdefault(r[1], len(streams))
    lines = [f"# one training step, AnomalyUNet B={args.batch} @{args.size}^2, {'eager' if args.eager else 'graph replay'}: "
             f"{(t1 - t0) / 1e3:.1f} us between the step's first and last stamp (CUDA events: {e0.elapsed_time(e1) * 1e3:.1f} us); "
             f"{len(rows)} stamped calls on {len(streams)} streams"]
    tens = [(r[3], r[4]) for r in rows if r[0] in TENSOR]
    busy_t = union_length(tens)
    lines.append(f"# tensor-core kernels: sum of durations {sum(e - s for s, e in tens) / 1e3:.1f} us, union (>= 1 running or "
                 f"waiting for SMs) {busy_t / 1e3:.1f} us, no tensor kernel in flight for {(t1 - t0 - busy_t) / 1e3:.1f} us")
    by_name = {}
    for r in rows:
        d = by_name.setdefault(r[0], [0, 0])
        d[0] += r[4] - r[3]
        d[1] += 1
    lines.append("# per entry point: total us, calls")
    for k, (tot, cnt) in sorted(by_name.items(), key=lambda kv: -kv[1][0]):
        lines.append(f"#   {tot / 1e3:9.1f} {cnt:4d}  {k}")
    # gaps of the tensor pipe: intervals where no tensor kernel is in flight
    iv = sorted(tens)
    gaps, cur = [], t0
    for s_, e_ in iv:
        if s_ > cur:
            gaps.append((cur, s_))
        cur = max(cur, e_)
    if cur < t1:
        gaps.append((cur, t1))
    lines.append("# tensor-pipe gaps longer than 15 us (start us, length us, what ran meanwhile)")
    for gs, ge in gaps:
        if ge - gs < 15000:
            continue
        during = [f"{r[0].replace('unetk_', '')}[{r[2]}]" for r in rows if r[3] < ge and r[4] > gs and r[0] not in TENSOR]
        lines.append(f"#   {(gs - t0) / 1e3:9.1f} {(ge - gs) / 1e3:7.1f}  {', '.join(during[:6])}")
    for sh, si in sorted(streams.items(), key=lambda kv: kv[1]):
        lines.append(f"## stream {si} (handle {sh:#x})")
        lines.append("#  start_us   dur_us   gap_us  call [note]")
        prev = None
        for r in sorted((r for r in rows if r[1] == sh), key=lambda r: r[3]):
            gap = (r[3] - prev) / 1e3 if prev is not None else 0.0
            lines.append(f"{(r[3] - t0) / 1e3:10.1f} {(r[4] - r[3]) / 1e3:8.1f} {gap:8.1f}  {r[0].replace('unetk_', '')} [{r[2]}]")
            prev = r[4]
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
