"""Headline benchmark: AnomalyUNet training throughput (images/s) at 256x256 on N B200s.

    python bench.py --gpus N --steps K --warmup W            # this build (CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

Workload = BASELINE.json configs[1]: AnomalyUNet dual-decoder training, bf16 storage / fp32
accumulate, 256x256, batch 16 per GPU, CombinedLoss (MSE + focal), Adam lr 1e-3 wd 1e-4, synthetic
seeded RGB batches, random-init weights of the reference architecture (seed 42).  For N > 1 the
script is launched with torch.distributed.run, one rank per GPU; the batch is sharded by rank
(weak scaling: 16 images per GPU) and gradients are averaged with bucketed NCCL all-reduce.

One JSON line on stdout (rank 0).  `value` = images/s with the batch already resident in HBM;
`e2e` = the same steps through the public API — train_utils.train_epoch, the call src/train.py makes — over a
loader of pinned HOST batches: H2D copy of image+mask and D2H read of the three loss scalars of every step inside
the timed region.  `roofline` = the dominant kernel
family (tcgen05 conv3x3 implicit GEMM: fwd + dgrad + wgrad launches), algorithmic FLOPs / CUDA-event
time of exactly those launches, measured in an instrumented pass of the same step.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "train_images_per_sec_256x256"
UNIT = "img/s"
PER_GPU_BATCH = 16
H = W = 256
TRAIN_GFLOP_PER_IMG = 475.68      # BASELINE.md §3 (fwd + dgrad + wgrad, AnomalyUNet @256^2)


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"tflops_sustained": p.get("bf16_tflops_sustained", 1400.1), "tflops_burst": p.get("bf16_tflops", 1683.7),
                "hbm_gbs": p.get("hbm_gbs", 6543.1), "source": "MEASURED_PEAKS.json (of measured)"}
    return {"tflops_sustained": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0,
            "source": "B200_PROFILING.md fallback (of fallback)"}


def ncu_traffic(kernel, key="dram_bytes_per_launch"):
    """DRAM bytes per launch (or another recorded figure) of `kernel` from the committed `ncu --set full`
    summary of this round (profiles/ncu_traffic.json, written by tools/ncu_summarise.py), or None."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        return t.get(kernel.split(" ")[0], {}).get(key)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, smax, reasons, power = [], [], set(), []
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); smax.append(float(c[2])); power.append(float(c[3]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            load = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
            out = {"sm_mhz": statistics.median(load), "sm_max_mhz": max(smax), "reasons": sorted(reasons),
                   "samples": len(sm), "power_w_max": max(power)}
        return out


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference algorithm (oracle port, fp32) on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_reference_rate(steps, warmup, budget_s):
    """images/s of the reference training step (AnomalyUNet + CombinedLoss + Adam, fp32, all host
    threads) on a bounded sample: the per-step batch is shrunk so the whole run fits the budget."""
    import torch
    from oracle import unet_oracle as O
    try:                      # torchrun exports OMP_NUM_THREADS=1: take every core we are allowed to use
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sd = O.synthetic_state_dict("anomaly_unet", seed=42)
    g = torch.Generator().manual_seed(0)

    def run(b, step):
        image = torch.randn(b, 3, H, W, generator=g)
        mask = (torch.rand(b, 1, H, W, generator=g) > 0.95).float()
        t0 = time.perf_counter()
        O.train_step(sd, image, mask, arch="anomaly_unet", opt_state=state, step=step)
        return time.perf_counter() - t0

    state = {}
    t1 = run(1, 1)                                    # probe (also first-touch warm-up)
    per_img = max(t1, 1e-3)
    b = int(max(1, min(PER_GPU_BATCH, budget_s / (per_img * (steps + warmup)))))
    for i in range(warmup):
        run(b, 2 + i)
    times = [run(b, 2 + warmup + i) for i in range(steps)]
    rate = b / statistics.median(times)
    cpu_reference_rate.last_step_ms = 1e3 * statistics.median(times)
    return rate, threads, f"AnomalyUNet train step (fwd+loss+bwd+Adam, fp32) on {b} of the {PER_GPU_BATCH} images of a 256x256 batch, " \
                          f"{steps} timed steps after {warmup} warm-up, median"


def workload_config(world):
    """`config` of the JSON line: the same dict on both arms (the reference arm times the same workload)."""
    return {"workload": "AnomalyUNet train 256x256, CombinedLoss(MSE+focal), Adam lr1e-3 wd1e-4 — BASELINE configs[1]",
            "per_gpu_batch": PER_GPU_BATCH, "global_batch": PER_GPU_BATCH * world,
            "parallelism": f"dp{world} (batch-sharded, bucketed NCCL all-reduce overlapped with backward, local BN)",
            "l2": "no explicit flush: one step streams ~8 GB of activations, far beyond the 126 MB L2",
            "streams": "decoder chains on 2 streams, weight gradients on side streams (engine.py)",
            "weights": "reference initialiser, seed 42"}


def run_reference(args):
    """The reference's algorithm for this path on the box's host cores (all threads), K timed steps after W warm-up
    steps exactly as asked; every step is a bounded sample of the workload's batch (see `cpu_baseline.sample`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    rate, threads, sample = cpu_reference_rate(args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": cpu_reference_rate.last_step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "device": "host CPU (reference algorithm restated in oracle/unet_oracle.py, PyTorch fp32 ops)",
                         "wall_s": round(time.perf_counter() - t0, 1)},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# this build
# ---------------------------------------------------------------------------------------------
def measure_inference(torch, dist, world, rank, dev, timed):
    """UNet eval: sigmoid(model(x)) + compute_anomaly_score on this rank's shard of 64 512x512 images."""
    from tiaozhanbei_unet_b200 import distributed as D, utils
    from tiaozhanbei_unet_b200.model import UNet
    total = 64
    b, e = D.shard_range(total, rank, world)
    torch.manual_seed(42)
    m = UNet(3, 1).to(dev).eval()
    gen = torch.Generator().manual_seed(2000 + rank)
    host = torch.randn(e - b, 3, 512, 512, generator=gen).pin_memory()
    x = host.to(dev)

    def run(img):
        with torch.no_grad():
            amap = torch.sigmoid(m(img))
            score = utils.compute_anomaly_score(img, img)       # UNet branch: reconstruction = images
        return amap, score

    for _ in range(3):
        run(x)
    k = 5
    ms = timed(lambda i: run(x), k)

    # end to end: images from pinned host memory, both maps back into pinned host buffers (what
    # src/test.py:99-106 does with .cpu(), without the pageable staging copy)
    host_amap = torch.empty(e - b, 1, 512, 512, dtype=torch.float32).pin_memory()
    host_score = torch.empty(e - b, 512, 512, dtype=torch.float32).pin_memory()

    def e2e_serial(i):
        amap, score = run(host.to(dev, non_blocking=True))
        host_amap.copy_(amap, non_blocking=True)
        host_score.copy_(score, non_blocking=True)
    e2e_serial(0)
    ms_serial = timed(e2e_serial, k)
    # ... and through the package's entry point for that loop: utils.predict_batches pipelines copy-in, forward and
    # copy-out over batches of the shard (two device slots, two side streams)
    chunk = 16 if (e - b) >= 32 else (e - b)     # small shards: one batch (batches of 2-4 images of 512^2 lose more in the forward pass than the overlap gives: 8 GPUs x 8 images, batches of 4: 0.73 of resident)
    res = utils.predict_batches(m, host, chunk)
    ms_e2e = timed(lambda i: utils.predict_batches(m, host, chunk, out=res), k)
    del m, x
    torch.cuda.empty_cache()
    return {"metric": "inference_images_per_sec_512x512", "value": total * k / (ms / 1e3), "unit": "img/s",
            "e2e": total * k / (ms_e2e / 1e3), "e2e_one_batch_serial": total * k / (ms_serial / 1e3), "ms_per_pass": ms / k,
            "e2e_api": f"utils.predict_batches(model, pinned host images, batch_size={chunk})",
            "config": f"UNet(3,1) eval, sigmoid anomaly map + anomaly score, {total} images of 512x512 sharded {world}-way "
                      f"({e - b} per GPU), no communication; e2e adds the H2D of the images and the D2H of both maps (pinned buffers), pipelined over batches by utils.predict_batches; e2e_one_batch_serial = copy in, one forward over the whole shard, copy out, on one stream",
            "e2e_h2d_bytes_per_pass": host.numel() * 4, "e2e_d2h_bytes_per_pass": (host_amap.numel() + host_score.numel()) * 4,
            "gflop_per_image": 385.34}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from tiaozhanbei_unet_b200 import _lib, distributed as D, ops, train_utils
    from tiaozhanbei_unet_b200.model import AnomalyUNet

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # the JSON line must be the only thing on stdout: libraries (NCCL prints its version banner there)
    # write to the stderr file descriptor for the rest of the run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    torch.manual_seed(42)                       # reference default --seed 42 (src/train.py:78)
    model = AnomalyUNet(3).to(dev).train()
    criterion = train_utils.CombinedLoss()
    optimizer = train_utils.get_optimizer(model, "adam", 1e-3, 1e-4)
    if world > 1:
        D.broadcast_parameters(model)
        D.attach(model)

    B = PER_GPU_BATCH
    gen = torch.Generator().manual_seed(1000 + rank)
    n_host = 4                                  # rotating pinned host batches (fresh data per step)
    host = [(torch.randn(B, 3, H, W, generator=gen).pin_memory(),
             (torch.rand(B, 1, H, W, generator=gen) > 0.95).float().pin_memory()) for _ in range(n_host)]
    resident = [(a.to(dev), b.to(dev)) for a, b in host]

    def step(image, mask):
        # the loop body of src/train_utils.py:122-135 (forward, CombinedLoss, zero_grad, backward, optimizer step)
        # through the package's public entry point for it
        train_utils.train_step(model, criterion, optimizer, image, mask)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, k):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # ---- resident-input throughput
    for i in range(args.warmup):
        step(*resident[i % n_host])
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    n0 = _lib.kernel_launch_count()
    ms = timed(lambda i: step(*resident[i % n_host]), args.steps)
    launches = _lib.kernel_launch_count() - n0
    clocks = sampler.stop() if sampler else None
    value = world * B * args.steps / (ms / 1e3)

    # ---- end to end through the public API with host batches: train_utils.train_epoch (the call src/train.py:232
    # makes) over a loader of pinned host batches in the dataset's dict form; every step copies its batch to the
    # device (train_epoch issues the next batch's copy on a side stream) and reads the three loss scalars back
    class HostLoader:
        def __init__(self, k):
            self.k = k

        def __len__(self):
            return self.k

        def __iter__(self):
            for i in range(self.k):
                hi, hm = host[i % n_host]
                yield {"image": hi, "mask": hm}

    train_utils.train_epoch(model, HostLoader(2), criterion, optimizer, dev, 0)
    ms_e2e = timed(lambda _: train_utils.train_epoch(model, HostLoader(args.steps), criterion, optimizer, dev, 0), 1)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    h2d = (B * 3 * H * W + B * H * W) * 4
    d2h = 12

    # ---- roofline of the dominant kernel family: instrumented pass (events around conv launches)
    # (every rank runs these steps: they contain the gradient all-reduce; rank 0 reports its own GPU)
    roof = None
    rec = ops.enable_conv_timing()
    k_inst = min(args.steps, 5)
    for i in range(k_inst):
        step(*resident[i % n_host])
    torch.cuda.synchronize()
    by_kernel = ops.conv_timing_by_kernel(rec)
    flops, tms, n_launch = ops.collect_conv_timing(rec)
    if rank == 0:
        peaks = read_peaks()
        # dominant kernel = the instantiation with the largest share of the step
        top = max(by_kernel, key=lambda k: by_kernel[k][1])
        tf, tms_top, n_top = by_kernel[top]
        achieved = tf / (tms_top / 1e3) / 1e12 if tms_top > 0 else 0.0
        fam = flops / (tms / 1e3) / 1e12 if tms > 0 else 0.0
        roof = {"bound": "tensor", "kernel": top + " (tcgen05 implicit GEMM: conv3x3 fwd/dgrad and ConvTranspose fwd/dgrad "
                                                   "launches with >= 256 output columns)",
                "achieved": achieved, "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["tflops_sustained"], "traffic": ncu_traffic(top),
                "tensor_pipe_active_pct_ncu": ncu_traffic(top, "tensor_pipe_active_pct_time_weighted"),
                "traffic_unit": "bytes per launch (dram read + write, ncu --set full, mean over one step's launches)",
                "flop_per_launch": tf / n_top, "launches_per_step": n_top // k_inst, "kernel_ms_per_step": tms_top / k_inst,
                "share_of_step": (tms_top / k_inst) / (ms / args.steps), "peak_source": peaks["source"],
                "timing": "CUDA events around every launch in a single-stream instrumented pass of the same step",
                "all_tcgen05_kernels": {"achieved": fam, "frac": fam / peaks["tflops_sustained"],
                                        "launches_per_step": n_launch // k_inst, "kernel_ms_per_step": tms / k_inst,
                                        "by_kernel": {k: {"tflops": v[0] / (v[1] / 1e3) / 1e12, "ms_per_step": v[1] / k_inst,
                                                          "launches_per_step": v[2] // k_inst} for k, v in by_kernel.items()}}}

    # ---- secondary figure: BASELINE configs[3], UNet inference anomaly-map scoring, 64 images of 512x512
    # sharded contiguously over the ranks, no communication (src/test.py:88-96 pattern)
    infer = None
    if not args.no_inference:
        infer = measure_inference(torch, dist, world, rank, dev, timed)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, threads, sample = cpu_reference_rate(steps=2, warmup=1, budget_s=25.0)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": workload_config(world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps,
                    "api": "train_utils.train_epoch(model, loader of pinned host batches, criterion, optimizer, device, epoch)"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "inference": infer,
            "tensor_frac_whole_step": world and (TRAIN_GFLOP_PER_IMG * B * args.steps / (ms / 1e3) / 1e3) / read_peaks()["tflops_sustained"],
            "loss": criterion.last_values()[0],
        }
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-inference", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
