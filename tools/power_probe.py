"""Power, SM clock and throughput of each kernel family run back to back for a few seconds: the B200 boxes of this
pool hold the training step at their power limit (`sw_power_cap`), so the cost of a kernel inside a long step is its
ENERGY (duration x power at the clock the governor settles on), not its stand-alone time at boost clocks.

    python tools/power_probe.py [--seconds 3]  ->  gpurun_out/power_probe.txt
"""
import argparse
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from tiaozhanbei_unet_b200 import ops, train_utils  # noqa: E402
from tiaozhanbei_unet_b200.model import AnomalyUNet  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=3.0)
args = ap.parse_args()
dev = torch.device("cuda:0")
B, H = 16, 256
lines = []


def log(s):
    print(s, flush=True)
    lines.append(s)


class Sampler:
    def __init__(self):
        self.path = tempfile.mktemp(suffix=".csv")

    def __enter__(self):
        self.f = open(self.path, "w")
        self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw.instant,power.draw", "--format=csv,noheader,nounits",
                                      "-i", "0", "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        self.t0 = time.time()
        return self

    def __exit__(self, *a):
        self.proc.terminate()
        self.proc.wait(timeout=5)
        self.f.close()
        rows = []
        for line in open(self.path):
            try:
                rows.append([float(x) for x in line.split(",")])
            except ValueError:
                pass
        os.unlink(self.path)
        rows = rows[len(rows) // 2:]            # the settled second half
        self.mhz = statistics.median(r[0] for r in rows) if rows else float("nan")
        self.w_inst = statistics.mean(r[1] for r in rows) if rows else float("nan")
        self.w_avg = statistics.mean(r[2] for r in rows) if rows else float("nan")


def bf(*shape):
    return (torch.randn(*shape, device=dev) * 0.5).to(torch.bfloat16)


def probe(name, fn, work, unit):
    """work: FLOP or bytes per call."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn(); torch.cuda.synchronize()
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    t_cold = e0.elapsed_time(e1)                # one call at whatever clock an idle GPU boosts to
    n_batch = max(1, int(0.05 / max(t_cold * 1e-3, 1e-6)))
    calls = 0
    with Sampler() as s:
        t_end = time.time() + args.seconds
        e0.record()
        while time.time() < t_end:
            for _ in range(n_batch):
                fn()
            calls += n_batch
            torch.cuda.synchronize()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / calls
    rate = work / ms / 1e9 if unit == "TF/s" else work / ms / 1e6
    joule = s.w_inst * ms * 1e-3
    log(f"{name:34s} burst {t_cold:8.3f} ms | sustained {ms:8.3f} ms {rate:8.0f} {unit:5s} | {s.mhz:6.0f} MHz {s.w_inst:6.0f} W (avg {s.w_avg:5.0f}) | "
        f"{joule * 1e3:8.2f} mJ/call  {joule / work * 1e12:7.3f} pJ per {'FLOP' if unit == 'TF/s' else 'byte'}")


log(f"# kernel families in a loop for {args.seconds:.0f} s each: burst = one call on an idle GPU, sustained = mean over the loop")
ONLY_EW = os.environ.get("ONLY_EW", "0") == "1"        # tools/power_probe_ew.py: the bandwidth-bound kernels only
if not ONLY_EW:
    a8 = torch.randn(8192, 8192, device=dev).to(torch.bfloat16)
    b8 = torch.randn(8192, 8192, device=dev).to(torch.bfloat16)
    probe("cuBLAS bf16 8192^3 (reference point)", lambda: torch.matmul(a8, b8), 2.0 * 8192 ** 3, "TF/s")
    del a8, b8
    src = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev)
    dst = torch.empty_like(src)
    probe("copy 2 GiB -> 2 GiB (reference point)", lambda: dst.copy_(src), 2.0 * src.numel() * 2, "GB/s")
    del src, dst


def conv_case(name, lvl, c1, c2, cout):
    h = H >> lvl
    cin = c1 + c2
    s1 = bf(B, h, h, c1)
    s2 = bf(B, h, h, c2) if c2 else None
    wf, wd = bf(cout * 9 * cin), bf(cin * 9 * cout)
    y = torch.empty(B, h, h, cout, dtype=torch.bfloat16, device=dev)
    stats = torch.zeros(2 * cout, device=dev)
    dx1 = torch.empty(B, h, h, c1, dtype=torch.bfloat16, device=dev)
    dx2 = torch.empty(B, h, h, c2, dtype=torch.bfloat16, device=dev) if c2 else None
    dw = torch.zeros(cout * 9 * cin, device=dev)
    flop = 2.0 * B * h * h * cout * 9 * cin
    probe(f"conv3x3_fwd   {name}", lambda: ops.conv3x3_fwd(s1, s2, wf, y, stats=stats), flop, "TF/s")
    probe(f"conv3x3_dgrad {name}", lambda: ops.conv3x3_dgrad(y, wd, dx1, False, dx2, False), flop, "TF/s")
    probe(f"conv3x3_wgrad {name}", lambda: ops.conv3x3_wgrad(y, s1, s2, dw), flop, "TF/s")


if not ONLY_EW:
    conv_case("up2.0 512->256 @64^2", 2, 256, 256, 256)
    conv_case("down3.3 512->512 @32^2", 3, 512, 0, 512)
    conv_case("up4.3 64->64 @256^2", 0, 64, 0, 64)
    conv_case("up4.0 128->64 @256^2", 0, 64, 64, 64)
    # transposed conv, the widest
    h = H >> 1
    a = bf(B, h, h, 128)
    wt = bf(4 * 64 * 128)
    bias = torch.zeros(64, device=dev)
    up = torch.empty(B, 2 * h, 2 * h, 64, dtype=torch.bfloat16, device=dev)
    da = torch.empty_like(a)
    dwt = torch.zeros(64 * 4 * 128, device=dev)
    fl = 2.0 * B * h * h * 128 * 4 * 64
    probe("convt2x2_fwd   up4 128->64", lambda: ops.convt2x2_fwd(a, wt, bias, up), fl, "TF/s")
    probe("convt2x2_dgrad up4", lambda: ops.convt2x2_dgrad(up, wt, da), fl, "TF/s")
    probe("convt2x2_wgrad up4", lambda: ops.convt2x2_wgrad(a, up, dwt), fl, "TF/s")
    del a, up, da
# bandwidth kernels at full resolution
y = bf(B, H, H, 64)
act = torch.empty_like(y)
dy = torch.empty_like(y)
sc, sh = torch.ones(64, device=dev), torch.zeros(64, device=dev)
sums = torch.zeros(128, device=dev)
dg, db = torch.empty(64, device=dev), torch.empty(64, device=dev)
nb = y.numel() * 2
probe("bnrelu_apply 64ch @256^2", lambda: ops.bnrelu_apply(y, sc, sh, act), 2.0 * nb, "GB/s")
probe("bn_bwd_reduce", lambda: ops.bn_bwd_reduce(act, y, sc, sh, sh, sc, sums), 2.0 * nb, "GB/s")
probe("bn_bwd_apply", lambda: ops.bn_bwd_apply(act, y, sc, sh, sh, sc, sums, dy, dg, db), 3.0 * nb, "GB/s")
del y, act, dy
if ONLY_EW:
    sys.exit(0)
# whole step
torch.manual_seed(42)
m = AnomalyUNet(3).to(dev).train()
crit = train_utils.CombinedLoss()
opt = train_utils.get_optimizer(m)
g = torch.Generator().manual_seed(0)
image = torch.randn(B, 3, H, H, generator=g).to(dev)
mask = (torch.rand(B, 1, H, H, generator=g) > 0.95).float().to(dev)
probe("train step (AnomalyUNet B=16)", lambda: train_utils.train_step(m, crit, opt, image, mask), 7610.95e9, "TF/s")
m.eval()
with torch.no_grad():
    probe("eval forward (AnomalyUNet B=16)", lambda: m(image), 16 * 158.64e9, "TF/s")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "power_probe.txt"), "w").write("\n".join(lines) + "\n")
