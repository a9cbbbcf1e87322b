"""Diagnostic: error tables of the CUDA path against (a) the fp32 oracle (= the reference's numbers)
and (b) the oracle under the build's bf16 storage contract.  Writes gpurun_out/parity_report.txt."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import unet_oracle as O  # noqa: E402
from tiaozhanbei_unet_b200 import train_utils  # noqa: E402
from tiaozhanbei_unet_b200.model import AnomalyUNet, UNet  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
lines = []


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    lines.append(s)


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()


def stats(got, ref):
    d = (got - ref).abs().flatten()
    return f"max {d.max().item():.2e} p99.9 {torch.quantile(d[:2000000], 0.999).item():.2e} mean {d.mean().item():.2e} (ref max {ref.abs().max().item():.2e})"


def make(arch, sd):
    m = (AnomalyUNet(3) if arch == "anomaly_unet" else UNet(3, 1))
    m.load_state_dict({k: v.clone() for k, v in sd.items()})
    return m.to(dev)


def run(arch, n, h, w, sd, tag):
    log(f"==== {arch} {n}x{h}x{w} weights={tag}")
    g = torch.Generator().manual_seed(33)
    image = torch.randn(n, 3, h, w, generator=g).to(dev)
    mask = (torch.rand(n, 1, h, w, generator=g) > 0.95).float().to(dev)
    fwd = O.anomaly_unet_forward if arch == "anomaly_unet" else O.unet_forward
    # eval
    m = make(arch, sd).eval()
    with torch.no_grad():
        out = m(image)
        r32 = fwd({k: v.clone() for k, v in sd.items()}, image, training=False)
        r16 = fwd({k: v.clone() for k, v in sd.items()}, image, training=False, storage="bf16")
    out = out if isinstance(out, tuple) else (out,)
    r32 = r32 if isinstance(r32, tuple) else (r32,)
    r16 = r16 if isinstance(r16, tuple) else (r16,)
    for i in range(len(out)):
        log(f"eval out{i} vs fp32: {stats(out[i], r32[i])}")
        log(f"eval out{i} vs bf16-contract: {stats(out[i], r16[i])}")
        if arch == "anomaly_unet":
            for mg in (1e-2, 2e-2, 3e-2, 5e-2):
                safe = (r32[i] - 0.5).abs() > mg
                bad = ((out[i] > 0.5) != (r32[i] > 0.5))[safe].sum().item()
                log(f"   mask flips outside margin {mg}: {bad} of {safe.sum().item()}")
    # train
    m = make(arch, sd).train()
    crit = train_utils.CombinedLoss()
    if arch == "anomaly_unet":
        r, a = m(image)
        total = crit(r, a, image, mask)["total_loss"]
        outs = (r, a)
    else:
        lg = m(image)
        p = torch.sigmoid(lg)
        total = F.mse_loss(p, mask) + crit.focal_loss(p, mask)
        outs = (lg,)
    total.backward()
    torch.cuda.synchronize()
    l32, g32 = O.train_step({k: v.clone() for k, v in sd.items()}, image, mask, arch=arch)
    l16, g16 = O.train_step({k: v.clone() for k, v in sd.items()}, image, mask, arch=arch, storage="bf16")
    log(f"loss ours {float(total):.6f} fp32 {l32['total_loss']:.6f} bf16-contract {l16['total_loss']:.6f}")
    w32 = w16 = 0.0
    for k, p in m.named_parameters():
        e32, e16 = rel(p.grad, g32[k]), rel(p.grad, g16[k])
        w32, w16 = max(w32, e32), max(w16, e16)
        log(f"grad {k:45s} vs fp32 {e32:.3e} vs bf16-contract {e16:.3e} (contract vs fp32 {rel(g16[k], g32[k]):.3e}) |ref| {g32[k].norm().item():.2e}")
    log(f"worst grad rel: vs fp32 {w32:.3e}, vs bf16-contract {w16:.3e}")


sd = O.synthetic_state_dict("anomaly_unet", seed=21, device=dev)
run("anomaly_unet", 2, 32, 32, sd, "synthetic21")
run("anomaly_unet", 4, 128, 128, sd, "synthetic21")
sd = O.synthetic_state_dict("unet", seed=21, device=dev)
run("unet", 1, 48, 64, sd, "synthetic21")
# the reference's own initialisation (seed 42, as src/train.py:106,177 would build it)
torch.manual_seed(42)
m0 = AnomalyUNet(3)
sd = {k: v.clone().to(dev) for k, v in m0.state_dict().items()}
run("anomaly_unet", 4, 128, 128, sd, "reference-init-seed42")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "parity_report.txt"), "w").write("\n".join(lines) + "\n")
