"""Check of the fused BatchNorm-backward reductions (UNETK_FUSE_BN_REDUCE=all): inside one training
step every fused per-channel sum is compared with the stand-alone reduction kernel run on the very same
gradient tensor (UNETK_FUSE_CHECK=1 in engine.py), so run-to-run atomics noise -- which bf16 storage
amplifies chaotically through 26 BatchNorm layers -- does not blur the comparison."""
import os
import sys

os.environ.setdefault("UNETK_FUSE_BN_REDUCE", "all")
os.environ["UNETK_FUSE_CHECK"] = "1"
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import unet_oracle as O  # noqa: E402
from tiaozhanbei_unet_b200 import train_utils  # noqa: E402
from tiaozhanbei_unet_b200.model import AnomalyUNet, UNet  # noqa: E402

dev = torch.device("cuda:0")
bad = 0
for arch, n, h, w in [("unet", 2, 32, 32), ("anomaly_unet", 2, 32, 32), ("anomaly_unet", 3, 64, 48), ("unet", 2, 128, 128),
                      ("anomaly_unet", 1, 52, 76), ("anomaly_unet", 4, 256, 256)]:
    sd = O.synthetic_state_dict(arch, seed=1, device=dev)
    model = (AnomalyUNet(3) if arch == "anomaly_unet" else UNet(3, 1))
    model.load_state_dict(sd)
    model = model.to(dev).train()
    g = torch.Generator().manual_seed(0)
    image = torch.randn(n, 3, h, w, generator=g).to(dev)
    mask = (torch.rand(n, 1, h, w, generator=g) > 0.95).float().to(dev)
    if arch == "anomaly_unet":
        crit = train_utils.CombinedLoss()
        recon, amap = model(image)
        loss = crit(recon, amap, image, mask)["total_loss"]
    else:
        p = torch.sigmoid(model(image))
        loss = torch.nn.functional.mse_loss(p, mask)
    loss.backward()
    torch.cuda.synchronize()
    recs = model._engine().fuse_check
    worst = max(recs, key=lambda r: r[1])
    ok = worst[1] < 2e-2
    bad += not ok
    print(f"[fuse check] {arch} {n}x{h}x{w}: {len(recs)} fused layers, worst rel err {worst[1]:.3e} at {worst[0]} -> {'OK' if ok else 'MISMATCH'}")
    if not ok:
        for name, err in recs:
            print(f"    {name:45s} {err:.3e}")
    model._engine().fuse_check.clear()
sys.exit(1 if bad else 0)
