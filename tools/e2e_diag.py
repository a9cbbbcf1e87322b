"""Per-epoch timing of train_utils.train_epoch over pinned host batches (prefetch / deferred loss readback)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tiaozhanbei_unet_b200 import train_utils  # noqa: E402
from tiaozhanbei_unet_b200.model import AnomalyUNet  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(42)
model = AnomalyUNet(3).to(dev).train()
crit = train_utils.CombinedLoss()
opt = train_utils.get_optimizer(model, "adam", 1e-3, 1e-4)
g = torch.Generator().manual_seed(0)
host = [{"image": torch.randn(16, 3, 256, 256, generator=g).pin_memory(),
         "mask": (torch.rand(16, 1, 256, 256, generator=g) > 0.95).float().pin_memory()} for _ in range(4)]
train_utils.tqdm = lambda it, **k: it          # no progress bar


class Loader:
    def __init__(self, k):
        self.k = k

    def __len__(self):
        return self.k

    def __iter__(self):
        for i in range(self.k):
            t = time.perf_counter()
            stamps.append(t)
            yield host[i % 4]


resident = [(b["image"].to(dev), b["mask"].to(dev)) for b in host]
for i in range(25):            # bench.py order: resident steps first (graphs get captured here)
    r, a = model(resident[i % 4][0])
    ld = crit(r, a, *resident[i % 4])
    opt.zero_grad()
    ld["total_loss"].backward()
    opt.step()
torch.cuda.synchronize()
stamps = []
train_utils.train_epoch(model, Loader(2), crit, opt, dev, 0)
for epoch in range(3):
    stamps = []
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    train_utils.train_epoch(model, Loader(20), crit, opt, dev, epoch)
    e1.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    gaps = [round((b - a) * 1e3, 1) for a, b in zip(stamps, stamps[1:])]
    print(f"epoch {epoch}: gpu {e0.elapsed_time(e1) / 20:.2f} ms/step, wall {(t1 - t0) * 1e3 / 20:.2f} ms/step; loader gaps ms {gaps}")
