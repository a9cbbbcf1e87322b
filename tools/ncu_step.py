"""Two AnomalyUNet training steps at the BASELINE config-2 shape (B=16 @256^2) for ncu captures: the first
step warms up (allocations, cold L2), kernel filters with --launch-skip select the second."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tiaozhanbei_unet_b200 import train_utils  # noqa: E402
from tiaozhanbei_unet_b200.model import AnomalyUNet  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(42)
m = AnomalyUNet(3).to(dev).train()
crit = train_utils.CombinedLoss()
opt = train_utils.get_optimizer(m, "adam", 1e-3, 1e-4)
g = torch.Generator().manual_seed(0)
image = torch.randn(16, 3, 256, 256, generator=g).to(dev)
mask = (torch.rand(16, 1, 256, 256, generator=g) > 0.95).float().to(dev)
for _ in range(int(os.environ.get("STEPS", "2"))):
    train_utils.train_step(m, crit, opt, image, mask)
torch.cuda.synchronize()
print("done", crit.last_values())
