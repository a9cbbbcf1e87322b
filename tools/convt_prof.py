"""A few transposed-conv launches (up4 / up3 / up1 shapes of the B=16 @256^2 step) for an ncu capture."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from tiaozhanbei_unet_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, H = 16, 256
which = os.environ.get("WHICH", "fwd")
for rep in range(2):
    for lvl, cin in ((1, 128), (2, 256), (4, 1024)):
        h = H >> lvl
        cout = cin // 2
        a = (torch.randn(B, h, h, cin, device=dev) * 0.5).to(torch.bfloat16)
        wt = (torch.randn(4 * cout * cin, device=dev) * 0.05).to(torch.bfloat16)
        bias = torch.zeros(cout, device=dev)
        up = torch.empty(B, 2 * h, 2 * h, cout, dtype=torch.bfloat16, device=dev)
        da = torch.empty_like(a)
        dw = torch.zeros(cout * 4 * cin, device=dev)
        if which == "fwd":
            ops.convt2x2_fwd(a, wt, bias, up)
        elif which == "dgrad":
            ops.convt2x2_dgrad(up, wt, da)
        else:
            ops.convt2x2_wgrad(a, up, dw)
torch.cuda.synchronize()
print("ok")
