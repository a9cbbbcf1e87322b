"""A handful of representative tensor-core conv launches for one ncu capture (run under ncu)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tiaozhanbei_unet_b200 import ops
dev = torch.device("cuda:0")
B = 16
def bf(*s): return (torch.randn(*s, device=dev) * 0.5).to(torch.bfloat16)
def run(h, c1, c2, cout, reps=2):
    cin = c1 + c2
    s1 = bf(B, h, h, c1); s2 = bf(B, h, h, c2) if c2 else None
    wf = bf(cout * 9 * cin); wd = bf(cin * 9 * cout)
    y = torch.empty(B, h, h, cout, dtype=torch.bfloat16, device=dev)
    stats = torch.zeros(2 * cout, device=dev)
    dx1 = torch.empty(B, h, h, c1, dtype=torch.bfloat16, device=dev)
    dx2 = torch.empty(B, h, h, c2, dtype=torch.bfloat16, device=dev) if c2 else None
    dw = torch.zeros(cout * 9 * cin, device=dev)
    for _ in range(reps):
        ops.conv3x3_fwd(s1, s2, wf, y, stats=stats)
        ops.conv3x3_dgrad(y, wd, dx1, False, dx2, False)
        ops.conv3x3_wgrad(y, s1, s2, dw)
    torch.cuda.synchronize()
import sys as _s
which = _s.argv[1] if len(_s.argv) > 1 else "all"
if which in ("all", "n64"):
    run(256, 64, 0, 64)      # inc.3 / up4.3: N=64, K=576
    run(128, 64, 0, 128)     # down1.0: N=128, K=576
if which in ("all", "mid"):
    run(128, 128, 0, 128)    # down1.3 / up3.3
    run(32, 512, 512, 512)   # up1.0: long K
print("done")
