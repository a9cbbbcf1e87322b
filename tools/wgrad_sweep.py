"""Split-K sweep of the weight-gradient kernels at the config-2 layer shapes (calibrates the cost
model in csrc/wgrad.cu: choose_ksplit).  UNETK_WGRAD_KSPLIT forces the split factor."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tiaozhanbei_unet_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, H = 16, 256


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def bf(*shape):
    return (torch.randn(*shape, device=dev) * 0.5).to(torch.bfloat16)


LAYERS = [("inc.3", 0, 64, 0, 64), ("down1.0", 1, 64, 0, 128), ("down1.3", 1, 128, 0, 128), ("down2.0", 2, 128, 0, 256),
          ("down2.3", 2, 256, 0, 256), ("down3.0", 3, 256, 0, 512), ("down3.3", 3, 512, 0, 512), ("down4.0", 4, 512, 0, 1024),
          ("down4.3", 4, 1024, 0, 1024), ("up1.0", 3, 512, 512, 512), ("up2.0", 2, 256, 256, 256), ("up3.0", 1, 128, 128, 128),
          ("up4.0", 0, 64, 64, 64)]
KS = [1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 49, 64, 74, 98, 148, 222, 296]
print("# conv3x3 wgrad: ms per split-K factor (default = cost model)")
for name, lvl, c1, c2, cout in LAYERS:
    h = H >> lvl
    cin = c1 + c2
    s1 = bf(B, h, h, c1); s2 = bf(B, h, h, c2) if c2 else None
    dy = bf(B, h, h, cout)
    dw = torch.zeros(cout * 9 * cin, device=dev)
    os.environ.pop("UNETK_WGRAD_KSPLIT", None)
    t0 = timeit(lambda: ops.conv3x3_wgrad(dy, s1, s2, dw))
    res = []
    for ks in KS:
        if ks > B * (h // 8) * (h // 8):
            continue
        os.environ["UNETK_WGRAD_KSPLIT"] = str(ks)
        res.append((ks, timeit(lambda: ops.conv3x3_wgrad(dy, s1, s2, dw))))
    best = min(res, key=lambda r: r[1])
    print(f"{name:8s} cin={cin:5d} cout={cout:5d} pix={B*h*h:8d} default {t0:.3f} | best ks={best[0]} {best[1]:.3f} | " +
          " ".join(f"{k}:{t:.3f}" for k, t in res), flush=True)
    del s1, s2, dy, dw
print("# convT wgrad")
for lvl, cin in [(4, 1024), (3, 512), (2, 256), (1, 128)]:
    h = H >> lvl
    cout = cin // 2
    a = bf(B, h, h, cin); up = bf(B, 2 * h, 2 * h, cout)
    dw = torch.zeros(cout * 4 * cin, device=dev)
    os.environ.pop("UNETK_WGRAD_KSPLIT", None)
    t0 = timeit(lambda: ops.convt2x2_wgrad(a, up, dw))
    res = []
    for ks in KS:
        if ks > B * h * h // 64:
            continue
        os.environ["UNETK_WGRAD_KSPLIT"] = str(ks)
        res.append((ks, timeit(lambda: ops.convt2x2_wgrad(a, up, dw))))
    best = min(res, key=lambda r: r[1])
    print(f"convT cin={cin:5d} default {t0:.3f} | best ks={best[0]} {best[1]:.3f} | " + " ".join(f"{k}:{t:.3f}" for k, t in res), flush=True)
os.environ.pop("UNETK_WGRAD_KSPLIT", None)
