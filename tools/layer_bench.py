"""Per-layer timing of the conv kernels at the BASELINE config-2 shapes (AnomalyUNet, B=16, 256x256)
plus the whole training step.  CUDA events on the launching stream, after warm-up.  Diagnostic tool
(writes gpurun_out/layer_bench.txt); the judged numbers come from bench.py."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tiaozhanbei_unet_b200 import ops, packing, _lib  # noqa: E402

dev = torch.device("cuda:0")
B = int(os.environ.get("B", 16))
H = int(os.environ.get("H", 256))
lines = []


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    lines.append(s)


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def bf(*shape):
    return (torch.randn(*shape, device=dev) * 0.5).to(torch.bfloat16)


# (name, level, c1, c2, cout, count)
LAYERS = [("inc.3", 0, 64, 0, 64, 1), ("down1.0", 1, 64, 0, 128, 1), ("down1.3", 1, 128, 0, 128, 1),
          ("down2.0", 2, 128, 0, 256, 1), ("down2.3", 2, 256, 0, 256, 1), ("down3.0", 3, 256, 0, 512, 1),
          ("down3.3", 3, 512, 0, 512, 1), ("down4.0", 4, 512, 0, 1024, 1), ("down4.3", 4, 1024, 0, 1024, 1),
          ("up1.0", 3, 512, 512, 512, 2), ("up1.3", 3, 512, 0, 512, 2), ("up2.0", 2, 256, 256, 256, 2),
          ("up2.3", 2, 256, 0, 256, 2), ("up3.0", 1, 128, 128, 128, 2), ("up3.3", 1, 128, 0, 128, 2),
          ("up4.0", 0, 64, 64, 64, 2), ("up4.3", 0, 64, 0, 64, 2)]
ONLY_STEP = os.environ.get("ONLY_STEP", "0") == "1"
if ONLY_STEP:
    LAYERS = []
tot = {"fwd": 1e-9, "dgrad": 1e-9, "wgrad": 1e-9}
totf = 0.0
log(f"# conv3x3 layers at B={B}, {H}x{H}: ms and TFLOP/s per pass")
for name, lvl, c1, c2, cout, cnt in LAYERS:
    h = H >> lvl
    cin = c1 + c2
    s1 = bf(B, h, h, c1)
    s2 = bf(B, h, h, c2) if c2 else None
    wf = bf(cout * 9 * cin)
    wd = bf(cin * 9 * cout)
    y = torch.empty(B, h, h, cout, dtype=torch.bfloat16, device=dev)
    stats = torch.zeros(2 * cout, device=dev)
    dx1 = torch.empty(B, h, h, c1, dtype=torch.bfloat16, device=dev)
    dx2 = torch.empty(B, h, h, c2, dtype=torch.bfloat16, device=dev) if c2 else None
    dw = torch.zeros(cout * 9 * cin, device=dev)
    flop = 2.0 * B * h * h * cout * 9 * cin
    t_f = timeit(lambda: ops.conv3x3_fwd(s1, s2, wf, y, stats=stats))
    t_d = timeit(lambda: ops.conv3x3_dgrad(y, wd, dx1, False, dx2, False))
    t_w = timeit(lambda: ops.conv3x3_wgrad(y, s1, s2, dw))
    extra = ""
    if not c2:      # dgrad with the fused BatchNorm-backward reduction (single destination)
        ones = torch.ones(c1, device=dev); zer = torch.zeros(c1, device=dev); sm = torch.zeros(2 * c1, device=dev)
        ctx = ops.bn_ctx(s1, ones, zer, zer, ones, sm)
        t_b = timeit(lambda: ops.conv3x3_dgrad(y, wd, dx1, False, bn=ctx))
        extra = f" | dgrad+bnred {t_b:7.3f} ms"
    log(f"{name:8s} M={B*h*h:8d} N={cout:5d} K={9*cin:5d} x{cnt}  fwd {t_f:7.3f} ms {flop/t_f/1e9:7.1f} TF | "
        f"dgrad {t_d:7.3f} ms {flop/t_d/1e9:7.1f} TF | wgrad {t_w:7.3f} ms {flop/t_w/1e9:7.1f} TF" + extra)
    tot["fwd"] += t_f * cnt; tot["dgrad"] += t_d * cnt; tot["wgrad"] += t_w * cnt
    totf += flop * cnt
    del s1, s2, wf, wd, y, dx1, dx2, dw
log(f"sum conv3x3 (tcgen05 layers): fwd {tot['fwd']:.2f} ms, dgrad {tot['dgrad']:.2f} ms, wgrad {tot['wgrad']:.2f} ms; "
    f"{totf/1e12:.2f} TFLOP per pass -> fwd {totf/tot['fwd']/1e9:.0f} TF/s dgrad {totf/tot['dgrad']/1e9:.0f} wgrad {totf/tot['wgrad']/1e9:.0f}")

log("# convT layers")
for k, (lvl, cin) in enumerate([] if ONLY_STEP else [(4, 1024), (3, 512), (2, 256), (1, 128)]):
    h = H >> lvl
    cout = cin // 2
    a = bf(B, h, h, cin)
    wt = bf(4 * cout * cin)
    bias = torch.zeros(cout, device=dev)
    up = torch.empty(B, 2 * h, 2 * h, cout, dtype=torch.bfloat16, device=dev)
    da = torch.empty_like(a)
    dw = torch.zeros(cout * 4 * cin, device=dev)
    flop = 2.0 * B * h * h * cin * 4 * cout
    t_f = timeit(lambda: ops.convt2x2_fwd(a, wt, bias, up))
    t_d = timeit(lambda: ops.convt2x2_dgrad(up, wt, da))
    t_w = timeit(lambda: ops.convt2x2_wgrad(a, up, dw))
    ones = torch.ones(cin, device=dev); zer = torch.zeros(cin, device=dev); sm = torch.zeros(2 * cin, device=dev)
    ctx = ops.bn_ctx(a, ones, zer, zer, ones, sm)
    t_b = timeit(lambda: ops.convt2x2_dgrad(up, wt, da, bn=ctx))
    log(f"up{k+1}.up  cin={cin:5d}  fwd {t_f:7.3f} ms {flop/t_f/1e9:7.1f} TF | dgrad {t_d:7.3f} ms {flop/t_d/1e9:7.1f} TF | wgrad {t_w:7.3f} ms {flop/t_w/1e9:7.1f} TF | dgrad+bnred {t_b:7.3f} ms")

log("# bandwidth kernels at full resolution (64 ch)")
y = bf(B, H, H, 64)
a = torch.empty_like(y)
sc = torch.ones(64, device=dev); sh = torch.zeros(64, device=dev)
nbytes = y.numel() * 2
t = timeit(lambda: ops.bnrelu_apply(y, sc, sh, a))
log(f"bnrelu_apply  {t:.3f} ms  {2*nbytes/t/1e6:.0f} GB/s")
pooled = torch.empty(B, H // 2, H // 2, 64, dtype=torch.bfloat16, device=dev)
idx = torch.empty(B, H // 2, H // 2, 64, dtype=torch.uint8, device=dev)
t = timeit(lambda: ops.bnrelu_pool(y, sc, sh, a, pooled, idx, True))
log(f"bnrelu_pool   {t:.3f} ms  {(2*nbytes + nbytes/4 + nbytes/8)/t/1e6:.0f} GB/s")
sums = torch.zeros(128, device=dev)
t = timeit(lambda: ops.bn_bwd_reduce(a, y, sc, sh, sh, sc, sums))
log(f"bn_bwd_reduce {t:.3f} ms  {2*nbytes/t/1e6:.0f} GB/s")
dy = torch.empty_like(y); dg = torch.empty(64, device=dev); db = torch.empty(64, device=dev)
t = timeit(lambda: ops.bn_bwd_apply(a, y, sc, sh, sh, sc, sums, dy, dg, db))
log(f"bn_bwd_apply  {t:.3f} ms  {3*nbytes/t/1e6:.0f} GB/s")
x = torch.randn(B, 3, H, H, device=dev)
w0 = torch.randn(64, 3, 3, 3, device=dev)
st = torch.zeros(128, device=dev)
t = timeit(lambda: ops.first_conv_fwd(x, w0, y, st))
log(f"first_conv_fwd {t:.3f} ms")
dw0 = torch.empty(64, 3, 3, 3, device=dev)
t = timeit(lambda: ops.first_conv_wgrad(x, y, dw0))
log(f"first_conv_wgrad {t:.3f} ms")
out = torch.empty(B, 3, H, H, device=dev)
hw = torch.randn(3, 64, device=dev); hb = torch.zeros(3, device=dev)
t = timeit(lambda: ops.head_fwd(y, hw, hb, out, True))
log(f"head_fwd {t:.3f} ms  {nbytes/t/1e6:.0f} GB/s")
dwh = torch.empty(3, 64, device=dev); dbh = torch.empty(3, device=dev)
t = timeit(lambda: ops.head_bwd(y, hw, out, out, a, dwh, dbh, True))
log(f"head_bwd {t:.3f} ms  {2*nbytes/t/1e6:.0f} GB/s")
ctx = ops.bn_ctx(y, sc, sh, sh, sc, sums)
t = timeit(lambda: ops.head_bwd(y, hw, out, out, a, dwh, dbh, True, bn=ctx))
log(f"head_bwd+bnred {t:.3f} ms  {2*nbytes/t/1e6:.0f} GB/s")
# the fused output stage of the training step (head + loss term + head backward + BatchNorm-backward sums)
_stats = torch.cat([y.float().sum((0, 1, 2)), (y.float() ** 2).sum((0, 1, 2))]).contiguous()
_g, _b = torch.ones(64, device=dev), torch.zeros(64, device=dev)
_rm, _rv, _nbt = torch.zeros(64, device=dev), torch.ones(64, device=dev), torch.zeros((), dtype=torch.int64, device=dev)
_aux = [torch.empty(64, device=dev) for _ in range(4)]
_fin = ops.bn_finalize_args(_stats, _g, _b, _rm, _rv, _nbt, *_aux, B * H * H, 1e-5, 0.1)
_work, _out3 = ops.head_loss_workspace(dev), torch.zeros(3, device=dev)
_img, _msk = torch.rand(B, 3, H, H, device=dev), (torch.rand(B, 1, H, H, device=dev) > 0.95).float()
_hw1, _hb1 = torch.randn(1, 64, device=dev), torch.zeros(1, device=dev)
_dw1, _db1 = torch.zeros(1, 64, device=dev), torch.zeros(1, device=dev)
_s128 = torch.zeros(128, device=dev)
t = timeit(lambda: ops.head_loss_bwd(y, _fin, hw, hb, _img, None, a, dwh, dbh, _s128, _work, _out3, 0, 1.0, 1.0, 0.25, 2.0))
log(f"head_loss_bwd<3, mse> {t:.3f} ms  {(2*nbytes + _img.numel()*4)/t/1e6:.0f} GB/s")
t2 = timeit(lambda: ops.head_loss_bwd(y, _fin, _hw1, _hb1, _msk, None, a, _dw1, _db1, _s128, _work, _out3, 1, 1.0, 1.0, 0.25, 2.0))
log(f"head_loss_bwd<1, focal> {t2:.3f} ms  {(2*nbytes + _msk.numel()*4)/t2/1e6:.0f} GB/s")
dp_ = bf(B, H // 2, H // 2, 64)
idx8 = torch.randint(0, 4, (B, H // 2, H // 2, 64), dtype=torch.uint8, device=dev)
t = timeit(lambda: ops.maxpool_bwd(dp_, idx8, a, accumulate=True))
log(f"maxpool_bwd(acc) {t:.3f} ms")
t = timeit(lambda: ops.maxpool_bwd(dp_, idx8, a, accumulate=True, bn=ctx))
log(f"maxpool_bwd(acc)+bnred {t:.3f} ms")
del y, a, dy, pooled, idx, x, out

log("# whole training step (AnomalyUNet + CombinedLoss + fused Adam)")
from tiaozhanbei_unet_b200.model import AnomalyUNet  # noqa: E402
from tiaozhanbei_unet_b200 import train_utils  # noqa: E402
torch.manual_seed(42)
m = AnomalyUNet(3).to(dev).train()
crit = train_utils.CombinedLoss()
opt = train_utils.get_optimizer(m)
g = torch.Generator().manual_seed(0)
image = torch.randn(B, 3, H, H, generator=g).to(dev)
mask = (torch.rand(B, 1, H, H, generator=g) > 0.95).float().to(dev)


def step():
    train_utils.train_step(m, crit, opt, image, mask)


n0 = _lib.launch_count()
t = timeit(step, iters=8, warm=3)
n1 = _lib.launch_count()
log(f"train step {t:.3f} ms -> {B/t*1e3:.1f} img/s ; {(n1-n0)//11} C-ABI calls per step; loss {crit.last_values()}")
t0 = time.perf_counter()
for _ in range(5):
    step()
cpu_enqueue = (time.perf_counter() - t0) / 5 * 1e3
torch.cuda.synchronize()
log(f"python enqueue time per step (no sync): {cpu_enqueue:.3f} ms")
m.eval()
with torch.no_grad():
    t = timeit(lambda: m(image), iters=8, warm=3)
log(f"eval forward {t:.3f} ms -> {B/t*1e3:.1f} img/s")
log("# SegmentationUNet(3, 4) + CombinedSegmentationLoss(ce + dice) + fused Adam, B=8 @512x512 (Gear pipeline shape)")
from tiaozhanbei_unet_b200.model import SegmentationUNet  # noqa: E402
from tiaozhanbei_unet_b200 import metrics  # noqa: E402
del m, opt
torch.cuda.empty_cache()
torch.manual_seed(42)
ms_ = SegmentationUNet(3, 4).to(dev).train()
crit_s = metrics.CombinedSegmentationLoss(1.0, 1.0)
opt_s = train_utils.get_optimizer(ms_)
xs = torch.randn(8, 3, 512, 512, generator=g).to(dev)
ts = torch.randint(0, 4, (8, 512, 512), generator=g).to(dev)


def seg_step():
    loss = crit_s(ms_(xs), ts)
    opt_s.zero_grad()
    loss.backward()
    opt_s.step()


t = timeit(seg_step, iters=8, warm=3)
log(f"segmentation train step {t:.3f} ms -> {8/t*1e3:.1f} img/s @512x512 ({8*288.78*4/t:.0f} TFLOP/s); loss terms {crit_s.last_terms()}")
zl = torch.randn(8, 4, 512, 512, device=dev, requires_grad=True)
t_f = timeit(lambda: crit_s(zl, ts))
lv = crit_s(zl, ts)
t_b = timeit(lambda: lv.backward(retain_graph=True))
nb = zl.numel() * 4 + ts.numel() * 8
log(f"seg loss fwd {t_f:.3f} ms ({nb/t_f/1e6:.0f} GB/s of logits+labels), bwd {t_b:.3f} ms ({(nb + zl.numel()*4)/t_b/1e6:.0f} GB/s)")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "layer_bench.txt"), "w").write("\n".join(lines) + "\n")
