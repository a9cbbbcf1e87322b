"""Stand-alone timing of unetk_fused_update over the parameter groups of an AnomalyUNet (bytes = 36 per conv weight:
read staged gradient, p, m, v; write zero, g, p, m, v, two bf16 copies)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from tiaozhanbei_unet_b200 import ops, train_utils  # noqa: E402
from tiaozhanbei_unet_b200.model import AnomalyUNet  # noqa: E402

dev = torch.device("cuda", 0)
m = AnomalyUNet(3).to(dev).train()
eng = m._engine()
eng.ensure(dev)
opt = train_utils.get_optimizer(m, "adam", 1e-3, 1e-4)
upd = opt.fused_update_args(eng)
ops.set_floats(eng.hyper, upd.values)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=5):
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


for part, name in ((("dec", 0), "decoder 0"), (("enc", 3), "encoder 3-4"), (("enc", 0), "encoder 0-2")):
    tab = eng.update_tables[part]
    n = sum(e[7] for e in tab.entries)
    for mb in (0, 148, 296, 444, 592):
        t = timeit(lambda: ops.fused_update(tab, eng.flat_p, eng.flat_g, upd.m, upd.v, eng.gstage, eng.wpack, eng.hyper, upd.mode, max_blocks=mb))
        print(f"fused_update {name:12s} {n / 1e6:6.2f} M params, {tab.n_blocks:5d} units, max_blocks {mb:4d}: {t * 1e3:7.1f} us  {36 * n / t / 1e6:7.0f} GB/s")
t = timeit(lambda: ops.adam_step(eng.flat_p, eng.flat_g, upd.m, upd.v, 1e-3, 0.9, 0.999, 1e-8, 1e-4, 3))
print(f"adam_step (all {eng.total / 1e6:.1f} M): {t * 1e3:.1f} us  {28 * eng.total / t / 1e6:.0f} GB/s")
# control: the same kernel over the whole flat buffer as ONE "plain" tensor (no tiles, no packed copies: a streaming
# pass of 36 bytes per element) -- what the tiled access pattern costs
from tiaozhanbei_unet_b200 import packing  # noqa: E402
from tiaozhanbei_unet_b200 import _lib  # noqa: E402
ctl = packing.UpdateTable(_lib.get().unetk_update_plain_chunk_elems())
ctl.add_plain(0, eng.total)
ctl.finalize(dev)
t = timeit(lambda: ops.fused_update(ctl, eng.flat_p, eng.flat_g, upd.m, upd.v, eng.gstage, eng.wpack, eng.hyper, upd.mode))
print(f"fused_update, everything as one plain tensor ({eng.total / 1e6:.1f} M): {t * 1e3:.1f} us  {36 * eng.total / t / 1e6:.0f} GB/s")
