"""Run under torchrun (N ranks, one GPU each): data-parallel gradient parity.
Every rank trains on its own shard; after backward all ranks must hold bit-identical gradients, and
those must equal the mean of the per-shard gradients (checked on rank 0 against the fp32 oracle run
on every shard: SURVEY.md §7.2 item 6, "reference per shard + gradient mean")."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import unet_oracle as O  # noqa: E402
from tiaozhanbei_unet_b200 import distributed as D, train_utils  # noqa: E402
from tiaozhanbei_unet_b200.model import AnomalyUNet  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
import datetime
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))

sd = O.synthetic_state_dict("anomaly_unet", seed=3)
m = AnomalyUNet(3)
m.load_state_dict(sd)
m = m.to(dev).train()
D.broadcast_parameters(m)
SYNC_BN = "--sync-bn" in sys.argv       # exact global-batch semantics instead of per-shard BatchNorm
red = D.attach(m, sync_bn=SYNC_BN)
crit = train_utils.CombinedLoss()
opt = train_utils.get_optimizer(m)
n, h, w = 4, 64, 64


def shard(r):
    # shards with different statistics (scale / offset per rank), so that per-shard and global-batch
    # BatchNorm are clearly different models
    g = torch.Generator().manual_seed(500 + r)
    return torch.randn(n, 3, h, w, generator=g) * (1.0 + 0.5 * r) + 0.3 * r, (torch.rand(n, 1, h, w, generator=g) > 0.95).float()


image, mask = (t.to(dev) for t in shard(rank))
FUSED = "--fused" in sys.argv           # the one-sequence training step (Engine.train_step) instead of model / criterion / backward
UPDATE = "--update" in sys.argv         # ... with the optimizer step inside it (unetk_fused_update after the all-reduces)
if UPDATE and not SYNC_BN:
    # (a tiny learning rate: the third step's gradients are still comparable with the oracle at the initial weights)
    opt = train_utils.get_optimizer(m, "adam", 1e-9, 0.0)
    assert train_utils._fusable(m, crit, image, mask) and train_utils.FUSED_UPDATE
    for _ in range(3):
        train_utils.train_step(m, crit, opt, image, mask)
    eng = m._engine()
    assert eng.gstage_zero and not eng.weights_dirty and float(opt.state_dict()["state"][0]["step"]) == 3.0
elif FUSED and not SYNC_BN:
    assert train_utils._fusable(m, crit, image, mask)
    for _ in range(3):                  # eager call, graph capture (forward half; the all-reduces stay eager), replay
        opt.zero_grad()
        m._engine().train_step(image, mask, (crit.recon_weight, crit.seg_weight, crit.focal_alpha, crit.focal_gamma))
else:
    r_, a_ = m(image)
    ld = crit(r_, a_, image, mask)
    opt.zero_grad()
    ld["total_loss"].backward()
torch.cuda.synchronize()
flat = m._engine().flat_g
# 1) identical on all ranks
chk = torch.stack([flat.double().sum(), flat.double().abs().sum()])
allc = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
same = all(torch.equal(allc[0], c) for c in allc)
ok = same
msg = f"identical across ranks: {same}; reduced {red.bytes_reduced / 1e6:.1f} MB in buckets"
if UPDATE and not SYNC_BN:
    # the updated weights are identical on every rank, their packed bf16 copies are current, the staging buffer is zero
    eng = m._engine()
    pc = torch.stack([eng.flat_p.double().sum(), eng.flat_p.double().abs().sum()])
    allp = [torch.zeros_like(pc) for _ in range(world)]
    dist.all_gather(allp, pc)
    same_p = all(torch.equal(allp[0], c) for c in allp)
    have = eng.wpack.clone()
    eng.weights_dirty = True
    eng._pack_weights(split=False)
    torch.cuda.synchronize()
    packed = torch.equal(have.view(torch.int16), eng.wpack.view(torch.int16))
    clean = not eng.gstage.any().item()
    moved = (eng.flat_p - torch.cat([torch.nn.functional.pad(sd[k].to(dev).flatten(), (0, (-sd[k].numel()) % 64)) for k in eng.param_names])).abs().max().item()
    ok = ok and same_p and packed and clean and 0.0 < moved < 1e-7
    msg += f"; after 3 fused updates: weights identical across ranks {same_p}, packed copies current {packed}, staging zero {clean}, max |dp| {moved:.1e}"
if rank == 0:
    if SYNC_BN:      # the reference on the concatenated global batch (mean loss over all images)
        im = torch.cat([shard(r)[0] for r in range(world)]).to(dev)
        mk = torch.cat([shard(r)[1] for r in range(world)]).to(dev)
        _, ref = O.train_step({k: v.to(dev).clone() for k, v in sd.items()}, im, mk, arch="anomaly_unet")
    else:
        ref = None
        for r in range(world):
            im, mk = (t.to(dev) for t in shard(r))
            _, g = O.train_step({k: v.to(dev).clone() for k, v in sd.items()}, im, mk, arch="anomaly_unet")
            ref = g if ref is None else {k: ref[k] + g[k] for k in g}
        ref = {k: v / world for k, v in ref.items()}
    worst_cos, head = 1.0, 0.0
    for k, p in m.named_parameters():
        g, gr = p.grad, ref[k]
        cos = ((g * gr).sum() / (g.norm() * gr.norm() + 1e-30)).item()
        worst_cos = min(worst_cos, cos)
        if k.startswith("outc"):
            head = max(head, ((g - gr).norm() / gr.norm()).item())
        nr = abs(g.norm().item() - gr.norm().item()) / gr.norm().item()
        ok = ok and cos > 0.6 and nr < 0.2
    ok = ok and head < 3e-2
    msg += (f"; vs {'the oracle on the global batch (SyncBN)' if SYNC_BN else 'mean of per-shard oracle grads'}: "
            f"worst cosine {worst_cos:.3f}, head rel err {head:.2e}")
    if SYNC_BN:      # control: the per-shard-BatchNorm model must NOT match (the check discriminates the two semantics)
        alt = None
        for r in range(world):
            ims, mks = (t.to(dev) for t in shard(r))
            _, g = O.train_step({k: v.to(dev).clone() for k, v in sd.items()}, ims, mks, arch="anomaly_unet")
            alt = g if alt is None else {k: alt[k] + g[k] for k in g}
        ctrl = max(((p.grad - alt[k] / world).norm() / (alt[k].norm() / world)).item() for k, p in m.named_parameters() if k.startswith("outc"))
        ok = ok and ctrl > 2 * max(head, 1e-2)
        msg += f" (control vs per-shard BatchNorm: {ctrl:.2e})"
    if SYNC_BN:      # running statistics are those of the global batch
        msd, rsd = m.state_dict(), {k: v.to(dev).clone() for k, v in sd.items()}
        O.anomaly_unet_forward(rsd, im, training=True)
        dev_rs = max(((msd[k] - rsd[k]).abs().max() / (rsd[k].abs().max() + 1e-3)).item() for k in rsd
                     if k.endswith(("running_mean", "running_var")))
        ok = ok and dev_rs < 5e-2
        msg += f", running-stat rel dev {dev_rs:.2e}"
if not UPDATE:
    opt.step()
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(("DDP CHECK PASSED: " if flag.item() else "DDP CHECK FAILED: ") + msg, flush=True)
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
