"""Network-level parity of the CUDA path (model -> engine -> C ABI) against the oracle and the golden
vectors recorded from the reference, on identical seeded inputs and weights.

Tolerances.  The build stores activations and their gradients in bf16 (fp32 accumulation).  The
noise floor of that storage contract was measured on the reference algorithm itself
(oracle storage="bf16" vs storage="fp32", tests/checks/parity_report.py, profiles/parity_r01.txt): on the
He-scaled synthetic weights used here, sigmoid outputs deviate by mean 3e-3 / p99.9 6e-2 / max 1e-1,
and deep-layer gradients by 50-60 % in relative L2 (cosine ~0.85) although their norms agree to a few
percent and loss curves agree to 3e-3 — the deviation between two bf16 evaluations of the SAME
algorithm is as large as either one's deviation from fp32, i.e. it is rounding chaos through 23
BatchNorm/ReLU layers, not a property of these kernels (each kernel is checked tightly on its own in
tests/test_ops_gpu.py).  Hence, against the fp32 reference numbers:
  * outputs: mean |d| <= 8e-3 * max|ref|, max |d| <= 1.5e-1 * max|ref|
  * losses: relative 2e-3
  * gradients: per tensor cosine >= 0.6 and norm within 30 % (the bf16-contract ORACLE is itself 18 % off
    the fp32 norm on inc.double_conv.1.weight of the 2x32x32 UNet case, 7 % on the others, so 20 % had no
    margin: the case flipped between runs with the order of the fp32 atomics); the output-side tensors
    (heads and the last BatchNorm), which see no accumulated noise, within 3e-2 relative L2
  * threshold masks: bit-exact wherever the reference margin |p - 0.5| exceeds 1.5e-1, and fewer than
    0.2 % flips outside a 5e-2 margin
  * loss curve over 100 Adam steps within 1 %
"""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import unet_oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "unet_golden.pt")
OUT_MAX, OUT_MEAN, LOSS_REL, GRAD_COS, GRAD_NORM, GRAD_REL_HEAD, MARGIN, SOFT_MARGIN = 1.5e-1, 8e-3, 2e-3, 0.6, 0.3, 3e-2, 1.5e-1, 5e-2
HEAD_KEYS = ("outc", "up4.conv.double_conv.4", "up4_recon.conv.double_conv.4", "up4_seg.conv.double_conv.4")


def _check_masks(got, ref, what):
    safe = (ref - 0.5).abs() > MARGIN
    assert torch.equal((got > 0.5)[safe], (ref > 0.5)[safe]), f"{what}: mask flip outside margin {MARGIN}"
    soft = (ref - 0.5).abs() > SOFT_MARGIN
    flips = ((got > 0.5) != (ref > 0.5))[soft].float().mean().item() if soft.any() else 0.0
    assert flips <= 2e-3, f"{what}: {flips:.2e} of the masks flip outside margin {SOFT_MARGIN}"


def _check_grads(named_params, ref_grads, what):
    worst_cos, worst_head = 1.0, 0.0
    for k, p in named_params:
        g, gr = p.grad, ref_grads[k]
        assert g is not None and g.shape == gr.shape and torch.isfinite(g).all(), k
        ng, nr = g.norm().item(), gr.norm().item()
        cos = (g * gr).sum().item() / (ng * nr + 1e-30)
        if k.startswith(HEAD_KEYS):
            relerr = (g - gr).norm().item() / (nr + 1e-30)
            worst_head = max(worst_head, relerr)
            assert relerr <= GRAD_REL_HEAD, f"{what} grad {k}: rel err {relerr:.3e}"
        worst_cos = min(worst_cos, cos)
        assert cos >= GRAD_COS, f"{what} grad {k}: cosine {cos:.3f}"
        assert abs(ng - nr) <= GRAD_NORM * nr + 1e-9, f"{what} grad {k}: norm {ng:.3e} vs {nr:.3e}"
    return worst_cos, worst_head


def _inputs(seed, n, h, w, device):
    g = torch.Generator().manual_seed(seed)
    image = torch.randn(n, 3, h, w, generator=g)
    mask = (torch.rand(n, 1, h, w, generator=g) > 0.95).float()
    return image.to(device), mask.to(device)


def _model(arch, sd, device):
    from tiaozhanbei_unet_b200.model import AnomalyUNet, UNet
    m = AnomalyUNet(3) if arch == "anomaly_unet" else UNet(3, 1)
    m.load_state_dict(sd, strict=True)
    return m.to(device)


def _check_out(got, ref, what):
    scale = ref.abs().max().item()
    d = (got - ref).abs()
    assert d.max().item() <= OUT_MAX * scale, f"{what}: max err {d.max().item():.3e} (scale {scale:.3e})"
    assert d.mean().item() <= OUT_MEAN * scale, f"{what}: mean err {d.mean().item():.3e} (scale {scale:.3e})"


@pytest.mark.parametrize("case_name", ["anomaly_unet_2x32x32", "anomaly_unet_1x48x64", "unet_2x32x32", "unet_1x48x64"])
def test_eval_forward_matches_reference_golden(cuda_device, case_name):
    case = torch.load(GOLDEN, weights_only=False)["cases"][case_name]
    arch = case["arch"]
    n, h, w = case["shape"]
    sd = O.synthetic_state_dict(arch, seed=case["weight_seed"])
    m = _model(arch, sd, cuda_device).eval()
    image, _ = _inputs(case["input_seed"], n, h, w, cuda_device)
    with torch.no_grad():
        out = m(image)
    out = out if isinstance(out, tuple) else (out,)
    for i, (got, ref) in enumerate(zip(out, case["eval_out"])):
        ref = ref.to(cuda_device)
        _check_out(got, ref, f"{case_name} eval output {i}")
        if arch == "anomaly_unet":      # masks at threshold 0.5
            _check_masks(got, ref, f"{case_name} output {i}")


def _train_step_ours(m, arch, image, mask, criterion):
    if arch == "anomaly_unet":
        r, a = m(image)
        ld = criterion(r, a, image, mask)
        outs = (r, a)
    else:
        # config 1's step: p = sigmoid(UNet(x)); MSE(p, mask) + focal(p, mask)  
        logits = m(image)
        p = torch.sigmoid(logits)
        rl = torch.nn.functional.mse_loss(p, mask)
        ld = {"total_loss": rl + criterion.focal_loss(p, mask), "recon_loss": rl}
        outs = (logits,)
    ld["total_loss"].backward()
    return outs, {k: float(ld[k].detach()) for k in ("total_loss", "recon_loss")}


@pytest.mark.parametrize("arch,n,h,w", [("anomaly_unet", 2, 32, 32), ("unet", 2, 32, 32), ("anomaly_unet", 3, 64, 48)])
def test_train_step_matches_oracle(cuda_device, arch, n, h, w):
    from tiaozhanbei_unet_b200 import train_utils
    dev = cuda_device
    sd = O.synthetic_state_dict(arch, seed=21, device=dev)
    m = _model(arch, {k: v.clone() for k, v in sd.items()}, dev).train()
    image, mask = _inputs(33, n, h, w, dev)
    crit = train_utils.CombinedLoss()
    opt = train_utils.get_optimizer(m, "adam", 1e-3, 1e-4)
    opt.zero_grad()
    outs, losses = _train_step_ours(m, arch, image, mask, crit)
    # oracle (fp32, same device, TF32 off via the cuda_device fixture)
    ref_sd = {k: v.clone() for k, v in sd.items()}
    opt_state = {}
    ref_losses, ref_grads = O.train_step(ref_sd, image, mask, arch=arch, opt_state=opt_state, step=1)
    torch.cuda.synchronize()
    for k in ("total_loss", "recon_loss"):
        assert abs(losses[k] - ref_losses[k]) <= LOSS_REL * abs(ref_losses[k]), (k, losses[k], ref_losses[k])
    worst = _check_grads(m.named_parameters(), ref_grads, f"{arch} {n}x{h}x{w}")
    # BatchNorm running statistics and counters
    msd = m.state_dict()
    for k in msd:
        if k.endswith("num_batches_tracked"):
            assert int(msd[k]) == 1
        elif O.is_buffer(k):
            d = (msd[k] - ref_sd[k]).abs().max().item()
            assert d <= 5e-2 * (ref_sd[k].abs().max().item() + 1e-3), k
    # one Adam step: update direction is sign-like on step 1, compare in aggregate
    opt.step()
    torch.cuda.synchronize()
    num = den = 0.0
    for k, p in m.named_parameters():
        dp, dr = p.detach() - sd[k], ref_sd[k] - sd[k]
        num += (dp - dr).pow(2).sum().item()
        den += dr.pow(2).sum().item()
    # every element moves by ~lr on Adam's first step; sign flips of near-zero gradients dominate the
    # difference, so the aggregate bound is loose and only guards against a wrong rule / wrong lr
    assert (num / den) ** 0.5 <= 0.9, f"Adam update mismatch {(num / den) ** 0.5:.3e}"
    print(f"[parity] {arch} {n}x{h}x{w}: worst grad cosine {worst[0]:.3f}, worst head-side rel err {worst[1]:.3e}; "
          f"Adam update rel diff {(num / den) ** 0.5:.3f}; losses {losses} vs {ref_losses}")


def test_loss_kernels_match_reference_golden(cuda_device):
    from tiaozhanbei_unet_b200 import train_utils, utils
    dev = cuda_device
    gold = torch.load(GOLDEN, weights_only=False)
    c = gold["combined_loss"]
    g = torch.Generator().manual_seed(c["seed"])
    recon = torch.rand(2, 3, 24, 24, generator=g).to(dev).requires_grad_(True)
    amap = torch.rand(2, 1, 24, 24, generator=g).clamp(1e-4, 1 - 1e-4).to(dev).requires_grad_(True)
    img = torch.randn(2, 3, 24, 24, generator=g).to(dev)
    msk = (torch.rand(2, 1, 24, 24, generator=g) > 0.9).float().to(dev)
    crit = train_utils.CombinedLoss(recon_weight=c["weights"][0], seg_weight=c["weights"][1])
    ld = crit(recon, amap, img, msk)
    ld["total_loss"].backward()
    for k, v in c["values"].items():
        assert abs(ld[k].item() - v) <= 1e-5 * max(abs(v), 1e-3), (k, ld[k].item(), v)
    assert crit.last_values() == tuple(ld[k].item() for k in ("total_loss", "recon_loss", "seg_loss"))
    torch.testing.assert_close(recon.grad.cpu(), c["d_recon"], rtol=1e-4, atol=1e-8)
    torch.testing.assert_close(amap.grad.cpu(), c["d_amap"], rtol=2e-4, atol=1e-8)
    torch.testing.assert_close(utils.compute_anomaly_score(recon, img).cpu(), gold["anomaly_score"]["mse"], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(utils.compute_anomaly_score(recon, img, "l1").cpu(), gold["anomaly_score"]["l1"], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(utils.compute_anomaly_score(recon, img, "ssim").cpu(), gold["anomaly_score"]["mse"], rtol=1e-5, atol=1e-7)
    with pytest.raises(ValueError):
        utils.compute_anomaly_score(recon, img, "nope")
    # SSIM
    s = gold["ssim"]
    g = torch.Generator().manual_seed(s["seed"])
    a = torch.rand(*s["shape"], generator=g)
    b = (a + 0.1 * torch.randn(*s["shape"], generator=g)).clamp(0, 1)
    a, b = a.to(dev).requires_grad_(True), b.to(dev)
    ssim = train_utils.SSIMLoss()
    val = ssim(a, b)
    val.backward()
    assert abs(val.item() - s["value"]) <= 2e-5
    torch.testing.assert_close(a.grad.cpu(), s["grad"], rtol=2e-3, atol=2e-8)
    per = train_utils.SSIMLoss(size_average=False)(a.detach(), b)
    torch.testing.assert_close(per.cpu(), s["per_image"], rtol=1e-4, atol=2e-5)
    assert abs(ssim(b, b).item()) <= 1e-6        # SSIMLoss()(a, a) == 0 in the reference
    # gradient w.r.t. the second image and the per-image variant, against the oracle's autograd
    a2 = a.detach().clone().requires_grad_(True)
    b2 = b.detach().clone().requires_grad_(True)
    w = torch.tensor([0.3, -1.7], device=dev)
    (train_utils.SSIMLoss(size_average=False)(a2, b2) * w).sum().backward()
    a3 = a.detach().clone().requires_grad_(True)
    b3 = b.detach().clone().requires_grad_(True)
    (O.ssim_loss(a3, b3, size_average=False) * w).sum().backward()
    torch.testing.assert_close(a2.grad, a3.grad, rtol=2e-3, atol=2e-8)
    torch.testing.assert_close(b2.grad, b3.grad, rtol=2e-3, atol=2e-8)


@pytest.mark.parametrize("fused", [False, True])
def test_loss_curve_tracks_oracle(cuda_device, fused):
    """100 Adam steps from identical weights on fresh seeded batches: loss curves within 1 % (the bound
    BASELINE.json's north star states) -- through model(x) / criterion / backward / optimizer.step() called one by
    one, and through train_utils.train_step (what train_epoch and bench.py run: one launch sequence with the
    optimizer step inside the backward pass)."""
    from tiaozhanbei_unet_b200 import train_utils
    dev = cuda_device
    arch = "anomaly_unet"
    sd = O.synthetic_state_dict(arch, seed=5, device=dev)
    m = _model(arch, {k: v.clone() for k, v in sd.items()}, dev).train()
    crit = train_utils.CombinedLoss()
    opt = train_utils.get_optimizer(m, "adam", 1e-3, 1e-4)
    ref_sd = {k: v.clone() for k, v in sd.items()}
    opt_state = {}
    g = torch.Generator().manual_seed(0)
    ours, refs = [], []
    for step in range(1, 101):
        image = torch.randn(4, 3, 32, 32, generator=g).to(dev)
        mask = (torch.rand(4, 1, 32, 32, generator=g) > 0.95).float().to(dev)
        if fused:
            assert train_utils._fusable(m, crit, image, mask)
            train_utils.train_step(m, crit, opt, image, mask)
        else:
            r, a = m(image)
            ld = crit(r, a, image, mask)
            opt.zero_grad()
            ld["total_loss"].backward()
            opt.step()
        ours.append(crit.last_values()[0])
        rl, _ = O.train_step(ref_sd, image, mask, arch=arch, opt_state=opt_state, step=step)
        refs.append(rl["total_loss"])
    worst = max(abs(a - b) / abs(b) for a, b in zip(ours, refs))
    if fused:
        assert m._engine().gstage_zero and float(opt.state_dict()["state"][0]["step"]) == 100.0
    print(f"[parity] loss curve ({'train_step' if fused else 'plain'}) worst rel dev {worst:.3e}; first/last ours {ours[0]:.5f}/{ours[-1]:.5f} ref {refs[0]:.5f}/{refs[-1]:.5f}")
    assert worst <= 1e-2


def test_checkpoint_roundtrip_and_modes(cuda_device, tmp_path):
    """state_dict written here loads into the oracle's spec (reference layout) and back; eval after
    training uses the updated running statistics; a ragged last batch gets its own plan."""
    from tiaozhanbei_unet_b200 import train_utils, utils
    dev = cuda_device
    m = _model("unet", O.synthetic_state_dict("unet", seed=2), dev).train()
    opt = train_utils.get_optimizer(m, "adam")
    crit = train_utils.CombinedLoss()
    for n in (3, 2):      # second iteration: smaller (ragged) batch
        image, mask = _inputs(n, n, 32, 32, dev)
        p = torch.sigmoid(m(image))
        loss = crit.focal_loss(p, mask)
        opt.zero_grad()
        loss.backward()
        opt.step()
    path = str(tmp_path / "ck.pth")
    utils.save_checkpoint(m, opt, 3, 0.5, path)
    ck = torch.load(path, map_location="cpu")
    assert list(ck.keys()) == ["epoch", "model_state_dict", "optimizer_state_dict", "loss"]
    spec = O.state_dict_spec("unet")
    assert list(ck["model_state_dict"].keys()) == list(spec.keys())
    assert int(ck["model_state_dict"]["inc.double_conv.1.num_batches_tracked"]) == 2
    st = ck["optimizer_state_dict"]["state"]
    assert len(st) == 64 and set(st[0].keys()) == {"step", "exp_avg", "exp_avg_sq"} and float(st[0]["step"]) == 2.0
    m2 = _model("unet", O.synthetic_state_dict("unet", seed=9), dev)
    opt2 = train_utils.get_optimizer(m2, "adam")
    epoch, loss = utils.load_checkpoint(m2, opt2, path, dev)
    assert (epoch, loss) == (3, 0.5)
    image, _ = _inputs(1, 2, 32, 32, dev)
    m.eval(); m2.eval()
    with torch.no_grad():
        a, b = m(image), m2(image)
    assert torch.equal(a, b)
    # the oracle, fed the checkpoint, agrees with eval mode (running statistics path)
    ref = O.unet_forward({k: v.to(dev) for k, v in ck["model_state_dict"].items()}, image, training=False)
    _check_out(a, ref, "eval after training")
    # resumed optimizer continues identically
    m.train(); m2.train()
    for mm, oo in ((m, opt), (m2, opt2)):
        image, mask = _inputs(7, 2, 32, 32, dev)
        loss = crit.focal_loss(torch.sigmoid(mm(image)), mask)
        oo.zero_grad(); loss.backward(); oo.step()
    # identical state => same update, up to the summation order of the float atomics (BN statistics,
    # split-K) which bf16 rounding then amplifies: a few near-zero-gradient elements may flip the sign of
    # their +-lr Adam move, the bulk must coincide
    diffs = torch.cat([(p - q).abs().flatten() for p, q in zip(m.parameters(), m2.parameters())])
    print(f"[parity] resumed-vs-continued param diff: max {diffs.max().item():.3e} mean {diffs.mean().item():.3e}")
    assert diffs.max().item() <= 2.5e-3 and diffs.mean().item() <= 1e-4, "resumed optimizer diverges"


def test_full_size_config2_against_oracle(cuda_device):
    """BASELINE config 2 shape (AnomalyUNet, B=16, 256x256): forward outputs, loss and a sample of
    gradients against the fp32 oracle on the same device, plus batch independence in eval mode."""
    from tiaozhanbei_unet_b200 import train_utils
    dev = cuda_device
    arch = "anomaly_unet"
    sd = O.synthetic_state_dict(arch, seed=4, device=dev)
    m = _model(arch, {k: v.clone() for k, v in sd.items()}, dev).train()
    image, mask = _inputs(0, 16, 256, 256, dev)
    crit = train_utils.CombinedLoss()
    r, a = m(image)
    ld = crit(r, a, image, mask)
    ld["total_loss"].backward()
    ref_sd = {k: v.clone() for k, v in sd.items()}
    ref_losses, ref_grads = O.train_step(ref_sd, image, mask, arch=arch)
    torch.cuda.synchronize()
    assert abs(ld["total_loss"].item() - ref_losses["total_loss"]) <= LOSS_REL * ref_losses["total_loss"]
    worst = _check_grads(m.named_parameters(), ref_grads, "config-2 shape")
    print(f"[parity] config-2 shape: worst grad cosine {worst[0]:.3f}, worst head-side rel err {worst[1]:.3e}")
    del ref_grads
    m.eval()
    with torch.no_grad():
        full = [t.clone() for t in m(image)]
    # eval-mode batch independence: images 0..7 alone give bit-identical rows
    with torch.no_grad():
        half = m(image[:8].contiguous())
    for f, hlf in zip(full, half):
        assert torch.equal(f[:8], hlf)
    msd = {k: v.clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        ref = O.anomaly_unet_forward(msd, image, training=False)
    for i, (got, rf) in enumerate(zip(full, ref)):
        _check_out(got, rf, f"config-2 eval output {i}")
        _check_masks(got, rf, f"config-2 eval output {i}")


def test_config4_inference_512_sharded(cuda_device):
    """BASELINE config 4 shape class (UNet inference anomaly-map scoring at 512x512; 8 of the 64 images to
    bound test time): sigmoid(UNet(x)) against the fp32 oracle, the src/test.py:88-96 scoring pattern, and
    shard independence — contiguous image shards (what N ranks would each run) give bit-identical rows."""
    from tiaozhanbei_unet_b200 import distributed as D, utils
    dev = cuda_device
    sd = O.synthetic_state_dict("unet", seed=9, device=dev)
    m = _model("unet", {k: v.clone() for k, v in sd.items()}, dev).eval()
    image, _ = _inputs(0, 8, 512, 512, dev)
    with torch.no_grad():
        amap = torch.sigmoid(m(image))
        ref = torch.sigmoid(O.unet_forward({k: v.clone() for k, v in sd.items()}, image, training=False))
    assert amap.shape == (8, 1, 512, 512) and amap.dtype == torch.float32
    _check_out(amap, ref, "config-4 anomaly map")
    _check_masks(amap, ref, "config-4 anomaly map")
    # UNet branch of test_model: reconstruction = images, so the score map is identically zero
    score = utils.compute_anomaly_score(image, image)
    assert score.shape == (8, 512, 512) and float(score.abs().max()) == 0.0
    world = 4
    covered = 0
    for rank in range(world):
        b, e = D.shard_range(8, rank, world)
        with torch.no_grad():
            part = torch.sigmoid(m(image[b:e].contiguous()))
        assert torch.equal(part, amap[b:e]), f"shard {rank} differs from the unsharded run"
        covered += e - b
    assert covered == 8


def test_config5_strip_train_step(cuda_device):
    """BASELINE config 5 shape class (UNet training on a 512x1408 KolektorSDD-shaped strip, one image per
    GPU, grayscale replicated to RGB): loss and gradients of one step against the fp32 oracle."""
    dev = cuda_device
    sd = O.synthetic_state_dict("unet", seed=5, device=dev)
    m = _model("unet", {k: v.clone() for k, v in sd.items()}, dev).train()
    g = torch.Generator().manual_seed(0)
    image = torch.randn(1, 1, 512, 1408, generator=g).expand(-1, 3, -1, -1).contiguous().to(dev)
    mask = (torch.rand(1, 1, 512, 1408, generator=g) > 0.95).float().to(dev)
    p = torch.sigmoid(m(image))
    bce = torch.nn.functional.binary_cross_entropy(p, mask, reduction="none")
    loss = torch.nn.functional.mse_loss(p, mask) + (0.25 * (1 - torch.exp(-bce)) ** 2 * bce).mean()
    loss.backward()
    ref_losses, ref_grads = O.train_step({k: v.clone() for k, v in sd.items()}, image, mask, arch="unet")
    torch.cuda.synchronize()
    assert abs(loss.item() - ref_losses["total_loss"]) <= LOSS_REL * abs(ref_losses["total_loss"]), (loss.item(), ref_losses)
    worst = _check_grads(m.named_parameters(), ref_grads, "config-5 strip")
    print(f"[parity] config-5 strip 1x512x1408: loss {loss.item():.6f} vs {ref_losses['total_loss']:.6f}; "
          f"worst grad cosine {worst[0]:.3f}, worst head-side rel err {worst[1]:.3e}")


def test_odd_sizes_take_the_pad_path(cuda_device):
    """Input sizes that are not multiples of 16 (100x100, 52x76, 37x90): the reference zero-pads the
    up-sampled tensor to the skip's size (src/model.py:57-61); here the consumer kernels read it as zero
    beyond its extent.  Eval outputs against golden vectors of the reference, one training step (loss and
    gradients, floor pooling and its backward included) against the oracle."""
    from tiaozhanbei_unet_b200 import train_utils
    dev = cuda_device
    gold = torch.load(os.path.join(os.path.dirname(GOLDEN), "odd_golden.pt"), weights_only=False)
    for case in gold["cases"]:
        arch = case["arch"]
        sd = O.synthetic_state_dict(arch, seed=case["weights_seed"], device=dev)
        m = _model(arch, {k: v.clone() for k, v in sd.items()}, dev).eval()
        g = torch.Generator().manual_seed(case["input_seed"])
        x = torch.randn(*case["shape"], generator=g).to(dev)
        with torch.no_grad():
            y = m(x)
        outs = [y] if arch == "unet" else list(y)
        for i, (got, want) in enumerate(zip(outs, case["outputs"])):
            assert got.shape == want.shape
            _check_out(got, want.to(dev), f"{arch} {tuple(case['shape'])} output {i}")
    for arch, (n, h, w) in (("anomaly_unet", (2, 52, 76)), ("unet", (2, 100, 100))):
        sd = O.synthetic_state_dict(arch, seed=41, device=dev)
        m = _model(arch, {k: v.clone() for k, v in sd.items()}, dev).train()
        image, mask = _inputs(7, n, h, w, dev)
        crit = train_utils.CombinedLoss()
        outs, losses = _train_step_ours(m, arch, image, mask, crit)
        ref_losses, ref_grads = O.train_step({k: v.clone() for k, v in sd.items()}, image, mask, arch=arch)
        torch.cuda.synchronize()
        assert abs(losses["total_loss"] - ref_losses["total_loss"]) <= LOSS_REL * abs(ref_losses["total_loss"])
        worst = _check_grads(m.named_parameters(), ref_grads, f"{arch} {n}x{h}x{w}")
        print(f"[parity] {arch} {n}x{h}x{w} (pad path): loss {losses['total_loss']:.6f} vs {ref_losses['total_loss']:.6f}; "
              f"worst grad cosine {worst[0]:.3f}, worst head-side rel err {worst[1]:.3e}")


def test_bilinear_decoders(cuda_device):
    """bilinear=True (src/model.py:47-49): the x2 bilinear up-sampling kernels against F.interpolate, eval
    outputs against golden vectors of the reference (incl. an odd size), one training step against the oracle."""
    import torch.nn.functional as F
    from tiaozhanbei_unet_b200 import ops, train_utils
    from tiaozhanbei_unet_b200.model import AnomalyUNet, UNet
    dev = cuda_device
    g = torch.Generator().manual_seed(1)
    for n, h, w, c in ((2, 5, 7, 64), (1, 16, 16, 128), (1, 1, 3, 8)):
        x = torch.randn(n, c, h, w, generator=g).to(dev).to(torch.bfloat16)
        xb = x.permute(0, 2, 3, 1).contiguous()
        out = torch.empty(n, 2 * h, 2 * w, c, dtype=torch.bfloat16, device=dev)
        ops.upsample2x_fwd(xb, out)
        xr = x.float().requires_grad_(True)
        ref = F.interpolate(xr, scale_factor=2, mode="bilinear", align_corners=True)
        torch.cuda.synchronize()
        d = (out.permute(0, 3, 1, 2).float() - ref.detach()).abs().max().item()
        assert d <= 2e-2 * ref.abs().max().item(), (n, h, w, c, d)
        go = torch.randn(n, c, 2 * h, 2 * w, generator=g).to(dev).to(torch.bfloat16)
        ref.backward(go.float())
        dx = torch.empty_like(xb)
        ops.upsample2x_bwd(go.permute(0, 2, 3, 1).contiguous(), dx)
        base = dx.clone()
        ops.upsample2x_bwd(go.permute(0, 2, 3, 1).contiguous(), dx, accumulate=True)
        torch.cuda.synchronize()
        want = xr.grad
        assert (base.permute(0, 3, 1, 2).float() - want).abs().max().item() <= 2e-2 * want.abs().max().item()
        assert (dx.permute(0, 3, 1, 2).float() - 2 * want).abs().max().item() <= 4e-2 * want.abs().max().item()
    gold = torch.load(os.path.join(os.path.dirname(GOLDEN), "bilinear_golden.pt"), weights_only=False)
    for case in gold["cases"]:
        arch = case["arch"]
        m = UNet(3, 1, bilinear=True) if arch == "unet" else AnomalyUNet(3, bilinear=True)
        m.load_state_dict(O.synthetic_state_dict(arch, seed=case["weights_seed"], bilinear=True))
        m = m.to(dev).eval()
        gi = torch.Generator().manual_seed(case["input_seed"])
        x = torch.randn(*case["shape"], generator=gi).to(dev)
        with torch.no_grad():
            y = m(x)
        outs = [y] if arch == "unet" else list(y)
        for i, (got, want) in enumerate(zip(outs, case["outputs"])):
            _check_out(got, want.to(dev), f"bilinear {arch} {tuple(case['shape'])} output {i}")
    arch = "anomaly_unet"
    sd = O.synthetic_state_dict(arch, seed=61, device=dev, bilinear=True)
    m = AnomalyUNet(3, bilinear=True)
    m.load_state_dict({k: v.clone() for k, v in sd.items()})
    m = m.to(dev).train()
    image, mask = _inputs(9, 2, 64, 48, dev)
    crit = train_utils.CombinedLoss()
    outs, losses = _train_step_ours(m, arch, image, mask, crit)
    ref_losses, ref_grads = O.train_step({k: v.clone() for k, v in sd.items()}, image, mask, arch=arch)
    torch.cuda.synchronize()
    assert abs(losses["total_loss"] - ref_losses["total_loss"]) <= LOSS_REL * abs(ref_losses["total_loss"])
    worst = _check_grads(m.named_parameters(), ref_grads, "bilinear anomaly_unet 2x64x48")
    print(f"[parity] bilinear anomaly_unet 2x64x48: loss {losses['total_loss']:.6f} vs {ref_losses['total_loss']:.6f}; "
          f"worst grad cosine {worst[0]:.3f}, worst head-side rel err {worst[1]:.3e}")


@pytest.mark.parametrize("arch,n,h,w,n_classes", [("unet", 1, 16, 16, 1), ("unet", 3, 17, 31, 2), ("anomaly_unet", 1, 16, 48, 1),
                                                  ("unet", 2, 24, 40, 4)])
def test_minimum_and_unusual_shapes(cuda_device, arch, n, h, w, n_classes):
    """Edge shapes: the smallest legal input (16x16: a 1x1 bottleneck), odd sizes just above it, batch 1,
    2- and 4-class heads.  Eval outputs against the fp32 oracle; one training step must give finite
    gradients of the reference's shapes."""
    from tiaozhanbei_unet_b200.model import AnomalyUNet, UNet
    dev = cuda_device
    sd = O.synthetic_state_dict(arch, seed=71, n_classes=n_classes, device=dev)
    m = UNet(3, n_classes) if arch == "unet" else AnomalyUNet(3)
    m.load_state_dict({k: v.clone() for k, v in sd.items()})
    m = m.to(dev).eval()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(n, 3, h, w, generator=g).to(dev)
    with torch.no_grad():
        y = m(x)
        ref = O.unet_forward({k: v.clone() for k, v in sd.items()}, x) if arch == "unet" else \
            O.anomaly_unet_forward({k: v.clone() for k, v in sd.items()}, x)
    outs, refs = ([y], [ref]) if arch == "unet" else (list(y), list(ref))
    for i, (got, want) in enumerate(zip(outs, refs)):
        assert got.shape == want.shape
        _check_out(got, want, f"{arch} {n}x{h}x{w} output {i}")
    m.train()
    out = m(x)
    loss = out.square().mean() if arch == "unet" else (out[0].square().mean() + out[1].mean())
    loss.backward()
    torch.cuda.synchronize()
    for k, p in m.named_parameters():
        assert p.grad is not None and p.grad.shape == p.shape and torch.isfinite(p.grad).all(), k


def test_train_epoch_prefetch_matches_the_step_loop(cuda_device):
    """train_epoch (src/train_utils.py:107-153) over pinned host batches — the next batch's copy runs on a side
    stream during the current step — gives the averages of the plain forward / loss / backward / step loop,
    a ragged last batch and the per-batch weighting of AverageMeter included.  The learning rate is tiny so that
    the comparison is about which batch reached which step, not about Adam amplifying the last-bit differences of
    two runs (atomics) in noise-level gradients."""
    from tiaozhanbei_unet_b200 import train_utils
    dev = cuda_device
    arch = "anomaly_unet"
    sd = O.synthetic_state_dict(arch, seed=7, device=dev)
    g = torch.Generator().manual_seed(3)
    batches = []
    for i, n in enumerate((4, 4, 4, 4, 2)):      # every batch on its own scale: a stale or swapped batch shows
        batches.append({"image": (torch.randn(n, 3, 32, 32, generator=g) * (1.0 + 0.5 * i)).pin_memory(),
                        "mask": (torch.rand(n, 1, 32, 32, generator=g) > 0.9).float().pin_memory()})

    m1 = _model(arch, {k: v.clone() for k, v in sd.items()}, dev)
    c1 = train_utils.CombinedLoss()
    o1 = train_utils.get_optimizer(m1, "adam", 1e-7, 0.0)
    got = train_utils.train_epoch(m1, batches, c1, o1, dev, 1)

    m2 = _model(arch, {k: v.clone() for k, v in sd.items()}, dev).train()
    c2 = train_utils.CombinedLoss()
    o2 = train_utils.get_optimizer(m2, "adam", 1e-7, 0.0)
    tot = [0.0, 0.0, 0.0]
    count = 0
    for b in batches:
        image, mask = b["image"].to(dev), b["mask"].to(dev)
        r, a = m2(image)
        ld = c2(r, a, image, mask)
        o2.zero_grad()
        ld["total_loss"].backward()
        o2.step()
        n = image.size(0)
        for i, v in enumerate(c2.last_values()):
            tot[i] += v * n
        count += n
    for key, want in zip(("total_loss", "recon_loss", "seg_loss"), tot):
        want /= count
        # (two runs of the same step differ in the last bits — atomics — and bf16 storage turns some of those into
        # rounding flips: the small focal term moves by a few 1e-4; a stale batch would move the reconstruction
        # term, whose batches differ by up to 9x, by tens of percent)
        tol = 5e-3 if key == "seg_loss" else 1e-3
        assert abs(got[key] - want) <= tol * abs(want), (key, got[key], want)
    # every batch is different, so a wrong / stale batch in any step would show: the five per-step losses differ
    assert len({round(float(((b["image"] ** 2).mean())), 6) for b in batches}) == len(batches)
    for (k, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert torch.allclose(p1, p2, rtol=0, atol=2e-6), k


def test_gradient_accumulation_adds_up(cuda_device):
    """Two backward passes before the optimizer step (gradient accumulation, or zero_grad(set_to_none=False)):
    `.grad` is a view of the engine's flat gradient buffer, which the second backward rewrites — the engine must
    hand autograd the new gradients in fresh storage so that AccumulateGrad leaves old + new, here checked against
    the oracle's two gradients (head-side tensors tightly, the rest by direction / norm like every gradient test)."""
    from tiaozhanbei_unet_b200 import train_utils
    dev = cuda_device
    arch = "anomaly_unet"
    sd = O.synthetic_state_dict(arch, seed=5, device=dev)
    m = _model(arch, {k: v.clone() for k, v in sd.items()}, dev).train()
    crit = train_utils.CombinedLoss()
    opt = train_utils.get_optimizer(m, "adam", 1e-3, 0.0)
    batches = [_inputs(11, 2, 32, 32, dev), _inputs(12, 2, 32, 32, dev)]
    batches[1] = (batches[1][0] * 2.5 + 0.5, (batches[1][0][:, :1] > 0.3).float())      # a clearly different second batch
    opt.zero_grad()
    firsts = None
    for image, mask in batches:
        _train_step_ours(m, arch, image, mask, crit)
        if firsts is None:
            firsts = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    torch.cuda.synchronize()
    eng = m._engine()
    # the accumulated gradients still live in the flat buffer (one fused optimizer launch)
    assert all(p.grad.untyped_storage().data_ptr() == eng.flat_g.untyped_storage().data_ptr() for p in m.parameters())
    # running statistics differ between the two oracle steps: replay them in order on one state dict
    ref_sd = {k: v.clone() for k, v in sd.items()}
    total = None
    for image, mask in batches:
        _, g = O.train_step(ref_sd, image, mask, arch=arch)
        total = g if total is None else {k: total[k] + g[k] for k in g}
    for k, p in m.named_parameters():        # (a sum of two bf16-path gradients: the head-side bound is twice one step's)
        g, gr = p.grad, total[k]
        assert torch.isfinite(g).all(), k
        cos = ((g * gr).sum() / (g.norm() * gr.norm() + 1e-30)).item()
        assert cos >= GRAD_COS, f"accumulated grad {k}: cosine {cos:.3f}"
        if k.startswith(HEAD_KEYS):
            assert (g - gr).norm() <= 2 * GRAD_REL_HEAD * gr.norm(), f"accumulated grad {k}"
    # and it is NOT twice the last gradient nor the last gradient alone (the failure modes of aliased storage):
    # checked on the head-side tensor whose two gradients differ most (the second batch is scaled to make them differ)
    _, g_last = O.train_step({kk: v.clone() for kk, v in sd.items()}, *batches[1], arch=arch)
    named = dict(m.named_parameters())
    heads = [kk for kk in named if kk.startswith(HEAD_KEYS)]
    k = max(heads, key=lambda kk: ((firsts[kk] - g_last[kk]).norm() / g_last[kk].norm()).item())
    got = named[k].grad
    assert (firsts[k] - g_last[k]).norm() > 0.2 * g_last[k].norm(), "test inputs too similar to tell the cases apart"
    assert (got - 2 * g_last[k]).norm() > 0.1 * got.norm() and (got - g_last[k]).norm() > 0.1 * got.norm()
    assert (got - firsts[k] - g_last[k]).norm() <= 5e-2 * got.norm()
    # zero_grad(set_to_none=False) keeps the views; the next backward must give that step's gradient alone
    opt.zero_grad(set_to_none=False)
    _train_step_ours(m, arch, *batches[0], crit)
    torch.cuda.synchronize()
    assert (dict(m.named_parameters())[k].grad - firsts[k]).norm() <= 5e-2 * firsts[k].norm()
    opt.step()
    torch.cuda.synchronize()


def test_backward_after_any_other_forward_raises(cuda_device):
    """The BatchNorm statistics / scale / shift and the packed weights are per engine: a forward of ANOTHER shape or
    an eval forward between a training forward and its backward invalidates them, and backward must say so instead
    of using the wrong statistics."""
    from tiaozhanbei_unet_b200 import train_utils
    dev = cuda_device
    m = _model("anomaly_unet", O.synthetic_state_dict("anomaly_unet", seed=2), dev).train()
    crit = train_utils.CombinedLoss()
    image, mask = _inputs(1, 2, 32, 32, dev)
    other, _ = _inputs(2, 1, 48, 32, dev)
    for intruder in ("train_other_shape", "eval_same_shape"):
        m.train()
        r, a = m(image)
        loss = crit(r, a, image, mask)["total_loss"]
        if intruder == "train_other_shape":
            with torch.no_grad():
                m(other)
        else:
            m.eval()
            with torch.no_grad():
                m(image)
            m.train()
        with pytest.raises(RuntimeError, match="another forward"):
            loss.backward()
    # and the normal order still works afterwards
    r, a = m(image)
    crit(r, a, image, mask)["total_loss"].backward()
    torch.cuda.synchronize()


def test_plan_cache_is_lru(cuda_device):
    """Four live input shapes (train, ragged train, validation, ragged validation) must not evict the training plan
    every epoch: the cache holds UNETK_MAX_PLANS (4) plans and drops the least recently used one."""
    from tiaozhanbei_unet_b200 import engine as E
    dev = cuda_device
    m = _model("unet", O.synthetic_state_dict("unet", seed=2), dev).eval()
    eng = m._engine()
    shapes = [(2, 32, 32), (1, 32, 32), (2, 16, 48), (1, 16, 48)]
    assert E.MAX_PLANS >= 4
    with torch.no_grad():
        for n, h, w in shapes:
            m(torch.zeros(n, 3, h, w, device=dev))
        train_plan = eng.plans[shapes[0]]
        m(torch.zeros(*((shapes[0][0], 3) + shapes[0][1:]), device=dev))       # touch the oldest: now most recent
        for i in range(E.MAX_PLANS - 3):
            m(torch.zeros(1, 3, 16 * (i + 1), 64, device=dev))                 # new shapes until one must go
        m(torch.zeros(3, 3, 16, 16, device=dev))
    assert eng.plans.get(shapes[0]) is train_plan, "the most recently used plan was evicted"
    assert shapes[1] not in eng.plans, "the least recently used plan should have been evicted"
    assert len(eng.plans) == E.MAX_PLANS
    torch.cuda.synchronize()


@pytest.mark.parametrize("n,h,w", [(2, 32, 32), (3, 64, 48), (1, 36, 52)])
def test_fused_train_step_matches_the_plain_path(cuda_device, n, h, w):
    """train_utils.train_step on this package's AnomalyUNet + CombinedLoss runs forward, loss and backward as one
    launch sequence with each decoder's head, loss term and head backward in ONE kernel (csrc/head_fused.cu).  The
    arithmetic is the plain path's (model(x) -> criterion -> backward), element for element, so losses, the
    gradients next to the heads, the BatchNorm buffers and three optimizer steps (eager call, graph capture, graph
    replay) must agree with it; the deep gradients are compared like everywhere else (direction / norm: two runs
    of the SAME path already differ there through the order of the fp32 atomics)."""
    from tiaozhanbei_unet_b200 import train_utils
    dev = cuda_device
    arch = "anomaly_unet"
    sd = O.synthetic_state_dict(arch, seed=9, device=dev)
    batches = [_inputs(40 + i, n, h, w, dev) for i in range(3)]

    def run(fused):
        m = _model(arch, {k: v.clone() for k, v in sd.items()}, dev).train()
        crit = train_utils.CombinedLoss(recon_weight=0.7, seg_weight=1.3)
        # (a tiny learning rate: Adam's first steps move every weight by +-lr whatever the gradient's size, which
        # turns the last-bit differences of two runs -- fp32 atomics -- into different weights)
        opt = train_utils.get_optimizer(m, "adam", 1e-7, 1e-4)
        losses, first_grads = [], None
        for image, mask in batches:
            if fused:
                assert train_utils._fusable(m, crit, image, mask)
                ld = train_utils.train_step(m, crit, opt, image, mask)
            else:
                r, a = m(image)
                ld = crit(r, a, image, mask)
                opt.zero_grad()
                ld["total_loss"].backward()
                opt.step()
            losses.append([float(ld[k]) for k in ("total_loss", "recon_loss", "seg_loss")])
            assert tuple(losses[-1]) == pytest.approx(crit.last_values(), rel=1e-6)
            if first_grads is None:
                first_grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
        torch.cuda.synchronize()
        return m, losses, first_grads

    m_plain, l_plain, g_plain = run(False)
    m_fused, l_fused, g_fused = run(True)
    for a, b in zip(l_plain, l_fused):
        # (two runs of the SAME path differ in the last bits -- atomics -- and bf16 storage turns some of those into
        # rounding flips: the small focal term moves by a few 1e-3, see test_train_epoch_prefetch_matches_the_step_loop)
        for va, vb, tol in zip(a, b, (1e-3, 1e-3, 5e-3)):
            assert abs(va - vb) <= tol * abs(va), (l_plain, l_fused)
    assert abs(l_plain[0][0] - l_fused[0][0]) <= 5e-4 * abs(l_plain[0][0])     # same weights, same arithmetic, other atomics order
    assert abs(l_fused[0][0] - (0.7 * l_fused[0][1] + 1.3 * l_fused[0][2])) <= 1e-5 * abs(l_fused[0][0])
    for k in g_plain:
        gp, gf = g_plain[k], g_fused[k]
        cos = ((gp * gf).sum() / (gp.norm() * gf.norm() + 1e-30)).item()
        assert cos >= GRAD_COS and abs(gp.norm().item() - gf.norm().item()) <= GRAD_NORM * gp.norm().item(), (k, cos)
        if k.startswith(HEAD_KEYS):
            # (not tighter: the batch statistics are sums of fp32 atomics, their last bits move the BatchNorm scale /
            # shift, and an activation that rounds to +0 instead of a tiny positive value flips its ReLU mask; the
            # kernels themselves are compared on identical inputs in test_ops_gpu.py::test_head_loss_bwd_is_the_unfused_sequence)
            assert (gp - gf).norm() <= GRAD_REL_HEAD * gp.norm(), k
    sp, sf = m_plain.state_dict(), m_fused.state_dict()
    for k in sp:
        if k.endswith("num_batches_tracked"):
            assert int(sp[k]) == int(sf[k]) == 3
        elif O.is_buffer(k):      # (the bound of test_train_step_matches_oracle)
            assert (sp[k] - sf[k]).abs().max().item() <= 5e-2 * (sp[k].abs().max().item() + 1e-3), k
    # and against the oracle, like the plain path's own test
    ref_losses, ref_grads = O.train_step({k: v.clone() for k, v in sd.items()}, *batches[0], arch=arch)
    want = 0.7 * ref_losses["recon_loss"] + 1.3 * ref_losses["seg_loss"]
    assert abs(l_fused[0][0] - want) <= LOSS_REL * abs(want)
    assert abs(l_fused[0][1] - ref_losses["recon_loss"]) <= LOSS_REL * abs(ref_losses["recon_loss"])
    for k in ("outc_recon.conv.bias", "outc_seg.conv.bias", "outc_seg.conv.weight"):
        gr = ref_grads[k] * (0.7 if "recon" in k else 1.3)
        assert (g_fused[k] - gr).norm() <= GRAD_REL_HEAD * gr.norm(), k


def test_fused_train_step_falls_back_when_it_does_not_apply(cuda_device):
    """UNet, a frozen parameter, eval mode or a foreign criterion take the plain path of train_step."""
    from tiaozhanbei_unet_b200 import train_utils
    dev = cuda_device
    image, mask = _inputs(3, 2, 32, 32, dev)
    crit = train_utils.CombinedLoss()
    m = _model("anomaly_unet", O.synthetic_state_dict("anomaly_unet", seed=2), dev).train()
    assert train_utils._fusable(m, crit, image, mask)
    assert not train_utils._fusable(m, torch.nn.MSELoss(), image, mask)
    assert not train_utils._fusable(m.eval(), crit, image, mask)
    m.train()
    m.outc_seg.conv.bias.requires_grad_(False)
    assert not train_utils._fusable(m, crit, image, mask)
    opt = train_utils.get_optimizer(m, "adam", 1e-3, 1e-4)
    ld = train_utils.train_step(m, crit, opt, image, mask)          # plain path: autograd leaves the frozen bias alone
    torch.cuda.synchronize()
    assert m.outc_seg.conv.bias.grad is None and m.outc_seg.conv.weight.grad is not None
    assert torch.isfinite(ld["total_loss"]).item()


def test_fused_train_step_can_return_the_outputs(cuda_device):
    """Engine.train_step(want_outs=True) also stores the two sigmoid outputs (a caller that logs reconstructions):
    they are the plain training-mode forward's, and the gradients do not change."""
    from tiaozhanbei_unet_b200 import train_utils
    dev = cuda_device
    sd = O.synthetic_state_dict("anomaly_unet", seed=4, device=dev)
    image, mask = _inputs(8, 2, 48, 32, dev)
    cfg = (1.0, 1.0, 0.25, 2.0)
    m1 = _model("anomaly_unet", {k: v.clone() for k, v in sd.items()}, dev).train()
    r, a = m1(image)
    m2 = _model("anomaly_unet", {k: v.clone() for k, v in sd.items()}, dev).train()
    for _ in range(3):      # eager, capture, replay
        m2.load_state_dict(sd)          # (same weights and running statistics every time)
        out3, outs = m2._engine().train_step(image, mask, cfg, want_outs=True)
        torch.cuda.synchronize()
        # (two runs of the same forward differ through the atomics' order, amplified by bf16 storage: the bounds of
        # every other output comparison in this file)
        _check_out(outs[0], r.detach(), "train_step reconstruction")
        _check_out(outs[1], a.detach(), "train_step anomaly map")
        ld = train_utils.CombinedLoss()(r.detach(), a.detach(), image, mask)
        assert abs(float(out3[0]) - float(ld["total_loss"])) <= 1e-3 * float(ld["total_loss"])
    m3 = _model("anomaly_unet", {k: v.clone() for k, v in sd.items()}, dev).train()
    m3._engine().train_step(image, mask, cfg)
    torch.cuda.synchronize()
    for (k, p2), (_, p3) in zip(m2.named_parameters(), m3.named_parameters()):
        if k.startswith(HEAD_KEYS):
            assert (p2.grad - p3.grad).norm() <= GRAD_REL_HEAD * p3.grad.norm(), k


@pytest.mark.parametrize("kind", ["adam", "adamw", "sgd"])
def test_fused_update_is_the_unfused_sequence(cuda_device, kind):
    """unetk_fused_update (one pass per parameter group: staged gradient -> `.grad`, optimizer update, new weights ->
    packed bf16 layouts, staging range zeroed) against the three passes it replaces on the same buffers
    (unpack_conv_grads / permute_batched, adam_step, pack_conv_weights / permute_batched / first_conv_pack):
    bit for bit, over every tensor of an AnomalyUNet, for two consecutive steps (first-step flag, bias corrections)
    and a gradient scale != 1 (the data-parallel average)."""
    from tiaozhanbei_unet_b200 import ops, train_utils
    dev = cuda_device
    m = _model("anomaly_unet", O.synthetic_state_dict("anomaly_unet", seed=5), dev).train()
    eng = m._engine()
    eng.ensure(torch.device(dev))
    opt_a = train_utils.get_optimizer(m, kind, 3e-3, 1e-2)
    opt_b = train_utils.get_optimizer(m, kind, 3e-3, 1e-2)
    g = torch.Generator(device=dev).manual_seed(11)
    p0 = eng.flat_p.clone()
    scale = 0.25
    ref = {}
    # --- the separate passes (twice), starting from p0 ---
    stages = [torch.randn(eng.total, device=dev, generator=g) * 1e-2 for _ in range(2)]
    for step in range(2):
        eng.gstage.copy_(stages[step])
        for part in (2, 1, 0):
            eng._unpack(part, scale)
        for p, v in zip(eng.params, eng.grad_views()):
            p.grad = v
        ref[("g", step)] = eng.flat_g.clone()
        opt_a.step()
        eng.weights_dirty = True
        eng._pack_weights(split=False)
        ref[("p", step)], ref[("w", step)] = eng.flat_p.clone(), eng.wpack.clone()
        ref[("m", step)] = [b.clone() for b in opt_a._flat]
    torch.cuda.synchronize()
    # --- the fused pass (twice), from the same start ---
    eng.flat_p.copy_(p0)
    eng.wpack.zero_()
    eng.flat_g.zero_()
    for p in eng.params:
        p.grad = None
    for step in range(2):
        eng.gstage.copy_(stages[step])
        upd = opt_b.fused_update_args(eng)
        assert upd is not None
        vals = list(upd.values)
        vals[8] = scale
        ops.set_floats(eng.hyper, vals)
        eng._upd = upd
        for tag in (("dec", 1), ("dec", 0), ("enc", 3), ("enc", 0)):        # the order the backward pass finishes them in
            eng._update_segment(tag)
        eng._upd = None
        torch.cuda.synchronize()
        assert torch.equal(eng.flat_g, ref[("g", step)]), f"{kind} step {step}: gradients"
        assert torch.equal(eng.flat_p, ref[("p", step)]), f"{kind} step {step}: parameters"
        for a, b in zip(opt_b._flat, ref[("m", step)]):
            assert torch.equal(a, b), f"{kind} step {step}: optimizer state"
        assert torch.equal(eng.wpack.view(torch.int16), ref[("w", step)].view(torch.int16)), f"{kind} step {step}: packed weights"
        for k, p in zip(eng.param_names, eng.params):       # every parameter's staging range is zero again
            o = eng.offsets[k]
            assert not eng.gstage[o: o + p.numel()].any(), k
    sa, sb = opt_a.state_dict(), opt_b.state_dict()
    if kind != "sgd":
        assert float(sa["state"][0]["step"]) == float(sb["state"][0]["step"]) == 2.0


@pytest.mark.parametrize("kind,bilinear,shape", [("adam", False, (2, 32, 48)), ("adamw", True, (2, 32, 32)), ("sgd", False, (1, 36, 52))])
def test_fused_update_inside_train_step(cuda_device, kind, bilinear, shape):
    """train_utils.train_step with this package's optimizer: the optimizer step runs inside the backward pass
    (Engine.train_step(update=...)).  Deterministic invariants after eager, captured and replayed steps: the packed
    bf16 weights the next step will read are exactly pack(current fp32 weights), the staging buffer is zero, the
    step count advances, a learning-rate change and weights written by torch code in between (load_state_dict)
    are picked up; and the trajectory agrees with the separate optimizer.step() path like two runs of one path do.
    Adam / AdamW / SGD; transposed-conv and bilinear (parameter-free up-sampling) decoders; an input size that is not
    a multiple of 16 (serial decoder chains in backward)."""
    from tiaozhanbei_unet_b200 import train_utils
    from tiaozhanbei_unet_b200.model import AnomalyUNet
    dev = cuda_device
    arch = "anomaly_unet"
    sd = O.synthetic_state_dict(arch, seed=12, device=dev, bilinear=bilinear)
    batches = [_inputs(70 + i, *shape, dev) for i in range(5)]
    lr = 1e-7 if kind != "sgd" else 1e-9        # (SGD moves a weight by lr * gradient, Adam by ~lr)

    def run(fused_update):
        old = train_utils.FUSED_UPDATE
        train_utils.FUSED_UPDATE = fused_update
        try:
            m = AnomalyUNet(3, bilinear=bilinear)
            m.load_state_dict({k: v.clone() for k, v in sd.items()})
            m = m.to(dev).train()
            crit = train_utils.CombinedLoss()
            opt = train_utils.get_optimizer(m, kind, lr, 1e-4)
            eng = m._engine()
            losses = []
            for i, (image, mask) in enumerate(batches):
                if i == 3:
                    opt.param_groups[0]["lr"] = 2 * lr
                    m.load_state_dict({k: v.clone() for k, v in sd.items()})      # torch writes the weights in place
                ld = train_utils.train_step(m, crit, opt, image, mask)
                losses.append(float(ld["total_loss"]))
                if fused_update:
                    torch.cuda.synchronize()
                    assert eng.gstage_zero and not eng.weights_dirty
                    assert not eng.gstage.any(), f"step {i}: staging buffer not zero"
                    have = eng.wpack.clone()
                    eng.weights_dirty = True
                    eng._pack_weights(split=False)
                    torch.cuda.synchronize()
                    assert torch.equal(have.view(torch.int16), eng.wpack.view(torch.int16)), f"step {i}: stale packed weights"
                    assert all(p.grad is not None and p.grad.data_ptr() == eng.flat_g.data_ptr() + 4 * eng.offsets[k]
                               for k, p in zip(eng.param_names, eng.params))
            torch.cuda.synchronize()
            if kind != "sgd":
                assert float(opt.state_dict()["state"][0]["step"]) == len(batches)
            return m, losses
        finally:
            train_utils.FUSED_UPDATE = old

    m_sep, l_sep = run(False)
    m_fus, l_fus = run(True)
    for a, b in zip(l_sep, l_fus):
        assert abs(a - b) <= 1e-3 * abs(a), (l_sep, l_fus)
    # steps 0 and 3 start from the same weights (and BatchNorm buffers close to each other): same loss up to the atomics
    assert abs(l_fus[0] - l_sep[0]) <= 5e-4 * abs(l_sep[0])
    moved = 0.0
    for (k, a), (_, b) in zip(m_sep.named_parameters(), m_fus.named_parameters()):
        # lr 1e-7..2e-7, five steps: every weight moved by at most ~1e-6; a stale or mis-laid-out update would show
        assert (a - b).abs().max().item() <= 4e-6, k
        moved = max(moved, (a - sd[k]).abs().max().item())
    assert moved > 0


@pytest.mark.parametrize("arch", ["unet", "anomaly_unet"])
def test_predict_batches_is_the_batch_loop(cuda_device, arch):
    """utils.predict_batches (the body of the reference's test_model loop, src/test.py:80-106, as a copy-in / compute /
    copy-out pipeline over a host tensor) returns exactly what the batch-by-batch loop returns: ragged last batch,
    slots reused more than once, both model kinds."""
    from tiaozhanbei_unet_b200 import utils
    dev = cuda_device
    sd = O.synthetic_state_dict(arch, seed=6, device=dev)
    m = _model(arch, {k: v.clone() for k, v in sd.items()}, dev).eval()
    g = torch.Generator().manual_seed(21)
    host = torch.randn(11, 3, 48, 80, generator=g)
    got = utils.predict_batches(m, host, batch_size=3, want_reconstruction=True)
    torch.cuda.synchronize()
    assert got["anomaly_maps"].is_pinned() and got["anomaly_maps"].shape == (11, 1, 48, 80)
    assert got["anomaly_scores"].shape == (11, 48, 80)
    with torch.no_grad():
        for b0 in range(0, 11, 3):
            x = host[b0: b0 + 3].to(dev)
            if arch == "anomaly_unet":
                recon, amap = m(x)
            else:
                amap, recon = torch.sigmoid(m(x)), x
            score = utils.compute_anomaly_score(recon, x)
            assert torch.equal(got["anomaly_maps"][b0: b0 + 3], amap.cpu()), b0
            assert torch.equal(got["anomaly_scores"][b0: b0 + 3], score.cpu()), b0
            assert torch.equal(got["reconstructions"][b0: b0 + 3], recon.cpu()), b0
    if arch == "unet":
        assert float(got["anomaly_scores"].abs().max()) == 0.0       # reconstruction = images (src/test.py:93)
    with pytest.raises(RuntimeError):
        utils.predict_batches(m.train(), host)


def test_train_validate_train_cycle(cuda_device):
    """The loops of src/train.py:229-244 back to back: train_epoch (fused step, optimizer inside it) -> validate_epoch
    (eval forward: must read the weights the last in-step update packed) -> train_epoch again.  The eval outputs after
    training equal those of a FRESH model loaded from the trained state_dict bit for bit (eval mode has no atomics), so
    the packed bf16 copies left behind by the fused update are exactly what a re-pack of the fp32 weights gives; the
    second epoch continues from the first one's optimizer state."""
    from tiaozhanbei_unet_b200 import train_utils
    from tiaozhanbei_unet_b200.model import AnomalyUNet
    dev = cuda_device
    sd = O.synthetic_state_dict("anomaly_unet", seed=15, device=dev)
    m = _model("anomaly_unet", {k: v.clone() for k, v in sd.items()}, dev)
    crit = train_utils.CombinedLoss()
    opt = train_utils.get_optimizer(m, "adam", 1e-4, 1e-4)
    g = torch.Generator().manual_seed(3)

    def batches(n, bs):
        out = []
        for i in range(n):
            image = torch.randn(bs, 3, 32, 48, generator=g)
            mask = (torch.rand(bs, 1, 32, 48, generator=g) > 0.9).float()
            # (label 0 everywhere: with mixed labels the reference's validate_epoch ends in a ValueError, src/train_utils.py:216-219,
            #  which this package reproduces -- tests/test_eval_gpu.py)
            out.append({"image": image, "mask": mask, "label": torch.zeros(bs, dtype=torch.int64)})
        return out

    train_a, val, train_b = batches(4, 2), batches(2, 3), batches(3, 2)
    r1 = train_utils.train_epoch(m, train_a, crit, opt, dev, 0)
    eng = m._engine()
    assert eng.gstage_zero and not eng.weights_dirty and float(opt.state_dict()["state"][0]["step"]) == 4.0
    v1 = train_utils.validate_epoch(m, val, crit, dev)
    assert not m.training and np.isfinite(v1["total_loss"])
    # a fresh model with the trained weights gives the same evaluation outputs, bit for bit
    fresh = AnomalyUNet(3)
    fresh.load_state_dict({k: v.clone() for k, v in m.state_dict().items()})
    fresh = fresh.to(dev).eval()
    x = val[0]["image"].to(dev)
    with torch.no_grad():
        ra, aa = m(x)
        rb, ab = fresh(x)
    assert torch.equal(ra, rb) and torch.equal(aa, ab), "eval after the in-step update read stale packed weights"
    r2 = train_utils.train_epoch(m, train_b, crit, opt, dev, 1)
    torch.cuda.synchronize()
    assert m.training and float(opt.state_dict()["state"][0]["step"]) == 7.0
    assert all(np.isfinite(v) for v in list(r1.values()) + list(r2.values()))
    assert eng.gstage_zero and not eng.gstage.any()


def test_checkpoint_resume_with_the_in_step_optimizer(cuda_device, tmp_path):
    """save_checkpoint / load_checkpoint (src/utils.py:37-58) around fused training steps: the optimizer state the
    in-step update keeps in flat buffers is what state_dict() writes; a new model + optimizer that loads the checkpoint
    continues exactly where the first one stood (same moments, same step count), and the next fused step of both moves
    the weights alike."""
    from tiaozhanbei_unet_b200 import train_utils, utils
    dev = cuda_device
    sd = O.synthetic_state_dict("anomaly_unet", seed=18, device=dev)
    batches = [_inputs(90 + i, 2, 32, 32, dev) for i in range(3)]
    m1 = _model("anomaly_unet", {k: v.clone() for k, v in sd.items()}, dev).train()
    crit = train_utils.CombinedLoss()
    o1 = train_utils.get_optimizer(m1, "adam", 1e-6, 1e-4)
    for image, mask in batches[:2]:
        train_utils.train_step(m1, crit, o1, image, mask)
    torch.cuda.synchronize()
    path = str(tmp_path / "ckpt.pth")
    utils.save_checkpoint(m1, o1, 2, 0.5, path)
    m2 = _model("anomaly_unet", {k: v.clone() for k, v in sd.items()}, dev).train()
    o2 = train_utils.get_optimizer(m2, "adam", 1e-3, 0.0)            # (hyper-parameters come back from the checkpoint)
    epoch, loss = utils.load_checkpoint(m2, o2, path, dev)
    assert (epoch, loss) == (2, 0.5) and o2.param_groups[0]["lr"] == 1e-6 and o2.param_groups[0]["weight_decay"] == 1e-4
    s1, s2 = o1.state_dict()["state"], o2.state_dict()["state"]
    assert len(s1) == len(s2) == 98
    for i in s1:
        assert float(s1[i]["step"]) == float(s2[i]["step"]) == 2.0
        assert torch.equal(s1[i]["exp_avg"], s2[i]["exp_avg"].to(dev)) and torch.equal(s1[i]["exp_avg_sq"], s2[i]["exp_avg_sq"].to(dev))
    for (k, a), (_, b) in zip(m1.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    image, mask = batches[2]
    l1 = train_utils.train_step(m1, crit, o1, image, mask)
    l2 = train_utils.train_step(m2, crit, o2, image, mask)
    torch.cuda.synchronize()
    assert abs(float(l1["total_loss"]) - float(l2["total_loss"])) <= 5e-4 * abs(float(l1["total_loss"]))
    assert float(o1.state_dict()["state"][0]["step"]) == float(o2.state_dict()["state"][0]["step"]) == 3.0
    e2 = m2._engine()
    assert e2.gstage_zero and o2.state[e2.params[0]]["exp_avg"].data_ptr() == o2._flat[0].data_ptr()     # moved into the flat buffers
    for (k, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
        # one Adam step of 1e-6 each, from identical weights and moments: at most the sign of a few updates differs
        # (|update| <= ~lr each; a checkpoint that failed to restore lr = 1e-6 or the moments would show as >= 1e-4)
        assert (a - b).abs().max().item() <= 5e-6, k
