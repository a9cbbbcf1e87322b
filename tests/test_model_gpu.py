"""Network-level parity of the CUDA path (model -> engine -> C ABI) against the oracle and the golden
vectors recorded from the reference, on identical seeded inputs and weights.

Tolerances (bf16 storage of activations / gradients, fp32 accumulation; SURVEY.md §7.2 item 8 and
BASELINE.md §5 calibrated the reference itself under bf16 autocast at ~1.2e-2 of the output range):
  * sigmoid outputs and UNet logits: max |d| <= 3e-2 * max|ref| (outputs), mean |d| <= 4e-3 * max|ref|
  * losses: relative 5e-3
  * parameter gradients: ||g - g_ref|| <= 8e-2 * ||g_ref|| per tensor (+ a floor for near-zero grads)
  * threshold masks: bit-exact wherever the reference margin |p - thr| exceeds 3e-2
"""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import unet_oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "unet_golden.pt")
OUT_MAX, OUT_MEAN, LOSS_REL, GRAD_REL, MARGIN = 3e-2, 4e-3, 5e-3, 8e-2, 3e-2


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
        if arch == "anomaly_unet":      # masks at threshold 0.5, bit-exact outside the margin
            safe = (ref - 0.5).abs() > MARGIN
            assert torch.equal((got > 0.5)[safe], (ref > 0.5)[safe])


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
    return outs, {k: float(ld[k]) for k in ("total_loss", "recon_loss")}


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
    worst = (0.0, None)
    for k, p in m.named_parameters():
        g, gr = p.grad, ref_grads[k]
        assert g is not None and g.shape == gr.shape, k
        err = (g - gr).norm().item()
        ref_n = gr.norm().item()
        rel = err / (ref_n + 1e-12)
        if rel > worst[0]:
            worst = (rel, k)
        assert err <= GRAD_REL * ref_n + 1e-7, f"grad {k}: rel err {rel:.3e} (|g_ref| {ref_n:.3e})"
    # BatchNorm running statistics and counters
    msd = m.state_dict()
    for k in msd:
        if k.endswith("num_batches_tracked"):
            assert int(msd[k]) == 1
        elif O.is_buffer(k):
            d = (msd[k] - ref_sd[k]).abs().max().item()
            assert d <= 2e-2 * (ref_sd[k].abs().max().item() + 1e-3), k
    # one Adam step: update direction is sign-like on step 1, compare in aggregate
    opt.step()
    torch.cuda.synchronize()
    num = den = 0.0
    for k, p in m.named_parameters():
        dp, dr = p.detach() - sd[k], ref_sd[k] - sd[k]
        num += (dp - dr).pow(2).sum().item()
        den += dr.pow(2).sum().item()
    assert (num / den) ** 0.5 <= 0.2, f"Adam update mismatch {(num / den) ** 0.5:.3e}"
    print(f"[parity] {arch} {n}x{h}x{w}: worst grad rel err {worst[0]:.3e} at {worst[1]}; losses {losses} vs {ref_losses}")


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


def test_loss_curve_tracks_oracle(cuda_device):
    """20 Adam steps from identical weights on fresh seeded batches: loss curves within 1 %."""
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
    for step in range(1, 21):
        image = torch.randn(4, 3, 32, 32, generator=g).to(dev)
        mask = (torch.rand(4, 1, 32, 32, generator=g) > 0.95).float().to(dev)
        r, a = m(image)
        ld = crit(r, a, image, mask)
        opt.zero_grad()
        ld["total_loss"].backward()
        opt.step()
        ours.append(crit.last_values()[0])
        rl, _ = O.train_step(ref_sd, image, mask, arch=arch, opt_state=opt_state, step=step)
        refs.append(rl["total_loss"])
    worst = max(abs(a - b) / abs(b) for a, b in zip(ours, refs))
    print(f"[parity] loss curve worst rel dev {worst:.3e}; first/last ours {ours[0]:.5f}/{ours[-1]:.5f} ref {refs[0]:.5f}/{refs[-1]:.5f}")
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
    for (k, p), (_, q) in zip(m.named_parameters(), m2.named_parameters()):
        assert torch.allclose(p, q, rtol=0, atol=1e-6), k


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
    worst = 0.0
    for k, p in m.named_parameters():
        err = (p.grad - ref_grads[k]).norm().item()
        ref_n = ref_grads[k].norm().item()
        worst = max(worst, err / (ref_n + 1e-12))
        assert err <= GRAD_REL * ref_n + 1e-7, f"grad {k}: rel {err / (ref_n + 1e-12):.3e}"
    print(f"[parity] config-2 shape worst grad rel err {worst:.3e}")
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
        safe = (rf - 0.5).abs() > MARGIN
        assert torch.equal((got > 0.5)[safe], (rf > 0.5)[safe])
