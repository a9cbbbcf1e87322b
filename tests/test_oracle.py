"""Pins the oracle (oracle/unet_oracle.py) against the reference:
  * golden vectors recorded from the unmodified reference (tests/golden/unet_golden.pt, produced
    by tests/golden/make_golden.py) — runs anywhere, CPU only;
  * the live reference in /root/reference, when present (build container only).
fp32 vs fp32 on the same torch build: tolerances are a few ulps of accumulated rounding."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import unet_oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "unet_golden.pt")


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLDEN, weights_only=False)


def _inputs(seed, n, h, w):
    g = torch.Generator().manual_seed(seed)
    image = torch.randn(n, 3, h, w, generator=g)
    mask = (torch.rand(n, 1, h, w, generator=g) > 0.95).float()
    return image, mask


def _summ(t, k=6):
    t = t.detach().float().flatten()
    idx = torch.linspace(0, t.numel() - 1, k).long()
    return float(t.norm()), float(t.sum()), t[idx]


def test_known_answer_facts(golden):
    f = golden["facts"]
    assert f["unet_params"] == 31037633 and f["anomaly_params"] == 43228228
    for arch, keys in (("unet", f["unet_keys"]), ("anomaly_unet", f["anomaly_keys"])):
        spec = O.state_dict_spec(arch)
        assert list(spec.keys()) == keys
        n_params = sum(torch.Size(s).numel() for k, (s, _) in spec.items() if not O.is_buffer(k))
        assert n_params == f[arch.replace("anomaly_unet", "anomaly") + "_params"]
    assert len(f["unet_keys"]) == 118 and len(f["anomaly_keys"]) == 176


@pytest.mark.parametrize("case_name", ["anomaly_unet_2x32x32", "anomaly_unet_1x48x64", "unet_2x32x32", "unet_1x48x64"])
def test_oracle_matches_golden(golden, case_name):
    torch.set_num_threads(8)
    case = golden["cases"][case_name]
    arch = case["arch"]
    n, h, w = case["shape"]
    sd = O.synthetic_state_dict(arch, seed=case["weight_seed"])
    image, mask = _inputs(case["input_seed"], n, h, w)
    fwd = O.anomaly_unet_forward if arch == "anomaly_unet" else O.unet_forward
    with torch.no_grad():
        o = fwd(sd, image, training=False)
    o = o if isinstance(o, tuple) else (o,)
    for got, ref in zip(o, case["eval_out"]):
        torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-5)
    opt_state = {}
    losses, grads = O.train_step(sd, image, mask, arch=arch, opt_state=opt_state, step=1)
    for k, v in case["losses"].items():
        assert abs(losses[k] - v) <= 1e-5 * max(1.0, abs(v)), (k, losses[k], v)
    for k, ref in case["grads"].items():
        norm, s, samples = _summ(grads[k])
        assert abs(norm - ref["norm"]) <= 2e-3 * ref["norm"] + 1e-7, (k, norm, ref["norm"])
        torch.testing.assert_close(samples, ref["samples"], rtol=5e-3, atol=1e-5 * max(ref["norm"], 1e-3))
    for k, ref in case["buffers_after"].items():
        norm, s, samples = _summ(sd[k])
        assert abs(norm - ref["norm"]) <= 1e-4 * ref["norm"] + 1e-6, k
    for k, ref in case["params_after_adam"].items():
        norm, s, samples = _summ(sd[k])
        assert abs(norm - ref["norm"]) <= 1e-5 * ref["norm"] + 1e-6, k
        torch.testing.assert_close(samples, ref["samples"], rtol=1e-4, atol=2e-4)  # Adam's first step is +-lr


def test_losses_match_golden(golden):
    s = golden["ssim"]
    g = torch.Generator().manual_seed(s["seed"])
    a = torch.rand(*s["shape"], generator=g)
    b = (a + 0.1 * torch.randn(*s["shape"], generator=g)).clamp(0, 1)
    a.requires_grad_(True)
    val = O.ssim_loss(a, b)
    val.backward()
    assert abs(float(val) - s["value"]) < 1e-6
    torch.testing.assert_close(a.grad, s["grad"], rtol=1e-4, atol=1e-8)
    torch.testing.assert_close(O.ssim_loss(a.detach(), b, size_average=False), s["per_image"], rtol=1e-5, atol=1e-6)
    assert float(O.ssim_loss(b, b)) == s["self"] == 0.0
    c = golden["combined_loss"]
    g = torch.Generator().manual_seed(c["seed"])
    recon = torch.rand(2, 3, 24, 24, generator=g).requires_grad_(True)
    amap = torch.rand(2, 1, 24, 24, generator=g).clamp(1e-4, 1 - 1e-4).requires_grad_(True)
    img = torch.randn(2, 3, 24, 24, generator=g)
    msk = (torch.rand(2, 1, 24, 24, generator=g) > 0.9).float()
    ld = O.combined_loss(recon, amap, img, msk, *c["weights"])
    ld["total_loss"].backward()
    for k, v in c["values"].items():
        assert abs(float(ld[k]) - v) < 1e-6
    torch.testing.assert_close(recon.grad, c["d_recon"], rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(amap.grad, c["d_amap"], rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(O.anomaly_score(recon.detach(), img), golden["anomaly_score"]["mse"])
    torch.testing.assert_close(O.anomaly_score(recon.detach(), img, "l1"), golden["anomaly_score"]["l1"])


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="live reference only exists in the build container")
def test_oracle_matches_live_reference():
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden import import_reference
    ref = import_reference()
    torch.set_num_threads(8)
    sd = O.synthetic_state_dict("anomaly_unet", seed=3)
    model = ref["model"].AnomalyUNet(3)
    model.load_state_dict(sd)
    image, mask = _inputs(9, 2, 32, 48)
    model.train()
    r, a = model(image)
    ld = ref["train_utils"].CombinedLoss()(r, a, image, mask)
    ld["total_loss"].backward()
    losses, grads = O.train_step(sd, image, mask, arch="anomaly_unet")
    assert abs(losses["total_loss"] - float(ld["total_loss"])) < 1e-6
    for k, p in model.named_parameters():
        torch.testing.assert_close(grads[k], p.grad, rtol=2e-3, atol=1e-6 * float(p.grad.abs().max() + 1))
    for k, v in model.state_dict().items():
        if O.is_buffer(k):
            torch.testing.assert_close(sd[k], v, rtol=1e-5, atol=1e-6)


def test_oracle_matches_reference_on_odd_sizes():
    """Sizes that are not multiples of 16 (F.pad path, src/model.py:57-61): golden vectors of the reference."""
    gold = torch.load(os.path.join(os.path.dirname(GOLDEN), "odd_golden.pt"), weights_only=False)
    for case in gold["cases"]:
        sd = O.synthetic_state_dict(case["arch"], seed=case["weights_seed"])
        g = torch.Generator().manual_seed(case["input_seed"])
        x = torch.randn(*case["shape"], generator=g)
        with torch.no_grad():
            y = O.unet_forward(sd, x) if case["arch"] == "unet" else O.anomaly_unet_forward(sd, x)
        outs = [y] if case["arch"] == "unet" else list(y)
        for got, want in zip(outs, case["outputs"]):
            torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)


def test_oracle_matches_reference_bilinear():
    """bilinear=True decoders (src/model.py:47-49): golden vectors of the reference, incl. the known answer
    of SURVEY.md section 4 (UNet(3, 1, bilinear=True) has 17,262,977 parameters)."""
    gold = torch.load(os.path.join(os.path.dirname(GOLDEN), "bilinear_golden.pt"), weights_only=False)
    assert gold["cases"][0]["n_params"] == 17_262_977
    for case in gold["cases"]:
        sd = O.synthetic_state_dict(case["arch"], seed=case["weights_seed"], bilinear=True)
        assert list(sd.keys()) == case["keys"]
        assert sum(v.numel() for k, v in sd.items() if not O.is_buffer(k)) == case["n_params"]
        g = torch.Generator().manual_seed(case["input_seed"])
        x = torch.randn(*case["shape"], generator=g)
        with torch.no_grad():
            y = O.unet_forward(sd, x) if case["arch"] == "unet" else O.anomaly_unet_forward(sd, x)
        outs = [y] if case["arch"] == "unet" else list(y)
        for got, want in zip(outs, case["outputs"]):
            torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-5)


def test_bilinear_modules_have_the_reference_layout():
    from tiaozhanbei_unet_b200.model import AnomalyUNet, UNet
    for arch, m in (("unet", UNet(3, 1, bilinear=True)), ("anomaly_unet", AnomalyUNet(3, bilinear=True))):
        spec = O.state_dict_spec(arch, 3, 1, bilinear=True)
        sd = m.state_dict()
        assert list(sd.keys()) == list(spec.keys())
        assert all(tuple(sd[k].shape) == tuple(spec[k][0]) for k in spec)
