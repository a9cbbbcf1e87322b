"""SegmentationUNet row: the oracle's restatement of the reference's segmentation
losses and model against golden vectors recorded from the unmodified reference
(tests/golden/make_golden_seg.py), on CPU."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import unet_oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "seg_golden.pt")


def loss_inputs(seed, n=2, c=4, h=24, w=20, ignore=None):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(n, c, h, w, generator=g) * 2.0
    target = torch.randint(0, c, (n, h, w), generator=g)
    if ignore is not None:
        target[torch.rand(n, h, w, generator=g) < 0.1] = ignore
    return logits, target


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLDEN, weights_only=False)


def test_oracle_segmentation_losses_match_reference(golden):
    for case in golden["losses"]:
        logits, target = loss_inputs(case["seed"], ignore=case["ignore_in_target"])
        logits.requires_grad_(True)
        total, _ = O.combined_segmentation_loss(logits, target, **case["kwargs"])
        total.backward()
        assert abs(float(total) - case["value"]) <= 1e-6 * max(1.0, abs(case["value"])), case["kwargs"]
        torch.testing.assert_close(logits.grad, case["grad"], rtol=1e-5, atol=1e-9)
    d = golden["dice"]
    logits, target = loss_inputs(d["seed"])
    probs = torch.softmax(logits, dim=1).requires_grad_(True)
    val = O.seg_dice_loss(probs, target)
    val.backward()
    assert abs(float(val) - d["value"]) <= 1e-6
    torch.testing.assert_close(probs.grad, d["grad"], rtol=1e-5, atol=1e-10)
    f = golden["focal"]
    logits, target = loss_inputs(f["seed"], ignore=f["ignore_index"])
    logits.requires_grad_(True)
    val = O.seg_focal_loss(logits, target, alpha=f["alpha"], gamma=f["gamma"], ignore_index=f["ignore_index"])
    val.backward()
    assert abs(float(val) - f["value"]) <= 1e-6
    torch.testing.assert_close(logits.grad, f["grad"], rtol=1e-5, atol=1e-10)


def test_oracle_segmentation_unet_matches_reference(golden):
    m = golden["model"]
    assert m["n_params"] == 31_037_828                      # SURVEY.md section 4 known answer
    sd = O.synthetic_state_dict("unet", seed=m["weights_seed"], n_classes=4)
    assert list(sd.keys()) == m["keys"]                      # Dropout2d adds no state: the UNet layout
    g = torch.Generator().manual_seed(m["input_seed"])
    x = torch.randn(*m["shape"], generator=g)
    with torch.no_grad():
        z = O.segmentation_unet_forward(sd, x, training=False)
    torch.testing.assert_close(z, m["logits"], rtol=1e-4, atol=1e-5)
    assert torch.equal(torch.argmax(z, dim=1), m["argmax"])


def test_segmentation_module_surface():
    """Constructor signature, state_dict layout and error behaviour of the drop-in classes (no GPU needed)."""
    from tiaozhanbei_unet_b200 import metrics
    from tiaozhanbei_unet_b200.model import SegmentationUNet
    m = SegmentationUNet(n_channels=3, n_classes=4, bilinear=False, dropout=0.1)
    spec = O.state_dict_spec("unet", 3, 4)
    sd = m.state_dict()
    assert list(sd.keys()) == list(spec.keys())
    assert all(tuple(sd[k].shape) == tuple(spec[k][0]) and sd[k].dtype == spec[k][1] for k in spec)
    assert sum(p.numel() for p in m.parameters()) == 31_037_828
    assert isinstance(m.dropout, torch.nn.Dropout2d) and m.dropout.p == 0.1
    assert isinstance(SegmentationUNet(dropout=0.0).dropout, torch.nn.Identity)
    crit = metrics.CombinedSegmentationLoss(focal_weight=0.5)          # ignore_index=None + focal: reference raises
    with pytest.raises(TypeError):
        crit(torch.zeros(1, 4, 2, 2), torch.zeros(1, 2, 2, dtype=torch.long))
    with pytest.raises(TypeError):
        metrics.focal_loss(torch.zeros(1, 4, 2, 2), torch.zeros(1, 2, 2, dtype=torch.long))
    with pytest.raises(RuntimeError):                                  # no CPU path
        metrics.CombinedSegmentationLoss()(torch.zeros(1, 4, 2, 2), torch.zeros(1, 2, 2, dtype=torch.long))
