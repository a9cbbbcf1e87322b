"""SegmentationUNet row on the GPU: the loss / argmax / dropout kernels and the
model through the C ABI, against golden vectors of the unmodified reference and the fp32 oracle."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import unet_oracle as O  # noqa: E402
from test_seg_oracle import GOLDEN, loss_inputs  # noqa: E402


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLDEN, weights_only=False)


def test_segmentation_losses_match_reference_golden(cuda_device, golden):
    from tiaozhanbei_unet_b200 import metrics
    dev = cuda_device
    for case in golden["losses"]:
        logits, target = loss_inputs(case["seed"], ignore=case["ignore_in_target"])
        logits = logits.to(dev).requires_grad_(True)
        crit = metrics.CombinedSegmentationLoss(**case["kwargs"])
        val = crit(logits, target.to(dev))
        val.backward()
        assert abs(val.item() - case["value"]) <= 2e-5 * max(1.0, abs(case["value"])), (case["kwargs"], val.item(), case["value"])
        torch.testing.assert_close(logits.grad.cpu(), case["grad"], rtol=2e-4, atol=2e-8)
        # the three terms are reported too and add up
        ce, dice, focal = crit.last_terms()
        kw = case["kwargs"]
        total = sum(wt * v for wt, v in ((kw["ce_weight"], ce), (kw["dice_weight"], dice), (kw["focal_weight"], focal)) if wt > 0)
        assert abs(total - val.item()) <= 1e-5 * max(1.0, abs(total))
    d = golden["dice"]
    logits, target = loss_inputs(d["seed"])
    probs = torch.softmax(logits, dim=1).to(dev).requires_grad_(True)
    val = metrics.dice_loss(probs, target.to(dev))
    val.backward()
    assert abs(val.item() - d["value"]) <= 2e-5
    torch.testing.assert_close(probs.grad.cpu(), d["grad"], rtol=2e-4, atol=1e-9)
    f = golden["focal"]
    logits, target = loss_inputs(f["seed"], ignore=f["ignore_index"])
    logits = logits.to(dev).requires_grad_(True)
    val = metrics.focal_loss(logits, target.to(dev), alpha=f["alpha"], gamma=f["gamma"], ignore_index=f["ignore_index"])
    val.backward()
    assert abs(val.item() - f["value"]) <= 2e-5
    torch.testing.assert_close(logits.grad.cpu(), f["grad"], rtol=2e-4, atol=1e-9)
    # upstream gradient scaling and a larger, ragged shape against the oracle's autograd
    g = torch.Generator().manual_seed(5)
    z = (torch.randn(3, 3, 37, 53, generator=g) * 3).to(dev)
    t = torch.randint(0, 3, (3, 37, 53), generator=g).to(dev)
    z1 = z.clone().requires_grad_(True)
    (metrics.CombinedSegmentationLoss(1.0, 2.0, 0.5, ignore_index=1, class_weights=[1.0, 0.0, 2.0])(z1, t) * 0.37).backward()
    z2 = z.clone().requires_grad_(True)
    ref, _ = O.combined_segmentation_loss(z2, t, 1.0, 2.0, 0.5, ignore_index=1, class_weights=[1.0, 0.0, 2.0])
    (ref * 0.37).backward()
    torch.testing.assert_close(z1.grad, z2.grad, rtol=5e-4, atol=1e-9)


def test_argmax_and_channel_scale(cuda_device, golden):
    from tiaozhanbei_unet_b200 import metrics, ops
    dev = cuda_device
    g = torch.Generator().manual_seed(2)
    z = torch.randn(3, 4, 19, 33, generator=g).to(dev)
    z[0, :, 0, 0] = 1.5                                   # ties: the first maximum wins
    z[1, 1:3, 2, 2] = 9.0
    assert torch.equal(metrics.argmax_predictions(z), torch.argmax(z, dim=1))
    m = golden["model"]
    assert torch.equal(metrics.argmax_predictions(m["logits"].to(dev)).cpu(), m["argmax"])
    x = torch.randn(2, 4, 6, 64, generator=g).to(dev).to(torch.bfloat16)
    mask = (torch.rand(2, 64, 1, 1, generator=g) > 0.3).float().div(0.7).to(dev)
    want = (x.float() * mask.view(2, 1, 1, 64)).to(torch.bfloat16)
    ops.channel_scale(x, mask)
    torch.cuda.synchronize()
    assert torch.equal(x, want)


def test_segmentation_unet_eval_matches_reference_golden(cuda_device, golden):
    from tiaozhanbei_unet_b200.model import SegmentationUNet
    from test_model_gpu import OUT_MAX, OUT_MEAN
    dev = cuda_device
    m = golden["model"]
    net = SegmentationUNet(3, 4)
    net.load_state_dict(O.synthetic_state_dict("unet", seed=m["weights_seed"], n_classes=4))
    net = net.to(dev).eval()
    g = torch.Generator().manual_seed(m["input_seed"])
    x = torch.randn(*m["shape"], generator=g).to(dev)
    with torch.no_grad():
        z = net(x)
    ref = m["logits"].to(dev)
    d = (z - ref).abs()
    scale = ref.abs().max().item()
    assert d.max().item() <= OUT_MAX * scale and d.mean().item() <= OUT_MEAN * scale, (d.max().item(), d.mean().item(), scale)
    # class decisions: bit-exact wherever the reference's top-2 margin exceeds the output tolerance
    top2 = ref.topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 2 * OUT_MAX * scale
    from tiaozhanbei_unet_b200 import metrics
    pred = metrics.argmax_predictions(z)
    assert safe.any() and torch.equal(pred[safe].cpu(), m["argmax"].to(dev)[safe].cpu())
    assert (pred.cpu() != m["argmax"]).float().mean().item() <= 0.02


def test_segmentation_unet_train_step_matches_oracle(cuda_device):
    """One training step with Dropout2d active: the same torch generator state gives the same channel
    mask as the reference's nn.Dropout2d, so loss and gradients are comparable with the oracle."""
    from tiaozhanbei_unet_b200 import metrics
    from tiaozhanbei_unet_b200.model import SegmentationUNet
    from test_model_gpu import LOSS_REL, _check_grads
    dev = cuda_device
    sd = O.synthetic_state_dict("unet", seed=13, n_classes=4, device=dev)
    net = SegmentationUNet(3, 4, dropout=0.25)
    net.load_state_dict({k: v.clone() for k, v in sd.items()})
    net = net.to(dev).train()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(3, 3, 64, 48, generator=g).to(dev)
    t = torch.randint(0, 4, (3, 64, 48), generator=g).to(dev)
    crit = metrics.CombinedSegmentationLoss(1.0, 1.0, 0.0, class_weights=[0.5, 1.0, 2.0, 1.0])
    torch.manual_seed(77)
    loss = crit(net(x), t)
    loss.backward()
    mask = net._engine().plans[(3, 64, 48)].drop_mask.clone()
    assert 0 < int((mask == 0).sum()) < mask.numel()          # some (image, channel) pairs were dropped
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if not O.is_buffer(k)}
    work = {k: (leaves[k] if k in leaves else v.clone()) for k, v in sd.items()}
    torch.manual_seed(77)
    ref, _ = O.combined_segmentation_loss(O.segmentation_unet_forward(work, x, training=True, dropout=0.25), t, 1.0, 1.0, 0.0,
                                          class_weights=[0.5, 1.0, 2.0, 1.0])
    ref.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - ref.item()) <= LOSS_REL * abs(ref.item()), (loss.item(), ref.item())
    worst = _check_grads(net.named_parameters(), {k: v.grad for k, v in leaves.items()}, "segmentation_unet 3x64x48")
    print(f"[parity] segmentation_unet 3x64x48 (dropout 0.25): loss {loss.item():.6f} vs {ref.item():.6f}; "
          f"worst grad cosine {worst[0]:.3f}, worst head-side rel err {worst[1]:.3e}")
    # eval mode: dropout is the identity and repeated calls agree bit for bit
    net.eval()
    with torch.no_grad():
        a, b = net(x), net(x)
    assert torch.equal(a, b)
