"""validate_epoch's reductions on the device through the C ABI: confusion counts and
order statistics bit-exact against the numpy oracle, validate_epoch against the golden vectors of the
unmodified reference."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import eval_oracle as E  # noqa: E402
from test_eval_oracle import CASES, GOLDEN, batches, collected  # noqa: E402


def _maps(seed, n, shape):
    g = torch.Generator().manual_seed(seed)
    true = (torch.rand(n, *shape, generator=g) > 0.8).float()
    pred = (0.35 * true + 0.65 * torch.rand(n, *shape, generator=g))
    pred.view(-1)[::7] = 0.5                                  # values ON a threshold are not above it
    return pred, true


@pytest.mark.parametrize("n,shape", [(7, (1, 24, 20)), (3, (1, 5, 7)), (1, (1, 1, 1)), (16, (1, 256, 256)), (2, (3, 33))])
def test_pixel_confusion_is_bit_exact(cuda_device, n, shape):
    from tiaozhanbei_unet_b200.utils import pixel_confusion_counts
    dev = cuda_device
    pred, true = _maps(n * 10 + len(shape), n, shape)
    thresholds = (0.3, 0.5, 0.7)
    labels = np.ones(n, dtype=np.int64)
    got = pixel_confusion_counts(pred.to(dev), true.to(dev), thresholds)
    np.testing.assert_array_equal(got.cpu().numpy(), E.pixel_confusion(pred.numpy(), true.numpy(), labels, thresholds))
    assert got.dtype == torch.int64 and int(got[0].sum()) == pred.numel()
    # anomalous images only, accumulated over two calls (validate_epoch's batches), 8 thresholds
    labels = (np.arange(n) % 3 != 1).astype(np.int64)
    thr8 = (0.05, 0.2, 0.3, 0.5, 0.5000001, 0.7, 0.95, 2.0)
    cut = max(1, n // 2)
    counts = pixel_confusion_counts(pred[:cut].to(dev), true[:cut].to(dev), thr8, include=labels[:cut] == 1)
    if cut < n:
        counts = pixel_confusion_counts(pred[cut:].to(dev), true[cut:].to(dev), thr8, include=labels[cut:] == 1, counts=counts)
    np.testing.assert_array_equal(counts.cpu().numpy(), E.pixel_confusion(pred.numpy(), true.numpy(), labels, thr8))
    # a view that is not 16-byte aligned takes the scalar path
    if pred[0].numel() > 4:
        flat_p, flat_t = pred.to(dev).view(n, -1)[:, 1:].contiguous(), true.to(dev).view(n, -1)[:, 1:].contiguous()
        off_p = torch.empty(flat_p.numel() + 1, device=dev)[1:].view_as(flat_p).copy_(flat_p)
        off_t = torch.empty(flat_t.numel() + 1, device=dev)[1:].view_as(flat_t).copy_(flat_t)
        got = pixel_confusion_counts(off_p, off_t, thresholds)
        np.testing.assert_array_equal(got.cpu().numpy(), E.pixel_confusion(flat_p.cpu().numpy(), flat_t.cpu().numpy(),
                                                                           np.ones(n, dtype=np.int64), thresholds))


def test_pixel_confusion_argument_errors(cuda_device):
    from tiaozhanbei_unet_b200 import _lib, ops
    dev = cuda_device
    p, t = torch.rand(2, 8, device=dev), torch.rand(2, 8, device=dev)
    counts = torch.zeros(9, 4, dtype=torch.int64, device=dev)
    with pytest.raises(_lib.UnetkError):
        ops.pixel_confusion(p, t, None, [0.1] * 9, counts)                    # more than 8 thresholds
    with pytest.raises(_lib.UnetkError):
        ops.pixel_confusion(p.cpu(), t, None, [0.5], counts[:1])              # no CPU path
    with pytest.raises(ValueError):
        ops.pixel_confusion(p, t[:1], None, [0.5], counts[:1])


def _select_cases():
    rng = np.random.default_rng(5)
    yield "one", np.array([3.5], dtype=np.float32)
    yield "two", np.array([2.0, -1.0], dtype=np.float32)
    yield "ties", rng.integers(-3, 4, 1000).astype(np.float32)
    yield "signed", rng.standard_normal(65537).astype(np.float32)
    yield "zeros", np.array([0.0, -0.0, 1e-45, -1e-45, 0.0, 5.0, -5.0], dtype=np.float32)
    yield "squares", (rng.random(7 * 24 * 20, dtype=np.float32) ** 2 * 0.05)
    yield "constant", np.full(4099, 0.25, dtype=np.float32)
    yield "inf", np.array([np.inf, -np.inf, 1.0, 2.0, np.inf], dtype=np.float32)
    yield "nan_last", np.array([1.0, np.nan, -2.0, 7.0], dtype=np.float32)


@pytest.mark.parametrize("name,values", list(_select_cases()), ids=[c[0] for c in _select_cases()])
def test_select_kth_is_exact(cuda_device, name, values):
    from tiaozhanbei_unet_b200 import ops
    x = torch.from_numpy(values).to(cuda_device)
    n = values.size
    for k in sorted({0, n // 3, n // 2, max(0, n - 2), n - 1}):
        got = ops.select_kth(x, k).cpu().numpy()
        want = np.array(E.kth_pair(values, k), dtype=np.float32)
        # compare values (0.0 == -0.0, NaN == NaN)
        np.testing.assert_array_equal(got, want, err_msg=f"{name} k={k}")
    assert torch.equal(x.cpu(), torch.from_numpy(values)) or name == "nan_last"      # input untouched
    # unaligned view: scalar path
    if n > 8:
        y = torch.empty(n + 1, device=cuda_device)[1:].copy_(x)
        np.testing.assert_array_equal(ops.select_kth(y, n // 2).cpu().numpy(), np.array(E.kth_pair(values, n // 2), dtype=np.float32))


def test_select_kth_argument_errors(cuda_device):
    from tiaozhanbei_unet_b200 import _lib, ops
    x = torch.rand(10, device=cuda_device)
    with pytest.raises(_lib.UnetkError):
        ops.select_kth(x, 10)
    with pytest.raises(_lib.UnetkError):
        ops.select_kth(x, 0, work=torch.empty(4, dtype=torch.int64, device=cuda_device))      # workspace too small


@pytest.mark.parametrize("n", [1, 2, 1001, 7 * 24 * 20, 64 * 512 * 512 + 3])
def test_percentile_is_numpys(cuda_device, n):
    """np.percentile(all_scores, 95) of src/train_utils.py:213 bit for bit, at the full evaluation size too
    (64 score maps of 512 x 512: the float32 virtual index is already inexact there)."""
    from tiaozhanbei_unet_b200.utils import percentile
    g = torch.Generator().manual_seed(n)
    a = torch.rand(n, generator=g) ** 2 * 0.05
    x = a.to(cuda_device)
    for q in ((0, 50, 95, 100) if n > 1 << 20 else (0, 5, 37.5, 50, 95, 99.9, 100)):
        got = percentile(x, q)
        want = np.percentile(a.numpy(), q)
        assert got == want and got.dtype == np.float32, (n, q, got, want)


class _Replay(torch.nn.Module):
    def __init__(self, data, dev):
        super().__init__()
        self.data, self.i, self.dev = data, 0, dev

    def forward(self, images):
        b = self.data[self.i]
        self.i += 1
        return b["recon"].to(self.dev), b["amap"].to(self.dev)


def test_validate_epoch_matches_reference_golden(cuda_device):
    from tiaozhanbei_unet_b200 import train_utils as T
    golden = torch.load(GOLDEN, weights_only=False)
    for name, labels in CASES.items():
        case = golden["cases"][name]
        data = batches(case["seed"], labels)
        if "error" in case:
            with pytest.raises(ValueError) as e:
                T.validate_epoch(_Replay(data, cuda_device), data, T.CombinedLoss(), cuda_device)
            assert str(e.value) == case["error"]
            continue
        res = T.validate_epoch(_Replay(data, cuda_device), data, T.CombinedLoss(), cuda_device)
        assert res["pixel_metrics"] == case["pixel_metrics"]                  # exact
        assert {k: float(v) for k, v in res["image_metrics"].items()} == case["image_metrics"]
        for key in ("total_loss", "recon_loss", "seg_loss"):
            assert abs(res[key] - case[key]) <= 2e-5 * max(1.0, abs(case[key])), (name, key, res[key], case[key])
        pred = res["predictions"]
        assert {k: tuple(v.shape) for k, v in pred.items()} == case["shapes"]
        assert {k: str(v.dtype) for k, v in pred.items()} == case["dtypes"]
        lab, scores, true, amap = collected(data)
        np.testing.assert_array_equal(pred["masks_pred"], amap)
        np.testing.assert_array_equal(pred["masks_true"], true)
        np.testing.assert_allclose(pred["scores"], scores, rtol=1e-5, atol=1e-9)
        assert abs(float(T.percentile(torch.from_numpy(scores).to(cuda_device), 95)) - case["percentile95"]) == 0.0


def test_confusion_and_select_bandwidth(cuda_device):
    """Not a pass/fail on speed: a sanity bound that both reductions stream (well above what a serialised reduction reaches, at the evaluation size:
    64 maps of 512 x 512) instead of serialising on atomics; the measured numbers are printed for the log."""
    from tiaozhanbei_unet_b200 import ops
    from tiaozhanbei_unet_b200.utils import pixel_confusion_counts
    dev = cuda_device
    pred = torch.rand(64, 1, 512, 512, device=dev)
    true = (torch.rand(64, 1, 512, 512, device=dev) > 0.9).float()
    scores = torch.rand(64 * 512 * 512, device=dev) ** 2

    def timed(fn, reps=10):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps * 1e-3

    counts = torch.zeros(3, 4, dtype=torch.int64, device=dev)
    t_conf = timed(lambda: ops.pixel_confusion(pred, true, None, [0.3, 0.5, 0.7], counts))
    out2 = torch.empty(2, device=dev)
    work = torch.empty(2048, dtype=torch.int64, device=dev)
    k = int(0.95 * (scores.numel() - 1))
    t_sel = timed(lambda: ops.select_kth(scores, k, out2, work))
    gb_conf = pred.numel() * 8 / 1e9
    gb_sel = scores.numel() * 4 * 5 / 1e9
    print(f"\npixel_confusion {t_conf * 1e3:.3f} ms {gb_conf / t_conf:.0f} GB/s; select_kth {t_sel * 1e3:.3f} ms "
          f"{gb_sel / t_sel:.0f} GB/s (5 passes)")
    assert gb_conf / t_conf > 500 and gb_sel / t_sel > 100
    want = torch.sort(scores).values[k:k + 2]
    assert torch.equal(out2, want)


SEG_CASES = [dict(), dict(ignore=255), dict(stray=True), dict(c=1), dict(c=8, ignore=7), dict(c=3, h=33, w=47),
             dict(n=4, c=4, h=512, w=512, ignore=255)]


@pytest.mark.parametrize("kw", SEG_CASES, ids=[str(i) for i in range(len(SEG_CASES))])
def test_segmentation_metrics_update_is_bit_exact(cuda_device, kw):
    """SegmentationMetrics.update (src/metrics.py:22-45) on the device: logits and index input, accumulation over
    updates, ignored and out-of-range labels, ties and NaN in the argmax, reads between updates."""
    from test_eval_oracle import seg_case
    from tiaozhanbei_unet_b200 import metrics as M
    dev = cuda_device
    logits, target = seg_case(20 + len(kw), **kw)
    logits[0, 0, 0, :5] = float("nan")                       # torch.argmax: a NaN is the maximum
    c, ignore = logits.shape[1], kw.get("ignore")
    cm, total = E.class_confusion(logits.numpy(), target.numpy(), c, ignore)
    np.testing.assert_array_equal(torch.argmax(logits, dim=1).numpy(), np.argmax(logits.numpy(), axis=1))
    m = M.SegmentationMetrics(c, ignore_index=ignore)
    m.update(logits[:2].to(dev), target[:2].to(dev))                               # (N, C, H, W) scores
    m.update(torch.argmax(logits[2:], dim=1).to(dev), target[2:].to(dev))          # (N, H, W) indices
    np.testing.assert_array_equal(m.confusion_matrix, cm)
    assert m.total_samples == total and m.confusion_matrix.dtype == np.int64
    m.update(logits.to(dev), target.to(dev))                                       # counting goes on after a read
    m.update(logits[:0].to(dev), target[:0].to(dev))                               # an empty batch adds nothing
    np.testing.assert_array_equal(m.confusion_matrix, 2 * cm)
    assert m.total_samples == 2 * total
    got = m.compute_all_metrics()
    tp, rows, cols = np.diag(2 * cm), (2 * cm).sum(1), (2 * cm).sum(0)
    np.testing.assert_array_equal(got["iou_per_class"], tp / np.maximum(rows + cols - tp, 1e-8))
    assert got["pixel_accuracy"] == tp.sum() / max((2 * cm).sum(), 1e-8)
    m.reset()
    assert m.total_samples == 0 and not m.confusion_matrix.any()
    m.update(logits, target.to(dev))                                               # host tensors are moved, not refused
    m.update(torch.argmax(logits, dim=1), target)
    np.testing.assert_array_equal(m.confusion_matrix, 2 * cm)
    full = M.compute_metrics_from_predictions(logits.to(dev), target.to(dev), c)
    np.testing.assert_array_equal(full["confusion_matrix"], E.class_confusion(logits.numpy(), target.numpy(), c)[0])
