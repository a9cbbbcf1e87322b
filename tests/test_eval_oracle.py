"""validate_epoch's reductions on the CPU: the numpy oracle against the golden vectors
recorded from the unmodified reference's validate_epoch, and the host half of the device percentile (index /
interpolation arithmetic) against numpy itself."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from oracle import eval_oracle as E  # noqa: E402
from make_golden_validate import CASES, batches  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "validate_golden.pt")


def collected(data):
    """The four arrays validate_epoch collects (src/train_utils.py:192-203) from a replayed validation set."""
    labels = np.concatenate([b["label"].numpy() for b in data])
    scores = np.concatenate([((b["recon"] - b["image"]) ** 2).mean(1).numpy() for b in data])
    true = np.concatenate([b["mask"].numpy() for b in data])
    pred = np.concatenate([b["amap"].numpy() for b in data])
    return labels, scores, true, pred


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLDEN, weights_only=False)


def test_oracle_matches_reference_validate_epoch(golden):
    for name, labels in CASES.items():
        case = golden["cases"][name]
        assert case["labels"] == labels
        lab, scores, true, pred = collected(batches(case["seed"], labels))
        assert float(E.score_threshold(scores)) == case["percentile95"]
        if "error" in case:
            continue
        assert E.pixel_metrics(pred, true, lab) == case["pixel_metrics"]       # exact: integer counts behind them
        assert scores.shape == case["shapes"]["scores"] and pred.shape == case["shapes"]["masks_pred"]
        assert abs(scores.astype(np.float64).sum() - case["scores_sum"]) < 1e-4
    assert golden["cases"]["mixed"]["error_type"] == "ValueError"
    assert golden["cases"]["anomalous"]["pixel_metrics"] and not golden["cases"]["normal"]["pixel_metrics"]


def test_confusion_counts_are_sklearns():
    from sklearn.metrics import confusion_matrix
    rng = np.random.default_rng(3)
    for n in (2, 17, 4096):
        t = rng.integers(0, 2, n)
        t[:2] = (0, 1)
        p = rng.integers(0, 2, n)
        assert E.confusion_counts(t, p) == tuple(int(v) for v in confusion_matrix(t, p).ravel())
    from tiaozhanbei_unet_b200.utils import calculate_metrics, metrics_from_confusion
    got = metrics_from_confusion(*E.confusion_counts(t, p))
    want = calculate_metrics(t, p)
    assert got == {k: want[k] for k in got} == E.count_metrics(*E.confusion_counts(t, p))


@pytest.mark.parametrize("n", [1, 2, 3, 7, 100, 1001, 65536, 7 * 24 * 20, (1 << 24) + 12345])
def test_percentile_index_and_interpolation_are_numpys(n):
    """The product's host half: given the exact order statistics, the result is np.percentile's bit for bit
    (float32 index arithmetic included: at 2^24+ elements the virtual index is rounded in float32)."""
    from tiaozhanbei_unet_b200.utils import percentile_lerp, percentile_rank
    rng = np.random.default_rng(n)
    a = (rng.random(n, dtype=np.float32) ** 2) * 0.05
    srt = np.sort(a)
    for q in (0, 5, 37.5, 50, 95, 99.9, 100):
        lo, hi, gamma = percentile_rank(n, q)
        assert (lo, hi) == (min(lo, n - 1), min(hi, n - 1))
        got = percentile_lerp(srt[lo], srt[hi], gamma)
        want = np.percentile(a, q)
        assert got == want and got.dtype == want.dtype == np.float32, (n, q, got, want)
        assert E.kth_pair(a, lo) == (srt[lo], srt[min(lo + 1, n - 1)])
    with pytest.raises(ValueError):
        percentile_rank(n, 101)


def test_lazy_predictions_mapping():
    from tiaozhanbei_unet_b200.train_utils import _Predictions
    labels = np.array([0, 1, 1])
    scores = [torch.arange(8.).view(2, 2, 2), torch.ones(1, 2, 2)]
    true = [torch.zeros(2, 1, 2, 2), torch.ones(1, 1, 2, 2)]
    pred = [torch.rand(2, 1, 2, 2), torch.rand(1, 1, 2, 2)]
    p = _Predictions(labels, scores, true, pred)
    assert list(p) == ["labels", "scores", "masks_true", "masks_pred"] and len(p) == 4 and "scores" in p
    assert p["labels"] is labels
    assert p["scores"].shape == (3, 2, 2) and p["scores"].dtype == np.float32
    assert p["scores"] is p["scores"]                      # copied once
    assert dict(p)["masks_pred"].shape == (3, 1, 2, 2)
    np.testing.assert_array_equal(p["masks_true"], torch.cat(true).numpy())
    with pytest.raises(KeyError):
        p["nope"]


# ---------------------------------------------------------------------------------------------
# SegmentationMetrics (src/metrics.py:9-133)
# ---------------------------------------------------------------------------------------------
REF_METRICS = "/root/reference/src/metrics.py"


def seg_case(seed, n=3, c=4, h=20, w=24, ignore=None, stray=False):
    """Seeded logits / labels: a dominant background, exact ties between channels, optional ignored / stray labels."""
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(n, c, h, w, generator=g)
    logits[:, 0] += 1.0
    if c > 1:
        logits[:, 1, ::3] = logits[:, 0, ::3]              # ties: the first maximum wins
    target = torch.randint(0, c, (n, h, w), generator=g)
    target[torch.rand(n, h, w, generator=g) < 0.5] = 0
    if ignore is not None:
        target[torch.rand(n, h, w, generator=g) < 0.1] = ignore
    if stray:
        target[torch.rand(n, h, w, generator=g) < 0.05] = 255      # a label outside range(C) that is NOT ignored
    return logits, target


def _reference_metrics_module():
    import importlib.util
    import types
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    spec = importlib.util.spec_from_file_location("_ref_metrics", REF_METRICS)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_class_confusion_oracle_is_sklearns():
    from sklearn.metrics import confusion_matrix
    for seed, kw in enumerate([dict(), dict(ignore=255), dict(stray=True), dict(c=1), dict(c=8, ignore=7)]):
        logits, target = seg_case(seed, **kw)
        c = logits.shape[1]
        ignore = kw.get("ignore")
        cm, total = E.class_confusion(logits.numpy(), target.numpy(), c, ignore)
        pred = torch.argmax(logits, dim=1).numpy().reshape(-1)
        t = target.numpy().reshape(-1)
        if ignore is not None:
            pred, t = pred[t != ignore], t[t != ignore]
        np.testing.assert_array_equal(cm, confusion_matrix(t, pred, labels=range(c)))
        assert total == t.size
        cm2, _ = E.class_confusion(torch.argmax(logits, dim=1).numpy(), target.numpy(), c, ignore)
        np.testing.assert_array_equal(cm, cm2)


@pytest.mark.skipif(not os.path.exists(REF_METRICS), reason="reference sources only exist in the build container")
def test_segmentation_metrics_match_the_live_reference():
    """Oracle counts == the reference's SegmentationMetrics.update; this package's compute_* formulas == the
    reference's on those counts (every key of compute_all_metrics, exactly)."""
    from tiaozhanbei_unet_b200 import metrics as M
    ref = _reference_metrics_module()
    for seed, kw in enumerate([dict(), dict(ignore=255), dict(stray=True), dict(c=8, ignore=7)]):
        logits, target = seg_case(10 + seed, **kw)
        c = logits.shape[1]
        r = ref.SegmentationMetrics(c, ignore_index=kw.get("ignore"))
        r.update(logits[:2], target[:2])
        r.update(torch.argmax(logits[2:], dim=1), target[2:])
        cm, total = E.class_confusion(logits.numpy(), target.numpy(), c, kw.get("ignore"))
        np.testing.assert_array_equal(cm, r.confusion_matrix)
        assert total == r.total_samples
        ours = M.SegmentationMetrics(c, ignore_index=kw.get("ignore"))
        ours.confusion_matrix = cm
        ours.total_samples = total
        want, got = r.compute_all_metrics(), ours.compute_all_metrics()
        assert set(want) == set(got)
        for k in want:
            np.testing.assert_array_equal(np.asarray(got[k]), np.asarray(want[k]), err_msg=k)
        assert ours.compute_iou(per_class=False) == r.compute_iou(per_class=False)
        assert ours.compute_dice(per_class=False) == r.compute_dice(per_class=False)
        assert ours.compute_precision_recall_f1(per_class=False) == r.compute_precision_recall_f1(per_class=False)
    empty = M.SegmentationMetrics(3)
    assert empty.compute_all_metrics()["mean_iou"] == ref.SegmentationMetrics(3).compute_all_metrics()["mean_iou"] == 0.0
    with pytest.raises(NotImplementedError):
        M.SegmentationMetrics(9)
    with pytest.raises(RuntimeError):
        M.SegmentationMetrics(4).update(torch.zeros(1, 4, 2, 2), torch.zeros(1, 2, 2, dtype=torch.long))
