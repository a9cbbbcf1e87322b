"""CPU ORACLE (test infrastructure, not product code) for validate_epoch's reductions, SURVEY.md §8(f) row 3.

numpy restatement of what /root/reference/src/train_utils.py:205-246 computes over the collected maps, each
function citing the lines it follows.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import
this module.  Pinned by tests/golden/validate_golden.pt, recorded from the UNMODIFIED reference's
`validate_epoch` (tests/golden/make_golden_validate.py).
"""
import numpy as np


def confusion_counts(true_flat, pred_flat):
    """(tn, fp, fn, tp) = sklearn.metrics.confusion_matrix(y_true, y_pred).ravel() for binary 0/1 arrays that
    contain both classes in y_true (src/utils.py:76): rows are the true class, columns the predicted one."""
    t = np.asarray(true_flat).astype(np.int64).reshape(-1)
    p = np.asarray(pred_flat).astype(np.int64).reshape(-1)
    m = np.bincount(t * 2 + p, minlength=4)
    return int(m[0]), int(m[1]), int(m[2]), int(m[3])


def count_metrics(tn, fp, fn, tp):
    """src/utils.py:78-82."""
    precision = tp / (tp + fp) if (tp + fp) > 0 else 0
    recall = tp / (tp + fn) if (tp + fn) > 0 else 0
    return {
        "accuracy": (tp + tn) / (tp + tn + fp + fn),
        "precision": precision,
        "recall": recall,
        "specificity": tn / (tn + fp) if (tn + fp) > 0 else 0,
        "f1_score": 2 * precision * recall / (precision + recall) if (precision + recall) > 0 else 0,
    }


def pixel_confusion(masks_pred, masks_true, labels, thresholds=(0.3, 0.5, 0.7)):
    """Per-threshold (tn, fp, fn, tp) over the images with label 1 (src/train_utils.py:232-241): prediction
    `> threshold` (the python float compared in the maps' float32), truth `> 0.5`."""
    masks_pred = np.asarray(masks_pred, dtype=np.float32)
    masks_true = np.asarray(masks_true, dtype=np.float32)
    keep = np.asarray(labels) == 1
    out = []
    for threshold in thresholds:
        pred = (masks_pred[keep] > threshold).astype(np.uint8).flatten()
        true = (masks_true[keep] > 0.5).astype(np.uint8).flatten()
        out.append(confusion_counts(true, pred))
    return np.array(out, dtype=np.int64).reshape(len(thresholds), 4)


def pixel_metrics(masks_pred, masks_true, labels, thresholds=(0.3, 0.5, 0.7)):
    """The `pixel_metrics` dict of validate_epoch (src/train_utils.py:230-246)."""
    labels = np.asarray(labels)
    out = {}
    if np.sum(labels == 1) > 0:
        counts = pixel_confusion(masks_pred, masks_true, labels, thresholds)
        for threshold, (tn, fp, fn, tp) in zip(thresholds, counts):
            if tp + fn > 0 and tn + fp > 0:     # len(np.unique(true_flat)) > 1
                out[f"pixel_f1_@{threshold}"] = count_metrics(tn, fp, fn, tp)["f1_score"]
    return out


def score_threshold(all_scores, q=95):
    """src/train_utils.py:213: np.percentile over every pixel of every score map (float32 in, float32 out)."""
    return np.percentile(np.asarray(all_scores, dtype=np.float32), q)


def kth_pair(values, k):
    """The two order statistics the linear percentile interpolates: (k-th, (k+1)-th) smallest, NaN last."""
    s = np.sort(np.asarray(values, dtype=np.float32).reshape(-1))
    return s[k], s[min(k + 1, s.size - 1)]


def class_confusion(pred, target, num_classes, ignore_index=None):
    """SegmentationMetrics.update (src/metrics.py:22-45) for numpy inputs: (counts int64 [C, C], total_samples).
    pred (N, C, H, W) scores -> argmax over axis 1 (first maximum) or (N, H, W) indices; pixels with the ignored
    label are removed, then sklearn's confusion_matrix(target, pred, labels=range(C)) counts the pairs whose label
    and prediction are both one of the C classes."""
    pred = np.asarray(pred)
    if pred.ndim == 4:
        pred = np.argmax(pred, axis=1)
    pred = pred.reshape(-1).astype(np.int64)
    target = np.asarray(target).reshape(-1).astype(np.int64)
    if ignore_index is not None:
        keep = target != ignore_index
        pred, target = pred[keep], target[keep]
    ok = (target >= 0) & (target < num_classes) & (pred >= 0) & (pred < num_classes)
    cm = np.bincount(target[ok] * num_classes + pred[ok], minlength=num_classes * num_classes)
    return cm.reshape(num_classes, num_classes).astype(np.int64), int(target.size)
