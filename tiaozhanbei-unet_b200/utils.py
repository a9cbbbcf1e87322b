"""The parts of the reference's `src/utils.py` that the hot path touches.

`compute_anomaly_score` (src/utils.py:205-215) is a CUDA kernel here, and so are the two reductions
validate_epoch runs over every map (`pixel_confusion_counts`, `percentile`: src/train_utils.py:213, 232-246); the checkpoint helpers keep the
reference's dict layout (src/utils.py:37-58); AverageMeter (src/utils.py:285-300) and
calculate_metrics (src/utils.py:61-94) are host-side bookkeeping restated so that this package does
not need the reference's file (which imports matplotlib / seaborn at top level).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def compute_anomaly_score(reconstruction, original, method="mse"):
    """Per-pixel channel-mean reconstruction error -> [B, H, W] fp32 (src/utils.py:205-215)."""
    if method in ("mse", "ssim"):      # the reference's 'ssim' branch is the MSE score too
        m = 0
    elif method == "l1":
        m = 1
    else:
        raise ValueError(f"Unknown method: {method}")
    r = reconstruction.detach().float().contiguous()
    o = original.detach().float().contiguous()
    if r.shape != o.shape:
        raise ValueError(f"shape mismatch: {tuple(r.shape)} vs {tuple(o.shape)}")
    n, c, h, w = r.shape
    out = torch.empty(n, h, w, dtype=torch.float32, device=r.device)
    return ops.anomaly_score(r, o, out, m)


def save_checkpoint(model, optimizer, epoch, loss, filepath):
    """Same four-key dict as the reference (src/utils.py:37-46)."""
    checkpoint = {
        "epoch": epoch,
        "model_state_dict": model.state_dict(),
        "optimizer_state_dict": optimizer.state_dict(),
        "loss": loss,
    }
    torch.save(checkpoint, filepath)
    print(f"Checkpoint saved to {filepath}")


def load_checkpoint(model, optimizer, filepath, device):
    """src/utils.py:49-58"""
    checkpoint = torch.load(filepath, map_location=device)
    model.load_state_dict(checkpoint["model_state_dict"])
    if optimizer:
        optimizer.load_state_dict(checkpoint["optimizer_state_dict"])
    epoch = checkpoint["epoch"]
    loss = checkpoint["loss"]
    print(f"Checkpoint loaded from {filepath}, epoch {epoch}, loss {loss:.4f}")
    return epoch, loss


class AverageMeter:
    """src/utils.py:285-300"""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def calculate_metrics(y_true, y_pred, y_scores=None):
    """Host-side metrics with the reference's keys and conventions (src/utils.py:61-94): inputs are
    cast to int and flattened, degenerate denominators give 0, AUROC / AUPRC fall back to 0.0 when
    sklearn rejects the inputs."""
    from sklearn.metrics import auc, confusion_matrix, precision_recall_curve, roc_auc_score
    yt = np.asarray(y_true, dtype=int).reshape(-1)
    yp = np.asarray(y_pred, dtype=int).reshape(-1)
    tn, fp, fn, tp = confusion_matrix(yt, yp).ravel()
    prec = tp / (tp + fp) if (tp + fp) > 0 else 0
    rec = tp / (tp + fn) if (tp + fn) > 0 else 0
    metrics = {
        "accuracy": (tp + tn) / (tp + tn + fp + fn),
        "precision": prec,
        "recall": rec,
        "specificity": tn / (tn + fp) if (tn + fp) > 0 else 0,
        "f1_score": 2 * prec * rec / (prec + rec) if (prec + rec) > 0 else 0,
    }
    if y_scores is not None:
        try:
            metrics["auroc"] = roc_auc_score(yt, y_scores)
            p_curve, r_curve, _ = precision_recall_curve(yt, y_scores)
            metrics["auprc"] = auc(r_curve, p_curve)
        except ValueError:
            metrics["auroc"] = 0.0
            metrics["auprc"] = 0.0
    return metrics


# ---------------------------------------------------------------------------------------------
# validate_epoch's reductions on the device: nothing but a few integers / two
# floats crosses to the host, instead of every score map and mask (src/train_utils.py:192-197)
# ---------------------------------------------------------------------------------------------
def pixel_confusion_counts(masks_pred, masks_true, thresholds=(0.3, 0.5, 0.7), include=None, counts=None,
                           truth_threshold=0.5):
    """Per-threshold (tn, fp, fn, tp) of `(masks_pred > t)` against `(masks_true > truth_threshold)`, the counts
    sklearn's `confusion_matrix(true_flat, pred_flat).ravel()` yields in src/train_utils.py:236-243, as an int64
    [len(thresholds), 4] CUDA tensor.  `include`: optional per-image flags (the reference keeps the images with
    label 1); `counts`: an earlier result to accumulate into (one call per batch)."""
    pred = masks_pred.detach().float().contiguous()
    true = masks_true.detach().float().contiguous()
    if counts is None:
        counts = ops.zero_(torch.empty(len(thresholds), 4, dtype=torch.int64, device=pred.device))
    if include is not None:
        include = torch.as_tensor(np.asarray(include).astype(np.uint8)).to(pred.device, non_blocking=False).contiguous()
    return ops.pixel_confusion(pred, true, include, list(thresholds), counts, truth_threshold)


def metrics_from_confusion(tn, fp, fn, tp):
    """The count-derived entries of `calculate_metrics` (src/utils.py:61-81), same degenerate-denominator rules."""
    tn, fp, fn, tp = (np.int64(v) for v in (tn, fp, fn, tp))
    prec = tp / (tp + fp) if (tp + fp) > 0 else 0
    rec = tp / (tp + fn) if (tp + fn) > 0 else 0
    return {
        "accuracy": (tp + tn) / (tp + tn + fp + fn),
        "precision": prec,
        "recall": rec,
        "specificity": tn / (tn + fp) if (tn + fp) > 0 else 0,
        "f1_score": 2 * prec * rec / (prec + rec) if (prec + rec) > 0 else 0,
    }


def percentile_rank(n, q, dtype=np.float32):
    """(previous index, next index, gamma) numpy's default 'linear' percentile uses for `n` values of `dtype`:
    the virtual index (n - 1) * q / 100 is computed IN the array's dtype (so is gamma), indices at or beyond the
    last element clamp to it."""
    quant = np.true_divide(q, dtype(100))
    if not 0 <= quant <= 1:
        raise ValueError("Percentiles must be in the range [0, 100]")
    virtual = np.asanyarray((n - 1) * quant)
    prev = np.floor(virtual)
    gamma = np.asanyarray(virtual - prev, dtype=virtual.dtype)
    if virtual >= n - 1:
        return n - 1, n - 1, gamma
    if virtual < 0:
        return 0, 0, gamma
    lo = int(prev)
    return lo, lo + 1, gamma


def percentile_lerp(lo, hi, gamma):
    """numpy's `_lerp` on two scalars of the array's dtype."""
    lo, hi = np.asanyarray(lo), np.asanyarray(hi)
    diff = hi - lo
    out = lo + diff * gamma
    if gamma >= 0.5:
        out = hi - diff * (1 - gamma)
    return out[()]


def percentile(scores, q):
    """`np.percentile(scores.cpu().numpy(), q)` (src/train_utils.py:213) without moving the scores: a radix
    selection on the device finds the two neighbouring order statistics, the host interpolates those two floats
    exactly as numpy does.  Returns a numpy float32 scalar, bit-identical to numpy's for NaN-free input."""
    x = scores.detach().float().contiguous().view(-1)
    n = x.numel()
    if n == 0:
        raise IndexError("percentile of an empty tensor")
    lo, hi, gamma = percentile_rank(n, q)
    pair = ops.select_kth(x, lo).cpu().numpy()
    return percentile_lerp(pair[0], pair[1] if hi != lo else pair[0], gamma)
