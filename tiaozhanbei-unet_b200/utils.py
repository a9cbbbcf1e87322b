"""The parts of the reference's `src/utils.py` that the hot path touches.

`compute_anomaly_score` (src/utils.py:205-215) is a CUDA kernel here; the checkpoint helpers keep the
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
