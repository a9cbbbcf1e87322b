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


_predict_streams = {}


def predict_batches(model, images, batch_size=16, method="mse", want_reconstruction=False, out=None):
    """The per-batch body of the reference's `test_model` loop (src/test.py:80-106: `images.to(device)`, forward,
    `sigmoid` for a plain UNet, `compute_anomaly_score`, `.cpu()` of every result) over a HOST tensor of images
    [N, C, H, W] (pinned or not), as a three-stage pipeline: while batch k runs on the caller's stream, batch k+1 is
    copied in and the results of batch k-1 are copied out on two side streams, through two device slots each.
    Returns pinned host tensors {"anomaly_maps" [N, 1, H, W], "anomaly_scores" [N, H, W]} (+ "reconstructions" when
    asked for; for a plain UNet the reference's dummy reconstruction is the input, so its score map is zero).
    `out`: the dict of an earlier call, to be filled again (pinning the result buffers costs milliseconds).
    The numbers are exactly those of the batch-by-batch loop (eval mode has no atomics)."""
    from .model import AnomalyUNet
    if model.training:
        raise RuntimeError("predict_batches: put the model in eval() mode first")
    if images.is_cuda or images.dim() != 4:
        raise ValueError("predict_batches: images must be a host tensor [N, C, H, W]")
    dev = next(model.parameters()).device
    images = images.float().contiguous()
    n, c, h, w = images.shape
    batch_size = max(1, min(int(batch_size), n))
    two = isinstance(model, AnomalyUNet)
    shapes = {"anomaly_maps": (n, 1, h, w), "anomaly_scores": (n, h, w)}
    if want_reconstruction:
        shapes["reconstructions"] = (n, c, h, w)
    if out is None:
        out = {}
    for key, shape in shapes.items():
        t = out.get(key)
        if t is None or tuple(t.shape) != shape or t.dtype != torch.float32 or t.is_cuda:
            out[key] = torch.empty(shape, dtype=torch.float32).pin_memory()
    with torch.cuda.device(dev), torch.no_grad():
        main = torch.cuda.current_stream(dev)
        if dev not in _predict_streams:
            _predict_streams[dev] = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
        s_in, s_out = _predict_streams[dev]
        x_slot = [torch.empty(batch_size, c, h, w, dtype=torch.float32, device=dev) for _ in range(2)]
        computed = [None, None]     # event: the forward pass that read x_slot[i] / wrote results[i] is done
        copied_out = [None, None]   # event: results[i] have left the device
        results = [None, None]
        s_in.wait_stream(main)
        s_out.wait_stream(main)
        for k, b0 in enumerate(range(0, n, batch_size)):
            i, nb = k & 1, min(batch_size, n - b0)
            with torch.cuda.stream(s_in):
                if computed[i] is not None:
                    s_in.wait_event(computed[i])                 # the slot's previous batch has been consumed ...
                    s_in.wait_event(copied_out[i])               # ... and copied out (a plain UNet's "reconstruction" IS the slot)
                x = x_slot[i][:nb]
                x.copy_(images[b0: b0 + nb], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(s_in)
            main.wait_event(ready)
            if copied_out[i] is not None:
                main.wait_event(copied_out[i])                   # (results[i] are about to be replaced)
            if two:
                recon, amap = model(x)
            else:
                amap = torch.sigmoid(model(x))
                recon = x
            score = compute_anomaly_score(recon, x, method)
            results[i] = (amap, score, recon)
            computed[i] = torch.cuda.Event()
            computed[i].record(main)
            with torch.cuda.stream(s_out):
                s_out.wait_event(computed[i])
                out["anomaly_maps"][b0: b0 + nb].copy_(amap, non_blocking=True)
                out["anomaly_scores"][b0: b0 + nb].copy_(score, non_blocking=True)
                if want_reconstruction:
                    out["reconstructions"][b0: b0 + nb].copy_(recon, non_blocking=True)
                for t in (amap, score, recon):
                    t.record_stream(s_out)
                copied_out[i] = torch.cuda.Event()
                copied_out[i].record(s_out)
        main.wait_stream(s_out)
        main.wait_stream(s_in)
    return out


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
