"""Segmentation losses of the reference's Gear / KolektorSDD pipelines on the sm_100a kernels
(SURVEY.md §8(f) row 1): drop-in for the loss half of /root/reference/src/metrics.py.

    dice_loss(pred, target, smooth=1e-8)                                   src/metrics.py:233-261
    focal_loss(pred, target, alpha=1, gamma=2, ignore_index=None)          src/metrics.py:264-282
    CombinedSegmentationLoss(ce_weight, dice_weight, focal_weight,
                             ignore_index, class_weights)(pred, target)    src/metrics.py:285-335
    argmax_predictions(logits) == torch.argmax(logits, dim=1)              test.py:98

Same names, argument meaning and error behaviour (e.g. focal with `ignore_index=None` raises the
TypeError F.cross_entropy raises in the reference).  One forward launch computes the cross entropy, dice
and focal sums in a single pass over the logits, one backward launch the gradient of their weighted sum;
there is no CPU path.

    SegmentationMetrics(num_classes, ignore_index=None).update(pred, target)   src/metrics.py:9-45

`update` runs in every step of the segmentation training loops (train.py:146): the reference copies the logits
to the host, takes the argmax there and calls sklearn's confusion_matrix.  Here one kernel fuses the argmax with
the C x C counts and accumulates them on the device; the matrix comes to the host when a `compute_*` method (or
`.confusion_matrix`) is read, i.e. once per epoch.  The derived metrics (src/metrics.py:47-133) are the same
numpy formulas; plotting stays the reference's.
"""
from __future__ import annotations

import numpy as np
import torch
from torch import nn
from torch.autograd.function import once_differentiable

from . import ops

_DEFAULT_IGNORE = -100          # F.cross_entropy's default when the reference passes no ignore_index


def _prep(pred, target):
    if not pred.is_cuda or not target.is_cuda:
        raise RuntimeError("pred and target must be CUDA tensors: the loss kernels have no CPU path")
    if pred.dim() != 4 or target.dim() != 3 or target.shape != (pred.shape[0], pred.shape[2], pred.shape[3]):
        raise ValueError(f"expected pred (N, C, H, W) and target (N, H, W), got {tuple(pred.shape)} and {tuple(target.shape)}")
    if pred.shape[1] > 8:
        raise NotImplementedError("at most 8 classes are supported by the segmentation loss kernels")
    return pred.detach().float().contiguous(), target.detach().long().contiguous()


class _SegLossFn(torch.autograd.Function):
    """total (and its three terms) of the combined segmentation loss; gradient w.r.t. the logits only."""

    @staticmethod
    def forward(ctx, pred, target, class_w, ignore_index, ce_w, dice_w, focal_w, alpha, gamma, is_prob):
        ctx.set_materialize_grads(False)
        z, t = _prep(pred, target)
        n, c = z.shape[:2]
        work = ops.seg_loss_workspace(n, c, z.device)
        out4 = torch.empty(4, dtype=torch.float32, device=z.device)
        ops.seg_loss_fwd(z, t, class_w, out4, work, ignore_index, ce_w, dice_w, focal_w, alpha, gamma, is_prob)
        ctx.saved = (z, t, class_w, work)
        ctx.cfg = (ignore_index, ce_w, dice_w, focal_w, alpha, gamma, is_prob)
        ctx.in_dtype = pred.dtype
        return out4[0], out4[1], out4[2], out4[3]

    @staticmethod
    @once_differentiable
    def backward(ctx, g_total, g_ce, g_dice, g_focal):
        z, t, class_w, work = ctx.saved
        ignore_index, ce_w, dice_w, focal_w, alpha, gamma, is_prob = ctx.cfg
        # the terms are exposed for reporting; gradients flow through `total` (or, for the stand-alone
        # functions, through the single term they return: its weight is the only non-zero one, so the
        # kernel's d total / d logits is that term's gradient)
        gs = [x for x in (g_total, g_ce, g_dice, g_focal) if x is not None]
        if not gs:
            return (None,) * 10
        g = gs[0] if len(gs) == 1 else torch.stack([x.float() for x in gs]).sum()
        dz = torch.empty_like(z)
        ops.seg_loss_bwd(z, t, class_w, work, g.detach().float().contiguous().view(1), dz, ignore_index, ce_w, dice_w,
                         focal_w, alpha, gamma, is_prob)
        return dz.to(ctx.in_dtype), None, None, None, None, None, None, None, None, None


def dice_loss(pred, target, smooth=1e-8):
    """Dice loss on class probabilities `pred` (N, C, H, W) vs integer labels (N, H, W): src/metrics.py:233-261."""
    if smooth != 1e-8:
        raise NotImplementedError("the kernel uses the reference's default smooth=1e-8")
    _, _, dice, _ = _SegLossFn.apply(pred, target, None, _DEFAULT_IGNORE, 0.0, 1.0, 0.0, 1.0, 2.0, True)
    return dice


def focal_loss(pred, target, alpha=1, gamma=2, ignore_index=None):
    """alpha (1 - pt)^gamma ce averaged over all pixels (src/metrics.py:264-282)."""
    if ignore_index is None:
        # the reference forwards ignore_index=None to F.cross_entropy, which rejects it
        raise TypeError("cross_entropy_loss(): argument 'ignore_index' (position 5) must be int, not NoneType")
    _, _, _, focal = _SegLossFn.apply(pred, target, None, int(ignore_index), 0.0, 0.0, 1.0, float(alpha), float(gamma), False)
    return focal


class CombinedSegmentationLoss(nn.Module):
    """ce_weight * CE + dice_weight * Dice + focal_weight * Focal (src/metrics.py:285-335)."""

    def __init__(self, ce_weight=1.0, dice_weight=1.0, focal_weight=0.0, ignore_index=None, class_weights=None):
        super().__init__()
        self.ce_weight = ce_weight
        self.dice_weight = dice_weight
        self.focal_weight = focal_weight
        self.ignore_index = ignore_index
        self.class_weights = class_weights
        if class_weights is not None:
            self.class_weights = torch.tensor(class_weights, dtype=torch.float32)
        self._last = None

    def forward(self, pred, target):
        if self.focal_weight > 0 and self.ignore_index is None:
            raise TypeError("cross_entropy_loss(): argument 'ignore_index' (position 5) must be int, not NoneType")
        cw = self.class_weights
        if cw is not None:
            if cw.numel() != pred.shape[1]:
                raise RuntimeError(f"weight tensor should be defined either for all {pred.shape[1]} classes or no classes "
                                   f"but got weight tensor of shape: [{cw.numel()}]")
            cw = cw.to(pred.device).contiguous()
            self.class_weights = cw
        ii = _DEFAULT_IGNORE if self.ignore_index is None else int(self.ignore_index)
        total, ce, dice, focal = _SegLossFn.apply(pred, target, cw, ii, float(self.ce_weight), float(self.dice_weight),
                                                  float(self.focal_weight), 1.0, 2.0, False)
        self._last = (ce.detach(), dice.detach(), focal.detach())
        return total

    def last_terms(self):
        """(ce, dice, focal) of the most recent forward as python floats (one device->host copy)."""
        return tuple(torch.stack(self._last).tolist())


def argmax_predictions(logits):
    """torch.argmax(logits, dim=1) for NCHW fp32 logits -> int64 (N, H, W) (test.py:98, src/metrics.py:31)."""
    if not logits.is_cuda:
        raise RuntimeError("logits must be a CUDA tensor (no CPU path)")
    z = logits.detach().float().contiguous()
    out = torch.empty(z.shape[0], z.shape[2], z.shape[3], dtype=torch.int64, device=z.device)
    return ops.argmax_channels(z, out)


class DeviceConfusion:
    """The state half of `SegmentationMetrics` (src/metrics.py:12-45) with the counts kept on the device.
    `confusion_matrix` / `total_samples` read like the reference's attributes (numpy int64 [C, C] / int) and
    synchronise; assigning to them (as `reset` does) is supported.  Also used as a mix-in over the reference's own
    class by `dropin.launch`, whose `compute_*` methods then run unmodified."""

    def __init__(self, num_classes, ignore_index=None):
        if num_classes > 8:
            raise NotImplementedError("at most 8 classes are supported by the confusion kernel")
        self.num_classes = num_classes
        self.ignore_index = ignore_index
        self.reset()

    def reset(self):
        self._cm_host = np.zeros((self.num_classes, self.num_classes), dtype=np.int64)
        self._kept_host = 0
        self._dev = None           # [C * C + 1] int64: the counts since the last read, and the kept pixels

    def _flush(self):
        if self._dev is not None:
            vals = self._dev.cpu().numpy()
            self._cm_host = self._cm_host + vals[:-1].reshape(self.num_classes, self.num_classes)
            self._kept_host += int(vals[-1])
            self._dev = None

    @property
    def confusion_matrix(self):
        self._flush()
        return self._cm_host

    @confusion_matrix.setter
    def confusion_matrix(self, value):
        self._flush()
        self._cm_host = np.asarray(value, dtype=np.int64)

    @property
    def total_samples(self):
        self._flush()
        return self._kept_host

    @total_samples.setter
    def total_samples(self, value):
        self._flush()
        self._kept_host = int(value)

    def update(self, pred, target):
        """pred: (N, C, H, W) logits or (N, H, W) class indices; target: (N, H, W) labels.  Host tensors (the
        reference accepts them) are copied to the GPU: the counting itself has no CPU path."""
        if not pred.is_cuda or not target.is_cuda:
            if not torch.cuda.is_available():
                raise RuntimeError("pred and target must be CUDA tensors: the confusion kernel has no CPU path")
            dev = pred.device if pred.is_cuda else (target.device if target.is_cuda else torch.device("cuda"))
            pred, target = pred.to(dev), target.to(dev)
        if pred.dim() == 4:
            if pred.shape[1] != self.num_classes:
                # (torch.argmax over more channels than classes: indices >= num_classes fall out of the matrix,
                # which the kernel handles for index input)
                pred = argmax_predictions(pred)
            else:
                pred = pred.detach().float().contiguous()
        if pred.dtype != torch.float32:
            pred = pred.detach().long().contiguous()
        target = target.detach().long().contiguous()
        if target.numel() == 0:
            return
        if self._dev is None or self._dev.device != target.device:
            self._flush()
            self._dev = ops.zero_(torch.empty(self.num_classes * self.num_classes + 1, dtype=torch.int64,
                                              device=target.device))
        ops.class_confusion(pred, target, self.num_classes, self.ignore_index, self._dev[:-1], self._dev[-1:])


class SegmentationMetrics(DeviceConfusion):
    """src/metrics.py:9-133 without the plotting: same constructor, `update`, `compute_*` and
    `compute_all_metrics` keys; every ratio guards its denominator with max(., 1e-8) like the reference."""

    def _parts(self):
        cm = self.confusion_matrix
        return np.diag(cm), cm.sum(axis=1), cm.sum(axis=0)      # hits, pixels per true class, per predicted class

    def compute_iou(self, per_class=True):
        tp, rows, cols = self._parts()
        iou = tp / np.maximum(rows + cols - tp, 1e-8)
        return iou if per_class else np.nanmean(iou)

    def compute_dice(self, per_class=True):
        tp, rows, cols = self._parts()
        dice = 2 * tp / np.maximum(rows + cols, 1e-8)
        return dice if per_class else np.nanmean(dice)

    def compute_pixel_accuracy(self):
        cm = self.confusion_matrix
        return np.diag(cm).sum() / max(cm.sum(), 1e-8)

    def compute_mean_accuracy(self):
        tp, rows, _ = self._parts()
        return np.nanmean(tp / np.maximum(rows, 1e-8))

    def compute_precision_recall_f1(self, per_class=True):
        tp, rows, cols = self._parts()
        precision = tp / np.maximum(tp + (cols - tp), 1e-8)
        recall = tp / np.maximum(tp + (rows - tp), 1e-8)
        f1 = 2 * precision * recall / np.maximum(precision + recall, 1e-8)
        if per_class:
            return precision, recall, f1
        return np.nanmean(precision), np.nanmean(recall), np.nanmean(f1)

    def compute_all_metrics(self):
        iou, dice = self.compute_iou(), self.compute_dice()
        precision, recall, f1 = self.compute_precision_recall_f1()
        return {
            "iou_per_class": iou, "mean_iou": np.nanmean(iou),
            "dice_per_class": dice, "mean_dice": np.nanmean(dice),
            "pixel_accuracy": self.compute_pixel_accuracy(), "mean_accuracy": self.compute_mean_accuracy(),
            "precision_per_class": precision, "recall_per_class": recall, "f1_per_class": f1,
            "mean_precision": np.nanmean(precision), "mean_recall": np.nanmean(recall), "mean_f1": np.nanmean(f1),
            "confusion_matrix": self.confusion_matrix,
        }


def compute_metrics_from_predictions(predictions, targets, num_classes, class_names=None):
    """src/metrics.py:209-230: all metrics of (N, C, H, W) / (N, H, W) predictions against (N, H, W) labels."""
    calc = SegmentationMetrics(num_classes)
    calc.update(predictions, targets)         # (the reference updates image by image: the same counts)
    return calc.compute_all_metrics()
