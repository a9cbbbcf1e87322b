"""Drop-in replacement for the reference's `src/train_utils.py`.

Same public names and call signatures (src/train.py:17-24 imports CombinedLoss, SSIMLoss,
get_optimizer, get_scheduler, train_epoch, validate_epoch); the losses are CUDA kernels
(csrc/head_loss.cu) wrapped in autograd nodes, the optimizer is optim.FusedOptimizer.
"""
from __future__ import annotations

from collections.abc import Mapping

import numpy as np
import torch
import torch.nn as nn
from torch.autograd.function import once_differentiable

from . import ops
from .optim import FusedOptimizer
from .utils import (AverageMeter, calculate_metrics, compute_anomaly_score, metrics_from_confusion, percentile,
                    pixel_confusion_counts)

try:  # progress bars are cosmetic; the reference uses tqdm (src/train_utils.py:5)
    from tqdm import tqdm
except Exception:  # pragma: no cover
    def tqdm(it, **kwargs):
        return it


def _as_f32_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the loss kernels have no CPU path")
    return t.detach().float().contiguous()


# ---------------------------------------------------------------------------------------------
# CombinedLoss — src/train_utils.py:10-44
# ---------------------------------------------------------------------------------------------
class _CombinedLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, reconstruction, anomaly_map, original_image, true_mask, owner):
        r = _as_f32_cuda(reconstruction, "reconstruction")
        a = _as_f32_cuda(anomaly_map, "anomaly_map")
        x = _as_f32_cuda(original_image, "original_image")
        t = _as_f32_cuda(true_mask, "true_mask")
        if r.shape != x.shape or a.shape != t.shape:
            raise ValueError(f"shape mismatch: reconstruction {tuple(r.shape)} vs image {tuple(x.shape)}, "
                             f"anomaly_map {tuple(a.shape)} vs mask {tuple(t.shape)}")
        need_r, need_a = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        d_r = torch.empty_like(r) if need_r else None
        d_a = torch.empty_like(a) if need_a else None
        out3 = torch.empty(3, dtype=torch.float32, device=r.device)
        ops.combined_loss(r, a, x, t, d_r, d_a, out3, owner._workspace(r.device), owner.recon_weight,
                          owner.seg_weight, owner.focal_alpha, owner.focal_gamma)
        ctx.d_r, ctx.d_a = d_r, d_a
        owner._last = out3
        total, rl, sl = out3[0], out3[1], out3[2]
        ctx.mark_non_differentiable(rl, sl)
        return total, rl, sl

    @staticmethod
    @once_differentiable
    def backward(ctx, g_total, _g1, _g2):
        d_r, d_a = ctx.d_r, ctx.d_a
        ctx.d_r = ctx.d_a = None
        g = g_total.detach().float().contiguous().view(1)
        if d_r is not None:
            ops.scale_by_scalar(d_r, g)
        if d_a is not None:
            ops.scale_by_scalar(d_a, g)
        return d_r, d_a, None, None, None


class CombinedLoss(nn.Module):
    """recon_weight * MSE(reconstruction, image) + seg_weight * focal(anomaly_map, mask)."""

    def __init__(self, recon_weight=1.0, seg_weight=1.0, focal_alpha=0.25, focal_gamma=2.0):
        super().__init__()
        self.recon_weight = recon_weight
        self.seg_weight = seg_weight
        self.focal_alpha = focal_alpha
        self.focal_gamma = focal_gamma
        self._work = None
        self._last = None
        self._pinned = None

    def _workspace(self, device):
        if self._work is None or self._work.device != device:
            self._work = ops.combined_loss_workspace(device)
        return self._work

    def focal_loss(self, pred, target):
        """Focal loss alone (src/train_utils.py:23-28): same kernel with recon_weight 0."""
        dummy = ops.zero_(torch.empty(4, dtype=torch.float32, device=pred.device))
        _, _, sl = _FocalOnlyFn.apply(pred, target, self, dummy)
        return sl

    def forward(self, reconstruction, anomaly_map, original_image, true_mask):
        total, rl, sl = _CombinedLossFn.apply(reconstruction, anomaly_map, original_image, true_mask, self)
        return {"total_loss": total, "recon_loss": rl, "seg_loss": sl}

    def last_values(self):
        """(total, recon, seg) of the most recent forward as python floats: ONE device->host copy
        instead of the three `.item()` syncs of src/train_utils.py:137-139."""
        return tuple(self._last.tolist())

    def last_values_async(self):
        """Start the device->host copy of (total, recon, seg) into pinned memory and return a callable that waits
        for it and gives the three floats: train_epoch reads a step's losses while the next step is already
        enqueued, so the GPU never idles on the `.item()` round trip of src/train_utils.py:137-139."""
        if self._pinned is None:
            self._pinned = [torch.empty(3, dtype=torch.float32).pin_memory() for _ in range(2)]
            self._turn = 0
        host = self._pinned[self._turn]
        self._turn ^= 1
        host.copy_(self._last, non_blocking=True)
        done = torch.cuda.Event()
        done.record()

        def wait():
            done.synchronize()
            return tuple(host.tolist())
        return wait


class _FocalOnlyFn(torch.autograd.Function):
    """focal(pred, target) with gradient w.r.t. pred; reuses the combined kernel on a 4-element dummy
    reconstruction pair so that no second kernel is needed."""

    @staticmethod
    def forward(ctx, pred, target, owner, dummy):
        a = _as_f32_cuda(pred, "pred")
        t = _as_f32_cuda(target, "target")
        d_a = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        out3 = torch.empty(3, dtype=torch.float32, device=a.device)
        ops.combined_loss(dummy, a, dummy, t, None, d_a, out3, owner._workspace(a.device), 0.0, 1.0,
                          owner.focal_alpha, owner.focal_gamma)
        ctx.d_a = d_a
        return out3[0], out3[1], out3[2]

    @staticmethod
    @once_differentiable
    def backward(ctx, _g0, _g1, g_seg):
        d_a = ctx.d_a
        ctx.d_a = None
        if d_a is not None:
            ops.scale_by_scalar(d_a, g_seg.detach().float().contiguous().view(1))
        return d_a, None, None, None


# ---------------------------------------------------------------------------------------------
# SSIMLoss — src/train_utils.py:47-104
# ---------------------------------------------------------------------------------------------
class _SSIMFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img1, img2, window_size, size_average):
        a = _as_f32_cuda(img1, "img1")
        b = _as_f32_cuda(img2, "img2")
        if a.shape != b.shape or a.dim() != 4:
            raise ValueError(f"SSIMLoss expects two [N, C, H, W] tensors of equal shape, got {tuple(a.shape)} and {tuple(b.shape)}")
        n = a.shape[0]
        need1, need2 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        maps = None
        if need1 or need2:
            maps = torch.empty((5 if need2 else 3,) + tuple(a.shape), dtype=torch.float32, device=a.device)
        sums = torch.empty(n, dtype=torch.float32, device=a.device)
        ops.ssim_fwd(a, b, sums, maps, window_size, need2)
        out = torch.empty(() if size_average else (n,), dtype=torch.float32, device=a.device)
        ops.ssim_finalize(sums, out, a.shape, size_average)
        ctx.saved = (a, b, maps, window_size, size_average, need1, need2)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        a, b, maps, window_size, size_average, need1, need2 = ctx.saved
        ctx.saved = None
        g = gout.detach().float().contiguous().view(-1)
        g1 = g2 = None
        if need1:
            g1 = torch.empty_like(a)
            ops.ssim_bwd(a, b, maps, g, g1, window_size, 0, not size_average)
        if need2:
            g2 = torch.empty_like(b)
            ops.ssim_bwd(a, b, maps, g, g2, window_size, 1, not size_average)
        return g1, g2, None, None


class SSIMLoss(nn.Module):
    """1 - SSIM with an 11x11 gaussian window (sigma 1.5), zero padding 5."""

    def __init__(self, window_size=11, size_average=True):
        super().__init__()
        self.window_size = window_size
        self.size_average = size_average
        self.channel = 1

    def forward(self, img1, img2):
        self.channel = img1.size(1)
        return _SSIMFn.apply(img1, img2, self.window_size, self.size_average)


# ---------------------------------------------------------------------------------------------
# training / validation loops — src/train_utils.py:107-260
# ---------------------------------------------------------------------------------------------
def _loss_values(criterion, loss_dict):
    if isinstance(criterion, CombinedLoss):
        return criterion.last_values()
    return (loss_dict["total_loss"].item(), loss_dict["recon_loss"].item(), loss_dict["seg_loss"].item())


def _loss_values_later(criterion, loss_dict):
    """A callable giving the three loss floats of this step (resolved one step later by train_epoch)."""
    if isinstance(criterion, CombinedLoss) and loss_dict["total_loss"].is_cuda:
        return criterion.last_values_async()
    values = _loss_values(criterion, loss_dict)
    return lambda: values


_prefetch = {}      # device -> {"stream": side stream, "bufs": {(name, slot, shape, dtype): device tensor}}


def _device_batches(loader, device):
    """Iterate `loader`, yielding (batch, images, masks) with the tensors on `device` (src/train_utils.py:118-119).
    On a CUDA device the host->device copy of the NEXT batch is issued on a side stream before the current step
    is enqueued, so it overlaps the step instead of preceding it (pinned loader memory makes it asynchronous).
    The copies land in two alternating, persistent device buffers per batch shape: nothing is allocated per step,
    and a buffer is overwritten only after the step that read it has been enqueued completely (event).  The
    yielded tensors are therefore valid for ONE step."""
    device = torch.device(device)
    if device.type != "cuda":
        for batch in loader:
            yield batch, batch["image"].to(device), batch["mask"].to(device)
        return
    state = _prefetch.get(device)
    if state is None:
        state = _prefetch[device] = {"stream": torch.cuda.Stream(device), "bufs": {}}
    side, bufs = state["stream"], state["bufs"]
    it = iter(loader)
    released = [None, None]        # per slot: event after the last step that used the slot's buffers
    start = torch.cuda.Event()
    start.record(torch.cuda.current_stream(device))      # buffers may still feed work enqueued before this epoch

    def load(slot):
        try:
            batch = next(it)
        except StopIteration:
            return None
        with torch.cuda.stream(side):
            side.wait_event(released[slot] if released[slot] is not None else start)
            out = []
            for name in ("image", "mask"):
                src = batch[name]
                key = (name, slot, tuple(src.shape), src.dtype)
                buf = bufs.get(key)
                if buf is None:
                    buf = bufs[key] = torch.empty(src.shape, dtype=src.dtype, device=device)
                buf.copy_(src, non_blocking=True)
                out.append(buf)
            ready = torch.cuda.Event()
            ready.record(side)
        return batch, out[0], out[1], ready

    k = 0
    nxt = load(0)
    while nxt is not None:
        batch, images, masks, ready = nxt
        torch.cuda.current_stream(device).wait_event(ready)
        nxt = load((k + 1) % 2)
        yield batch, images, masks
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(device))
        released[k % 2] = done
        k += 1


FUSED_STEP = __import__("os").environ.get("UNETK_FUSED_STEP", "1") != "0"
FUSED_UPDATE = __import__("os").environ.get("UNETK_FUSED_UPDATE", "1") != "0"


def _fusable(model, criterion, images, masks):
    """The loop body of src/train_utils.py:122-135 can run as Engine.train_step: this package's AnomalyUNet in
    training mode, this package's CombinedLoss, float32 CUDA batches of matching shapes, autograd on."""
    from .model import AnomalyUNet
    if not (FUSED_STEP and type(model) is AnomalyUNet and type(criterion) is CombinedLoss and model.training
            and torch.is_grad_enabled()):
        return False
    if not (torch.is_tensor(images) and torch.is_tensor(masks) and images.is_cuda and masks.is_cuda
            and images.dtype == torch.float32 and masks.dtype == torch.float32 and images.dim() == 4 and masks.dim() == 4
            and masks.shape == (images.shape[0], 1, images.shape[2], images.shape[3])
            and not images.requires_grad and not masks.requires_grad):
        return False
    p0 = next(model.parameters())
    return p0.device == images.device and model._engine().can_fuse_step()


def train_step(model, criterion, optimizer, images, masks):
    """One iteration of the reference's training loop body (src/train_utils.py:122-135): forward, CombinedLoss,
    `optimizer.zero_grad()`, backward, `optimizer.step()`.  Returns the loss dict (0-d device tensors, detached:
    the backward pass has already run).  For this package's AnomalyUNet + CombinedLoss the forward pass, the loss
    and the backward pass are ONE launch sequence (Engine.train_step: each decoder's 1x1 head, its loss term and the
    head's backward are a single kernel, and nothing separates the forward from the backward CUDA graph); any other
    combination takes the plain path below, which is also what `model(x)` / `criterion(...)` / `.backward()` called
    one by one do."""
    if _fusable(model, criterion, images, masks):
        eng = model._engine()
        update = None
        if FUSED_UPDATE and isinstance(optimizer, FusedOptimizer):
            # this package's optimizer on the model's flat parameter buffer: its step runs inside the backward pass,
            # group by group as the gradients become final (Engine.train_step, unetk_fused_update)
            eng.ensure(images.device)
            update = optimizer.fused_update_args(eng)
        optimizer.zero_grad()
        try:
            out3, _ = eng.train_step(images, masks, (criterion.recon_weight, criterion.seg_weight,
                                                     criterion.focal_alpha, criterion.focal_gamma), update=update)
        except BaseException:
            if update is not None:
                update.rollback()       # the optimizer step that was handed out did not run
            raise
        criterion._last = out3
        if update is None:
            optimizer.step()
        return {"total_loss": out3[0], "recon_loss": out3[1], "seg_loss": out3[2]}
    reconstruction, anomaly_map = model(images)
    loss_dict = criterion(reconstruction, anomaly_map, images, masks)
    optimizer.zero_grad()
    loss_dict["total_loss"].backward()
    optimizer.step()
    return loss_dict


def train_epoch(model, train_loader, criterion, optimizer, device, epoch):
    """Train for one epoch; returns the same dict of three running averages as the reference."""
    model.train()
    losses, recon_losses, seg_losses = AverageMeter(), AverageMeter(), AverageMeter()
    pbar = tqdm(train_loader, desc=f"Epoch {epoch}")
    def account(pending):
        values, batch_size = pending
        t, r, s = values()
        losses.update(t, batch_size)
        recon_losses.update(r, batch_size)
        seg_losses.update(s, batch_size)
        if hasattr(pbar, "set_postfix"):
            pbar.set_postfix({"Loss": f"{losses.avg:.4f}", "Recon": f"{recon_losses.avg:.4f}",
                              "Seg": f"{seg_losses.avg:.4f}"})

    pending = None
    for batch, images, masks in _device_batches(pbar, device):
        loss_dict = train_step(model, criterion, optimizer, images, masks)
        current = (_loss_values_later(criterion, loss_dict), images.size(0))
        if pending is not None:
            account(pending)        # the previous step's losses: its copy finished while this step was enqueued
        pending = current
    if pending is not None:
        account(pending)
    return {"total_loss": losses.avg, "recon_loss": recon_losses.avg, "seg_loss": seg_losses.avg}


class _Predictions(Mapping):
    """`validate_epoch(...)['predictions']`: the reference's four host arrays (src/train_utils.py:255-260), copied
    from the device the first time one of them is read.  The reference's own caller never reads them
    (src/train.py:241-265), so an epoch of validation no longer moves every map over PCIe."""

    _KEYS = ("labels", "scores", "masks_true", "masks_pred")

    def __init__(self, labels, scores, masks_true, masks_pred):
        self._host = {"labels": labels}
        self._dev = {"scores": scores, "masks_true": masks_true, "masks_pred": masks_pred}

    def __getitem__(self, key):
        if key not in self._host:
            if key not in self._dev:
                raise KeyError(key)
            parts = self._dev.pop(key)
            # np.array(list of per-image arrays), as the reference builds it
            self._host[key] = (torch.cat(parts).cpu().numpy() if parts else np.array([]))
        return self._host[key]

    def __iter__(self):
        return iter(self._KEYS)

    def __len__(self):
        return len(self._KEYS)


_PIXEL_THRESHOLDS = (0.3, 0.5, 0.7)     # src/train_utils.py:235


def validate_epoch(model, val_loader, criterion, device):
    """Validate for one epoch: same flow, keys, thresholds and values as src/train_utils.py:155-260.

    The per-pixel reductions run on the device while the maps are still there: the confusion counts behind
    `pixel_f1_@t` (anomalous images only) are accumulated batch by batch by `unetk_pixel_confusion`, the 95th
    percentile of the score maps by `unetk_select_kth`; `predictions` is materialised on the host only when read.
    """
    model.eval()
    losses, recon_losses, seg_losses = AverageMeter(), AverageMeter(), AverageMeter()
    all_labels, dev_scores, dev_true, dev_pred = [], [], [], []
    counts = None
    with torch.no_grad():
        pbar = tqdm(val_loader, desc="Validation")
        for batch in pbar:
            images = batch["image"].to(device)
            masks = batch["mask"].to(device)
            labels = batch["label"].numpy()
            reconstruction, anomaly_map = model(images)
            loss_dict = criterion(reconstruction, anomaly_map, images, masks)
            batch_size = images.size(0)
            t, r, s = _loss_values(criterion, loss_dict)
            losses.update(t, batch_size)
            recon_losses.update(r, batch_size)
            seg_losses.update(s, batch_size)
            all_labels.extend(labels)
            dev_scores.append(compute_anomaly_score(reconstruction, images))
            dev_true.append(masks)
            dev_pred.append(anomaly_map)
            if np.any(labels == 1):
                counts = pixel_confusion_counts(anomaly_map, masks, _PIXEL_THRESHOLDS, include=(labels == 1),
                                                counts=counts)
            if hasattr(pbar, "set_postfix"):
                pbar.set_postfix({"Loss": f"{losses.avg:.4f}", "Recon": f"{recon_losses.avg:.4f}",
                                  "Seg": f"{seg_losses.avg:.4f}"})
    all_labels = np.array(all_labels)
    predictions = _Predictions(all_labels, dev_scores, dev_true, dev_pred)
    if len(np.unique(all_labels)) > 1:
        # (the reference compares the [N] labels with the [N, H, W] thresholded score maps here and sklearn
        # rejects the lengths: same call, same ValueError)
        threshold = percentile(torch.cat([x.reshape(-1) for x in dev_scores]), 95)
        all_scores = predictions["scores"]
        image_metrics = calculate_metrics(all_labels, (all_scores > threshold).astype(int), all_scores)
    else:
        normal = all_labels[0] == 0
        image_metrics = {"accuracy": 1.0 if normal else 0.0, "precision": 0.0, "recall": 0.0,
                         "specificity": 1.0 if normal else 0.0, "f1_score": 0.0, "auroc": 0.0, "auprc": 0.0}
    pixel_metrics = {}
    if counts is not None:
        for threshold, (tn, fp, fn, tp) in zip(_PIXEL_THRESHOLDS, counts.cpu().numpy()):
            if tp + fn > 0 and tn + fp > 0:        # len(np.unique(true_flat)) > 1
                pixel_metrics[f"pixel_f1_@{threshold}"] = metrics_from_confusion(tn, fp, fn, tp)["f1_score"]
    return {
        "total_loss": losses.avg, "recon_loss": recon_losses.avg, "seg_loss": seg_losses.avg,
        "image_metrics": image_metrics, "pixel_metrics": pixel_metrics,
        "predictions": predictions,
    }


# ---------------------------------------------------------------------------------------------
# optimizer / scheduler factories — src/train_utils.py:263-284
# ---------------------------------------------------------------------------------------------
def get_optimizer(model, optimizer_name="adam", learning_rate=1e-3, weight_decay=1e-4):
    name = optimizer_name.lower()
    if name not in ("adam", "adamw", "sgd"):
        raise ValueError(f"Unknown optimizer: {optimizer_name}")
    engine = model._engine() if hasattr(model, "_engine") else None
    return FusedOptimizer(model.parameters(), name, lr=learning_rate, weight_decay=weight_decay, engine=engine)


def get_scheduler(optimizer, scheduler_name="cosine", num_epochs=100, eta_min=1e-6):
    """Host-side learning-rate schedules, identical to the reference's choices."""
    name = scheduler_name.lower()
    if name == "cosine":
        return torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, T_max=num_epochs, eta_min=eta_min)
    if name == "step":
        return torch.optim.lr_scheduler.StepLR(optimizer, step_size=num_epochs // 3, gamma=0.1)
    if name == "plateau":
        return torch.optim.lr_scheduler.ReduceLROnPlateau(optimizer, mode="min", patience=10, factor=0.5)
    return None
