"""Evaluation-time input transform of the reference's datasets on the device (the
deterministic half): `get_transforms(image_size, is_train=False)` of src/dataset.py:141-152,

    image: transforms.Resize((S, S)) -> ToTensor() -> Normalize(ImageNet mean / std)
    mask : transforms.Resize((S, S)) -> ToTensor()

applied to batches of decoded uint8 images after the host->device copy instead of per image in the loader
workers (Pillow resizes a 1024x1024 MVTec image in a few milliseconds per worker: at >1500 img/s per GPU the
loader is the bottleneck).  Bit-exact against torchvision + Pillow (`csrc/preprocess.cu`).  The random
training augmentations (flip, rotation, colour jitter: src/dataset.py:131-138) stay on the host.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops

IMAGENET_MEAN = (0.485, 0.456, 0.406)      # src/dataset.py:138, 145
IMAGENET_STD = (0.229, 0.224, 0.225)

_plans = {}


def _plan(in_size, out_size, device):
    key = (int(in_size), int(out_size), str(device))
    if key not in _plans:
        bounds, coeffs = ops.resize_coeffs(in_size, out_size)
        _plans[key] = (torch.from_numpy(bounds).to(device), torch.from_numpy(coeffs).to(device))
    return _plans[key]


def _as_u8_batch(x, channels, device):
    """PIL image / HWC uint8 array / uint8 tensor ([H, W, C], [N, H, W, C]; masks [H, W], [N, H, W]) -> CUDA
    uint8 [N, H, W, C]."""
    if not isinstance(x, torch.Tensor):
        x = torch.from_numpy(np.array(x))      # (a copy: PIL hands out read-only buffers)
    if x.dtype != torch.uint8:
        raise TypeError(f"expected decoded uint8 pixels, got {x.dtype}")
    single = False
    if channels == 1:
        if x.dim() == 2:
            x, single = x[None], True
        if x.dim() == 3:
            x = x[..., None]
    else:
        if x.dim() == 3:
            x, single = x[None], True
    if x.dim() != 4 or x.shape[-1] != channels:
        raise ValueError(f"expected [N, H, W, {channels}] uint8 pixels, got {tuple(x.shape)}")
    if device is None:
        device = x.device if x.is_cuda else torch.device("cuda")
    return x.to(device).contiguous(), single


def _resize(x, size, out_u8, out_f32, mean, std):
    oh, ow = (size, size) if isinstance(size, int) else size
    n, h, w, c = x.shape
    tmp = torch.empty(n, h, ow, c, dtype=torch.uint8, device=x.device)
    u8 = torch.empty(n, oh, ow, c, dtype=torch.uint8, device=x.device) if out_u8 else None
    f = torch.empty(n, c, oh, ow, dtype=torch.float32, device=x.device) if out_f32 else None
    ops.resize_u8(x, tmp, u8, f, _plan(w, ow, x.device), _plan(h, oh, x.device), mean, std)
    return u8, f


def resize_u8(images, size, device=None):
    """`PIL.Image.resize((S, S), BILINEAR)` (what transforms.Resize does to a PIL image) on uint8 pixels:
    RGB [H, W, 3] / [N, H, W, 3] or single-channel [H, W] / [N, H, W]; same layout back, on the device."""
    x = images if isinstance(images, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(images)))
    rgb = x.dim() >= 3 and x.shape[-1] == 3
    x, single = _as_u8_batch(x, 3 if rgb else 1, device)
    out, _ = _resize(x, size, True, False, None, None)
    if not rgb:
        out = out[..., 0]
    return out[0] if single else out


def eval_image_transform(images, image_size=256, mean=IMAGENET_MEAN, std=IMAGENET_STD, device=None):
    """The `is_train=False` image transform (src/dataset.py:141-146) -> fp32 [N, 3, S, S] ([3, S, S] for one image)."""
    x, single = _as_u8_batch(images, 3, device)
    _, out = _resize(x, image_size, False, True, mean, std)
    return out[0] if single else out


def eval_mask_transform(masks, image_size=256, device=None):
    """The mask transform (src/dataset.py:148-151): Resize -> ToTensor -> fp32 [N, 1, S, S] ([1, S, S] for one)."""
    x, single = _as_u8_batch(masks, 1, device)
    _, out = _resize(x, image_size, False, True, None, None)
    return out[0] if single else out


def get_transforms(image_size=256, is_train=True):
    """Device-side counterpart of src/dataset.py:127-154 for `is_train=False`: two callables that take a decoded
    PIL image / uint8 array (or a uint8 batch) and return what the reference's transforms return, on the GPU."""
    if is_train:
        raise NotImplementedError("the random training augmentations (flip / rotation / colour jitter) run on the "
                                  "host: use the reference's get_transforms(image_size, is_train=True)")
    return (lambda img: eval_image_transform(img, image_size)), (lambda mask: eval_mask_transform(mask, image_size))
