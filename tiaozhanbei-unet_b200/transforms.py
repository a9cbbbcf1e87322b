"""Evaluation-time input transform of the reference's datasets on the device (the
deterministic half): `get_transforms(image_size, is_train=False)` of src/dataset.py:141-152,

    image: transforms.Resize((S, S)) -> ToTensor() -> Normalize(ImageNet mean / std)
    mask : transforms.Resize((S, S)) -> ToTensor()

and the `is_train=True` one (src/dataset.py:131-138),

    image: Resize -> RandomHorizontalFlip(0.5) -> RandomRotation(10) -> ColorJitter(.1, .1, .1, .05)
           -> ToTensor() -> Normalize

applied to batches of decoded uint8 images after the host->device copy instead of per image in the loader
workers (Pillow resizes a 1024x1024 MVTec image in a few milliseconds per worker: at >1500 img/s per GPU the
loader is the bottleneck).  Bit-exact against torchvision + Pillow (`csrc/preprocess.cu`): the random
parameters are drawn on the host with the very calls torchvision makes, in its order, from torch's global
generator, so the same seed gives the same batch as the reference's single-process loader.
"""
from __future__ import annotations

import math

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
    x = images if isinstance(images, torch.Tensor) else torch.from_numpy(np.array(images))
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


_tables = {}


def label_mask_transform(masks, image_size=(512, 512), device=None):
    """The target transform of the segmentation datasets (src/gear_dataset.py:255-258,
    src/kolektorsdd_dataset.py:156-159): Resize(size, NEAREST) -> MaskToTensor -> int64 [N, H', W'] class indices
    ([H', W'] for one mask)."""
    x, single = _as_u8_batch(masks, 1, device)
    x = x[..., 0].contiguous()
    oh, ow = (image_size, image_size) if isinstance(image_size, int) else image_size
    n, h, w = x.shape
    tabs = []
    for a, b in ((w, ow), (h, oh)):
        key = (a, b, str(x.device))
        if key not in _tables:
            _tables[key] = torch.from_numpy(ops.resize_nearest_table(a, b)).to(x.device)
        tabs.append(_tables[key])
    out = torch.empty(n, oh, ow, dtype=torch.int64, device=x.device)
    ops.resize_nearest_u8(x, out, None, tabs[0], tabs[1])
    return out[0] if single else out


# ---------------------------------------------------------------------------------------------
# training augmentations: parameters on the host (torchvision's draws), pixels on the device
# ---------------------------------------------------------------------------------------------
def draw_train_params(n, degrees=10.0, brightness=0.1, contrast=0.1, saturation=0.1, hue=0.05, p_flip=0.5):
    """One dict per image with what RandomHorizontalFlip.forward, RandomRotation.get_params and
    ColorJitter.get_params draw (torchvision 0.26), from torch's global generator in the same order."""
    lo_hi = lambda v: (max(0.0, 1.0 - v), 1.0 + v)            # ColorJitter._check_input, center 1
    (b0, b1), (c0, c1), (s0, s1) = lo_hi(brightness), lo_hi(contrast), lo_hi(saturation)
    out = []
    for _ in range(n):
        flip = bool(torch.rand(1) < p_flip)
        angle = float(torch.empty(1).uniform_(float(-degrees), float(degrees)).item())
        order = [int(i) for i in torch.randperm(4)]
        b = float(torch.empty(1).uniform_(b0, b1))
        c = float(torch.empty(1).uniform_(c0, c1))
        s = float(torch.empty(1).uniform_(s0, s1))
        h = float(torch.empty(1).uniform_(-hue, hue))
        out.append({"flip": flip, "angle": angle, "order": order, "factors": (b, c, s, h)})
    return out


def _fix16(v):
    return int(math.floor(v * 65536.0 + 0.5))


def _geometry(flip, angle, w, h):
    """(mode, flip, a0..a5) of include/unetk.h: PIL.Image.rotate's fast paths and its inverse affine matrix about
    the image centre, in the 16.16 fixed point of Pillow's nearest-neighbour affine transform."""
    a = angle % 360.0
    if a == 0:
        return [0, int(flip), 0, 0, 0, 0, 0, 0]
    if a == 180:
        return [2, int(flip), 0, 0, 0, 0, 0, 0]
    if a in (90, 270) and w == h:
        return [3 if a == 90 else 4, int(flip), 0, 0, 0, 0, 0, 0]
    ang = -math.radians(a)
    m = [round(math.cos(ang), 15), round(math.sin(ang), 15), 0.0, round(-math.sin(ang), 15), round(math.cos(ang), 15), 0.0]
    cx, cy = w / 2, h / 2
    x, y = -cx - 0, -cy - 0
    m[2], m[5] = m[0] * x + m[1] * y + m[2], m[3] * x + m[4] * y + m[5]
    m[2] += cx
    m[5] += cy
    # (angles so small that sin rounds to 0 make Pillow take its scale-only path, which samples the same pixels as
    # the fixed-point formula below: identity for cos = 1, the 180 degree mirror for cos = -1)
    return [1, int(flip), _fix16(m[0]), _fix16(m[1]), _fix16(m[2] + m[0] * 0.5 + m[1] * 0.5),
            _fix16(m[3]), _fix16(m[4]), _fix16(m[5] + m[3] * 0.5 + m[4] * 0.5)]


def train_image_transform(images, image_size=256, params=None, mean=IMAGENET_MEAN, std=IMAGENET_STD, device=None,
                          **ranges):
    """The `is_train=True` image transform (src/dataset.py:131-139) on a uint8 batch -> fp32 [N, 3, S, S].
    `params`: the output of `draw_train_params` (drawn here when None, with `ranges` = its keyword arguments:
    the Gear pipeline uses degrees=10, brightness=contrast=saturation=0.2, hue=0.1, src/gear_dataset.py:236-243;
    KolektorSDD degrees=5, src/kolektorsdd_dataset.py:140-147)."""
    x, single = _as_u8_batch(images, 3, device)
    n = x.shape[0]
    if params is None:
        params = draw_train_params(n, **ranges)
    if len(params) != n:
        raise ValueError(f"{len(params)} parameter sets for {n} images")
    oh, ow = (image_size, image_size) if isinstance(image_size, int) else image_size
    if max(oh, ow) >= 32768:
        raise ValueError("image too large for the 16.16 fixed-point rotation")
    resized, _ = _resize(x, image_size, True, False, None, None)
    geom = torch.tensor([_geometry(p["flip"], p["angle"], ow, oh) for p in params], dtype=torch.int32)
    stage_ops = torch.tensor([p["order"] for p in params], dtype=torch.int32)
    fac = []
    for p in params:
        b, c, s, h = p["factors"]
        per_op = {0: b, 1: c, 2: s, 3: float(int(np.int32(h * 255).astype(np.uint8)))}   # hue: uint8 amount added to H
        fac.append([per_op[o] for o in p["order"]])
    factors = torch.tensor(fac, dtype=torch.float32)
    dev = x.device
    rotated = ops.flip_rotate_u8(resized, torch.empty_like(resized), geom.to(dev))
    ops.color_jitter_u8(rotated, stage_ops.to(dev), factors.to(dev))
    out = ops.u8_to_tensor(rotated, torch.empty(n, 3, oh, ow, dtype=torch.float32, device=dev), mean, std)
    return out[0] if single else out


def get_transforms(image_size=256, is_train=True):
    """Device-side counterpart of src/dataset.py:127-154: two callables that take a decoded PIL image / uint8
    array (or a uint8 batch) and return what the reference's transforms return, on the GPU."""
    if is_train:
        return (lambda img: train_image_transform(img, image_size)), (lambda mask: eval_mask_transform(mask, image_size))
    return (lambda img: eval_image_transform(img, image_size)), (lambda mask: eval_mask_transform(mask, image_size))
