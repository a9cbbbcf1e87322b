"""CPU ORACLE (test infrastructure, not product code) for the evaluation-time input transform, SURVEY.md §8(f)
row 4: numpy restatement of what /root/reference/src/dataset.py:141-152 makes torchvision + Pillow compute.

The arithmetic lives in third-party dependencies that are not under /root/reference: torchvision 0.26
(`transforms.Resize` on a PIL image -> `Image.resize(size, BILINEAR)`; `ToTensor`; `Normalize`) and Pillow 12.2
(src/libImaging/Resample.c: `precompute_coeffs`, `normalize_coeffs_8bpc`, `ImagingResampleHorizontal_8bpc`,
`ImagingResampleVertical_8bpc`), restated here from the published algorithm.  Pinned by running those very
libraries (installed in this image, and on the GPU box) on seeded inputs: tests/test_preprocess_oracle.py.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def ksize(in_size, out_size):
    return int(math.ceil(max(in_size / out_size, 1.0))) * 2 + 1


def coeffs(in_size, out_size):
    """precompute_coeffs + normalize_coeffs_8bpc for the bilinear (triangle, support 1) filter over the whole
    axis: (bounds int32 [out, 2] = (first tap, tap count), taps int32 [out, ksize] in 22-bit fixed point)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    k = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    taps = np.zeros((out_size, k), np.int32)
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        ss = 1.0 / filterscale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = []
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            w.append(1.0 - a if a < 1.0 else 0.0)
        ww = 0.0
        for v in w:                                 # plain left-to-right double sum, as the C loop
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            taps[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, taps


def resample_axis(img, out_size, axis):
    """One 8-bit pass: 2^21 + sum(pixel * tap) >> 22, clamped to a byte."""
    bounds, taps = coeffs(img.shape[axis], out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.zeros((out_size,) + src.shape[1:], np.int64)
    for xx in range(out_size):
        first, count = bounds[xx]
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(count):
            acc += src[first + x] * int(taps[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return np.moveaxis(out.astype(np.uint8), 0, axis)


def resize(img, oh, ow):
    """Image.resize((ow, oh), BILINEAR) of a uint8 [H, W] / [H, W, C] array: horizontal pass, then vertical
    (a pass whose size does not change is the identity)."""
    x = np.asarray(img, dtype=np.uint8)
    if ow != x.shape[1]:
        x = resample_axis(x, ow, 1)
    if oh != x.shape[0]:
        x = resample_axis(x, oh, 0)
    return x


def to_tensor(img_u8):
    """transforms.ToTensor on uint8 HWC: CHW float32 / 255."""
    x = np.asarray(img_u8)
    if x.ndim == 2:
        x = x[..., None]
    return (x.transpose(2, 0, 1).astype(np.float32) / np.float32(255)).astype(np.float32)


def normalize(chw, mean, std):
    m = np.asarray(mean, dtype=np.float32)[:, None, None]
    s = np.asarray(std, dtype=np.float32)[:, None, None]
    return ((chw - m) / s).astype(np.float32)


def eval_image_transform(img_u8, size, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """src/dataset.py:141-146."""
    return normalize(to_tensor(resize(img_u8, size, size)), mean, std)


def eval_mask_transform(mask_u8, size):
    """src/dataset.py:148-151."""
    return to_tensor(resize(mask_u8, size, size))
