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


# ---------------------------------------------------------------------------------------------
# the random training augmentations of src/dataset.py:131-138, given their drawn parameters:
# RandomHorizontalFlip -> RandomRotation(10) -> ColorJitter(.1, .1, .1, .05) on the resized PIL image.
# Restated from torchvision 0.26 (transforms.py, _functional_pil.py) and Pillow 12.2 (Image.rotate,
# ImageEnhance, Geometry.c affine_fixed, Blend.c, Convert.c rgb2l / rgb2hsv / hsv2rgb); the two colour-space
# conversions and the luma are checked against Pillow over all 2^24 colours (tests/test_preprocess_oracle.py).
# ---------------------------------------------------------------------------------------------
def draw_train_params(n, degrees=10.0, brightness=(0.9, 1.1), contrast=(0.9, 1.1), saturation=(0.9, 1.1),
                      hue=(-0.05, 0.05), p_flip=0.5):
    """The draws torchvision makes per image from torch's global generator, in its order: flip
    (`torch.rand(1) < p`), angle (`uniform_(-d, d)`), then ColorJitter.get_params (`randperm(4)`, b, c, s, h)."""
    import torch
    out = []
    for _ in range(n):
        flip = bool(torch.rand(1) < p_flip)
        angle = float(torch.empty(1).uniform_(float(-degrees), float(degrees)).item())
        order = [int(i) for i in torch.randperm(4)]
        b = float(torch.empty(1).uniform_(brightness[0], brightness[1]))
        c = float(torch.empty(1).uniform_(contrast[0], contrast[1]))
        s = float(torch.empty(1).uniform_(saturation[0], saturation[1]))
        h = float(torch.empty(1).uniform_(hue[0], hue[1]))
        out.append({"flip": flip, "angle": angle, "order": order, "factors": (b, c, s, h)})
    return out


def rotate_matrix(angle, w, h):
    """PIL.Image.rotate's inverse affine matrix (destination -> source) about the image centre."""
    ang = -math.radians(angle)
    m = [round(math.cos(ang), 15), round(math.sin(ang), 15), 0.0, round(-math.sin(ang), 15), round(math.cos(ang), 15), 0.0]
    cx, cy = w / 2, h / 2
    x, y = -cx - 0, -cy - 0
    m[2], m[5] = m[0] * x + m[1] * y + m[2], m[3] * x + m[4] * y + m[5]
    m[2] += cx
    m[5] += cy
    return m


def fix16(v):
    """Geometry.c FIX(): 16.16 fixed point, round half up."""
    return int(math.floor(v * 65536.0 + 0.5))


def affine_fixed_coeffs(m):
    """The six integers of affine_fixed: steps per x / per y and the origin sampled at the pixel centre."""
    return (fix16(m[0]), fix16(m[1]), fix16(m[2] + m[0] * 0.5 + m[1] * 0.5),
            fix16(m[3]), fix16(m[4]), fix16(m[5] + m[3] * 0.5 + m[4] * 0.5))


def rotate_nearest(img, angle):
    """Image.rotate(angle, NEAREST, expand=False, fillcolor=0) of a uint8 [H, W, C] array."""
    h, w = img.shape[:2]
    a = angle % 360.0
    if a == 0:
        return img.copy()
    if a == 180:
        return img[::-1, ::-1].copy()
    if a in (90, 270) and w == h:
        return np.rot90(img, 1 if a == 90 else 3).copy()
    a0, a1, a2, a3, a4, a5 = affine_fixed_coeffs(rotate_matrix(a, w, h))
    ys = np.arange(h, dtype=np.int64)[:, None]
    xs = np.arange(w, dtype=np.int64)[None, :]
    xin = (a2 + a1 * ys + a0 * xs) >> 16
    yin = (a5 + a4 * ys + a3 * xs) >> 16
    ok = (xin >= 0) & (xin < w) & (yin >= 0) & (yin < h)
    out = np.zeros_like(img)
    out[ok] = img[yin[ok], xin[ok]]
    return out


def blend(in1, in2, alpha):
    """Image.blend(im1, im2, alpha): float32 arithmetic, truncation; clamped when extrapolating."""
    f32 = np.float32
    al = f32(alpha)
    if al == 0:
        return in1.copy()
    if al == 1:
        return in2.copy()
    a, b = in1.astype(np.int32), in2.astype(np.int32)
    t = (a.astype(f32) + (al * (b - a).astype(f32)).astype(f32)).astype(f32)
    if 0 <= al <= 1:
        return t.astype(np.int32).astype(np.uint8)
    return np.where(t <= 0, 0, np.where(t >= 255, 255, t.astype(np.int32))).astype(np.uint8)


def luma(rgb):
    """convert('L'): ITU-R 601-2 in 16-bit fixed point."""
    r, g, b = (rgb[..., i].astype(np.int64) for i in range(3))
    return ((r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16).astype(np.uint8)


def rgb_to_hsv(rgb):
    """convert('HSV') (Convert.c rgb2hsv_row): float32 ratios, double for the hue wrap and the scaling."""
    f32, f64 = np.float32, np.float64
    r, g, b = (rgb[..., i].astype(np.int32) for i in range(3))
    maxc = np.maximum(r, np.maximum(g, b))
    minc = np.minimum(r, np.minimum(g, b))
    with np.errstate(all="ignore"):
        cr = (maxc - minc).astype(f32)
        s = cr / maxc.astype(f32)
        rc, gc, bc = (maxc - r).astype(f32) / cr, (maxc - g).astype(f32) / cr, (maxc - b).astype(f32) / cr
        h = np.where(r == maxc, (bc - gc).astype(f64),
                     np.where(g == maxc, 2.0 + rc.astype(f64) - bc, 4.0 + gc.astype(f64) - rc)).astype(f32)
        h = np.fmod(h.astype(f64) / 6.0 + 1.0, 1.0).astype(f32)
        uh = np.clip((h.astype(f64) * 255.0).astype(np.int32), 0, 255)
        us = np.clip((s.astype(f64) * 255.0).astype(np.int32), 0, 255)
    gray = minc == maxc
    return np.stack([np.where(gray, 0, uh), np.where(gray, 0, us), maxc], -1).astype(np.uint8)


def hsv_to_rgb(hsv):
    """HSV -> RGB (Convert.c hsv2rgb)."""
    f32, f64 = np.float32, np.float64
    h, s, v = (hsv[..., i].astype(np.int32) for i in range(3))
    hh = h.astype(f32).astype(f64) * 6.0 / 255.0
    i = np.floor(hh).astype(np.int32)
    f = (hh - i.astype(f32).astype(f64)).astype(f32).astype(f64)
    fs = (s.astype(f32).astype(f64) / 255.0).astype(f32).astype(f64)
    vv = v.astype(f32).astype(f64)

    def rnd(x):
        return np.clip(np.floor(x + 0.5), 0, 255).astype(np.int32)
    p, q, t = rnd(vv * (1.0 - fs)), rnd(vv * (1.0 - fs * f)), rnd(vv * (1.0 - fs * (1.0 - f)))
    sel = i % 6
    r = np.choose(sel, [v, q, p, p, t, v])
    g = np.choose(sel, [t, v, v, q, p, p])
    b = np.choose(sel, [p, p, t, v, v, q])
    z = s == 0
    return np.stack([np.where(z, v, r), np.where(z, v, g), np.where(z, v, b)], -1).astype(np.uint8)


def hue_delta(hue_factor):
    """torchvision adjust_hue: np.int32(hue_factor * 255).astype(np.uint8), added to H with uint8 wrap-around."""
    return int(np.int32(hue_factor * 255).astype(np.uint8))


def contrast_mean(rgb):
    """ImageEnhance.Contrast: int(mean of the luma + 0.5)."""
    l = luma(rgb)
    return int(int(l.astype(np.int64).sum()) / l.size + 0.5)


def color_jitter(img, order, factors):
    """ColorJitter.forward with drawn parameters: the four adjustments in `order` (0 brightness, 1 contrast,
    2 saturation, 3 hue) on a uint8 RGB array."""
    b, c, s, h = factors
    for fn in order:
        if fn == 0:
            img = blend(np.zeros_like(img), img, b)
        elif fn == 1:
            img = blend(np.full_like(img, contrast_mean(img)), img, c)
        elif fn == 2:
            img = blend(np.repeat(luma(img)[..., None], 3, -1), img, s)
        else:
            hsv = rgb_to_hsv(img)
            hsv[..., 0] = (hsv[..., 0].astype(np.int32) + hue_delta(h)).astype(np.uint8)
            img = hsv_to_rgb(hsv)
    return img


def train_image_transform(img_u8, size, params, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)):
    """src/dataset.py:131-138 for one image with the parameters of `draw_train_params`."""
    x = resize(img_u8, size, size)
    if params["flip"]:
        x = x[:, ::-1].copy()
    x = rotate_nearest(x, params["angle"])
    x = color_jitter(x, params["order"], params["factors"])
    return normalize(to_tensor(x), mean, std)


def nearest_table(in_size, out_size):
    """Image.resize(NEAREST) (Geometry.c ImagingScaleAffine): source index per output index; the sampling
    position starts at scale / 2 and is accumulated in double."""
    step = in_size / out_size
    pos = 0.0 + step * 0.5
    t = np.zeros(out_size, np.int32)
    for x in range(out_size):
        xin = -1 if pos < 0.0 else int(pos)
        t[x] = xin if 0 <= xin < in_size else -1
        pos += step
    return t


def label_mask_transform(mask_u8, oh, ow):
    """src/gear_dataset.py:255-258: Resize(NEAREST) -> int64 class indices."""
    m = np.asarray(mask_u8, dtype=np.uint8)
    return m[nearest_table(m.shape[0], oh)][:, nearest_table(m.shape[1], ow)].astype(np.int64)
