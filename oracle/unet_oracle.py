"""ORACLE — test infrastructure only.  Never imported by the product path.

A plain fp32 restatement (functional PyTorch, no nn.Module containers) of the reference's
U-Net hot path, written from the reference's behaviour:

    DoubleConv / Down / Up / OutConv / UNet / AnomalyUNet     /root/reference/src/model.py:6-210
    CombinedLoss (MSE + focal) and SSIMLoss                   /root/reference/src/train_utils.py:10-104
    compute_anomaly_score                                     /root/reference/src/utils.py:205-215
    Adam step built by get_optimizer                          /root/reference/src/train_utils.py:263-272
    SegmentationUNet (Dropout2d on the bottleneck)            /root/reference/src/model.py:111-153
    dice_loss / focal_loss / CombinedSegmentationLoss         /root/reference/src/metrics.py:233-335
    bilinear=True decoders, F.pad path for odd sizes          /root/reference/src/model.py:47-49, 57-61

The arithmetic underneath the reference lives in PyTorch (third party, `torch>=2.0.0` in the
reference's requirements.txt; this image has torch 2.11.0), so the restatement calls the same
functional ops the reference's nn.Modules dispatch to.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this file, and only as the checker.

Parity pinning: the reference ships no tests or golden vectors for this path, so
the oracle is pinned against outputs of the reference itself:
  * tests/golden/*.pt were produced by tests/golden/make_golden*.py (U-Net / AnomalyUNet path, segmentation
    pipeline, odd sizes, bilinear decoders), which import the unmodified reference from /root/reference/src
    and record its outputs, losses and gradients on seeded inputs; tests/test_oracle.py and
    tests/test_seg_oracle.py check this oracle against them (CPU, -m "not gpu");
  * where /root/reference is present (the build container), tests/test_oracle.py also runs the
    live reference side by side with the oracle.
Weights are addressed by the reference's own state_dict keys.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
ENC = [("inc", None), ("down1", "maxpool_conv.1"), ("down2", "maxpool_conv.1"),
       ("down3", "maxpool_conv.1"), ("down4", "maxpool_conv.1")]
WIDTHS = [64, 128, 256, 512, 1024]


# ---------------------------------------------------------------------------------------------
# parameter tables
# ---------------------------------------------------------------------------------------------
def state_dict_spec(arch: str = "anomaly_unet", n_channels: int = 3, n_classes: int = 1, bilinear: bool = False):
    """Ordered {key: (shape, dtype)} exactly as the reference's state_dict().  src/model.py:78-95 (UNet),
    :161-186 (AnomalyUNet); bilinear=True: parameter-free up-sampling, DoubleConv(in, out, in // 2) and a
    512-channel bottleneck (src/model.py:47-49, 90-95)."""
    spec = OrderedDict()
    factor = 2 if bilinear else 1

    def double_conv(prefix, cin, cout, mid=None):
        mid = mid or cout
        for ci, co, (i_conv, i_bn) in zip((cin, mid), (mid, cout), ((0, 1), (3, 4))):
            spec[f"{prefix}.double_conv.{i_conv}.weight"] = ((co, ci, 3, 3), torch.float32)
            spec[f"{prefix}.double_conv.{i_bn}.weight"] = ((co,), torch.float32)
            spec[f"{prefix}.double_conv.{i_bn}.bias"] = ((co,), torch.float32)
            spec[f"{prefix}.double_conv.{i_bn}.running_mean"] = ((co,), torch.float32)
            spec[f"{prefix}.double_conv.{i_bn}.running_var"] = ((co,), torch.float32)
            spec[f"{prefix}.double_conv.{i_bn}.num_batches_tracked"] = ((), torch.int64)

    double_conv("inc", n_channels, 64)
    for k in range(1, 5):
        double_conv(f"down{k}.maxpool_conv.1", WIDTHS[k - 1], WIDTHS[k] // (factor if k == 4 else 1))

    def decoder(suffix, n_out):
        for k in range(1, 5):
            cin = WIDTHS[5 - k]
            if bilinear:
                double_conv(f"up{k}{suffix}.conv", cin, (cin // 2) // (factor if k < 4 else 1), cin // 2)
                continue
            spec[f"up{k}{suffix}.up.weight"] = ((cin, cin // 2, 2, 2), torch.float32)
            spec[f"up{k}{suffix}.up.bias"] = ((cin // 2,), torch.float32)
            double_conv(f"up{k}{suffix}.conv", cin, cin // 2)
        spec[f"outc{suffix}.conv.weight"] = ((n_out, 64, 1, 1), torch.float32)
        spec[f"outc{suffix}.conv.bias"] = ((n_out,), torch.float32)

    if arch == "unet":
        decoder("", n_classes)
    elif arch == "anomaly_unet":
        decoder("_recon", n_channels)
        decoder("_seg", 1)
    else:
        raise ValueError(arch)
    return spec


def is_buffer(key: str) -> bool:
    return key.endswith(("running_mean", "running_var", "num_batches_tracked"))


def synthetic_state_dict(arch="anomaly_unet", seed=0, n_channels=3, n_classes=1, device="cpu", bilinear=False):
    """Deterministic non-trivial weights (NOT the reference initialiser): every tensor is drawn
    from a CPU generator in key order, so the same seed reproduces the same state_dict anywhere."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = OrderedDict()
    for key, (shape, dtype) in state_dict_spec(arch, n_channels, n_classes, bilinear).items():
        if key.endswith("num_batches_tracked"):
            t = torch.zeros((), dtype=torch.int64)
        elif key.endswith("running_var"):
            t = torch.rand(shape, generator=g) * 0.5 + 0.75
        elif key.endswith("running_mean"):
            t = torch.randn(shape, generator=g) * 0.1
        elif len(shape) == 1 and ".double_conv." in key and key.endswith("weight"):
            t = torch.rand(shape, generator=g) * 0.5 + 0.75          # BN gamma
        elif len(shape) == 1:
            t = torch.randn(shape, generator=g) * 0.05                # BN beta / biases
        else:
            fan_in = shape[1] * shape[2] * shape[3]
            if ".up.weight" in key:
                fan_in = shape[0]
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        sd[key] = t.to(device)
    return sd


# ---------------------------------------------------------------------------------------------
# storage contract of the B200 build (optional): bf16 storage, fp32 accumulation
# ---------------------------------------------------------------------------------------------
class _RoundBF16(torch.autograd.Function):
    """Value round trip through bf16 in forward AND backward: marks the tensors the CUDA path keeps
    in bf16 in HBM (conv / transposed-conv outputs, post-ReLU activations, their gradients, the
    packed conv weights).  Everything else (accumulation, BatchNorm statistics, heads, losses,
    optimizer) stays fp32, exactly like the reference algorithm above it."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def _rounder(storage):
    if storage == "fp32":
        return lambda t: t
    if storage == "bf16":
        return _RoundBF16.apply
    raise ValueError(storage)


# ---------------------------------------------------------------------------------------------
# forward
# ---------------------------------------------------------------------------------------------
def _double_conv(sd, prefix, x, training, r, first=False):
    # (conv3x3 pad 1 no bias -> BN -> ReLU) x 2                      src/model.py:13-20
    for i_conv, i_bn in ((0, 1), (3, 4)):
        w = sd[f"{prefix}.double_conv.{i_conv}.weight"]
        if not (first and i_conv == 0):      # the first conv reads the fp32 image with fp32 weights
            w = r(w)
        x = F.conv2d(x, w, None, padding=1)
        if training or (first and i_conv == 0):
            x = r(x)      # eval mode folds BN + ReLU into the conv epilogue: the raw output is never stored
        b = f"{prefix}.double_conv.{i_bn}"
        if training:
            sd[b + ".num_batches_tracked"] += 1
        x = F.batch_norm(x, sd[b + ".running_mean"], sd[b + ".running_var"], sd[b + ".weight"],
                         sd[b + ".bias"], training, BN_MOMENTUM, BN_EPS)
        x = r(F.relu(x))
    return x


def _up(sd, prefix, x1, x2, training, r):
    # ConvTranspose2d(k2, s2) -> pad to the skip's size -> cat([skip, up]) -> DoubleConv
    #                                                                   src/model.py:51,54-66
    if prefix + ".up.weight" in sd:
        x1 = r(F.conv_transpose2d(x1, r(sd[prefix + ".up.weight"]), sd[prefix + ".up.bias"], stride=2))
    else:      # bilinear=True: nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True)   src/model.py:47-49
        x1 = r(F.interpolate(x1, scale_factor=2, mode="bilinear", align_corners=True))
    dy = x2.size(2) - x1.size(2)
    dx = x2.size(3) - x1.size(3)
    x1 = F.pad(x1, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
    return _double_conv(sd, prefix + ".conv", torch.cat([x2, x1], dim=1), training, r)


def _encoder(sd, x, training, r):
    feats = [_double_conv(sd, "inc", x, training, r, first=True)]
    for k in range(1, 5):
        feats.append(_double_conv(sd, f"down{k}.maxpool_conv.1", F.max_pool2d(feats[-1], 2), training, r))
    return feats


def _decoder(sd, feats, suffix, training, r):
    x = feats[4]
    for k in range(1, 5):
        x = _up(sd, f"up{k}{suffix}", x, feats[4 - k], training, r)
    return F.conv2d(x, sd[f"outc{suffix}.conv.weight"], sd[f"outc{suffix}.conv.bias"])


def unet_forward(sd, x, training=False, storage="fp32"):
    """UNet.forward -> logits (no sigmoid).  src/model.py:97-108.  BN buffers in `sd` are updated
    in place when training, like nn.BatchNorm2d.  storage="bf16" applies the build's storage
    contract (see _RoundBF16); the default is the reference's plain fp32."""
    r = _rounder(storage)
    return _decoder(sd, _encoder(sd, x, training, r), "", training, r)


def segmentation_unet_forward(sd, x, training=False, dropout=0.1, storage="fp32"):
    """SegmentationUNet.forward -> logits: the UNet graph with nn.Dropout2d(dropout) on the bottleneck
    (src/model.py:111-153; dropout draws from torch's global generator like the module does)."""
    r = _rounder(storage)
    feats = _encoder(sd, x, training, r)
    if dropout > 0:
        feats[4] = r(F.dropout2d(feats[4], dropout, training))
    return _decoder(sd, feats, "", training, r)


def anomaly_unet_forward(sd, x, training=False, storage="fp32"):
    """AnomalyUNet.forward -> (sigmoid reconstruction, sigmoid anomaly map).  src/model.py:188-210."""
    r = _rounder(storage)
    feats = _encoder(sd, x, training, r)
    recon = torch.sigmoid(_decoder(sd, feats, "_recon", training, r))
    amap = torch.sigmoid(_decoder(sd, feats, "_seg", training, r))
    return recon, amap


# ---------------------------------------------------------------------------------------------
# segmentation losses (Gear / KolektorSDD pipelines) — src/metrics.py:233-335
# ---------------------------------------------------------------------------------------------
def seg_dice_loss(pred, target, smooth=1e-8):
    """src/metrics.py:233-261: pred = probabilities (N, C, H, W), target = labels (N, H, W)."""
    c = pred.size(1)
    onehot = F.one_hot(target, c).permute(0, 3, 1, 2).float()
    p = pred.reshape(pred.size(0), c, -1)
    y = onehot.reshape(onehot.size(0), c, -1)
    inter = (p * y).sum(dim=2)
    union = p.sum(dim=2) + y.sum(dim=2)
    return 1 - ((2 * inter + smooth) / (union + smooth)).mean()


def seg_focal_loss(pred, target, alpha=1, gamma=2, ignore_index=-100):
    """src/metrics.py:264-282 (ignore_index must be an int: the reference's None default is rejected by
    F.cross_entropy)."""
    ce = F.cross_entropy(pred, target, ignore_index=ignore_index, reduction="none")
    pt = torch.exp(-ce)
    return (alpha * (1 - pt) ** gamma * ce).mean()


def combined_segmentation_loss(pred, target, ce_weight=1.0, dice_weight=1.0, focal_weight=0.0, ignore_index=None,
                               class_weights=None):
    """CombinedSegmentationLoss.forward (src/metrics.py:285-335) -> (total, {term: value})."""
    terms = {}
    total = 0
    if ce_weight > 0:
        w = None if class_weights is None else torch.as_tensor(class_weights, dtype=torch.float32, device=pred.device)
        kw = {} if ignore_index is None else {"ignore_index": ignore_index}
        terms["ce"] = F.cross_entropy(pred, target, weight=w, **kw)
        total = total + ce_weight * terms["ce"]
    if dice_weight > 0:
        terms["dice"] = seg_dice_loss(F.softmax(pred, dim=1), target)
        total = total + dice_weight * terms["dice"]
    if focal_weight > 0:
        terms["focal"] = seg_focal_loss(pred, target, ignore_index=ignore_index)
        total = total + focal_weight * terms["focal"]
    return total, terms


# ---------------------------------------------------------------------------------------------
# losses / scoring
# ---------------------------------------------------------------------------------------------
def focal_loss(pred, target, alpha=0.25, gamma=2.0):
    # src/train_utils.py:23-28
    bce = F.binary_cross_entropy(pred, target, reduction="none")
    pt = torch.exp(-bce)
    return (alpha * (1 - pt) ** gamma * bce).mean()


def combined_loss(recon, amap, image, mask, recon_weight=1.0, seg_weight=1.0, alpha=0.25, gamma=2.0):
    # src/train_utils.py:30-44
    rl = F.mse_loss(recon, image)
    sl = focal_loss(amap, mask, alpha, gamma)
    return {"total_loss": recon_weight * rl + seg_weight * sl, "recon_loss": rl, "seg_loss": sl}


def _gauss_window(window_size, channel, sigma=1.5):
    # src/train_utils.py:57-65 (np.exp in float64, cast to fp32, normalise, outer product)
    g = torch.tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)],
                     dtype=torch.float32)
    g = (g / g.sum()).unsqueeze(1)
    w2d = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0)
    return w2d.expand(channel, 1, window_size, window_size).contiguous()


def ssim_loss(img1, img2, window_size=11, size_average=True):
    # src/train_utils.py:67-104
    c = img1.size(1)
    win = _gauss_window(window_size, c).to(img1.device, img1.dtype)
    pad = window_size // 2
    mu1 = F.conv2d(img1, win, padding=pad, groups=c)
    mu2 = F.conv2d(img2, win, padding=pad, groups=c)
    mu1_sq, mu2_sq, mu12 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.conv2d(img1 * img1, win, padding=pad, groups=c) - mu1_sq
    s2 = F.conv2d(img2 * img2, win, padding=pad, groups=c) - mu2_sq
    s12 = F.conv2d(img1 * img2, win, padding=pad, groups=c) - mu12
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    smap = ((2 * mu12 + c1) * (2 * s12 + c2)) / ((mu1_sq + mu2_sq + c1) * (s1 + s2 + c2))
    if size_average:
        return 1 - smap.mean()
    return 1 - smap.mean(1).mean(1).mean(1)


def anomaly_score(recon, orig, method="mse"):
    # src/utils.py:205-215  ('ssim' is the same as 'mse' in the reference)
    if method in ("mse", "ssim"):
        return F.mse_loss(recon, orig, reduction="none").mean(dim=1)
    if method == "l1":
        return F.l1_loss(recon, orig, reduction="none").mean(dim=1)
    raise ValueError(f"Unknown method: {method}")


# ---------------------------------------------------------------------------------------------
# one reference training step (train_epoch body, src/train_utils.py:117-139) and Adam
# ---------------------------------------------------------------------------------------------
def adam_update(p, g, m, v, step, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4):
    """torch.optim.Adam single-tensor math (L2 decay folded into the gradient)."""
    g = g + weight_decay * p
    m.lerp_(g, 1 - betas[0])
    v.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
    bc1 = 1 - betas[0] ** step
    bc2 = 1 - betas[1] ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


def train_step(sd, image, mask, arch="anomaly_unet", opt_state=None, step=1, lr=1e-3, weight_decay=1e-4,
               storage="fp32"):
    """Forward + loss + backward (+ Adam when opt_state is a dict).  Returns (loss dict, grads)."""
    params = {k: v for k, v in sd.items() if not is_buffer(k)}
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    work = OrderedDict((k, leaves[k] if k in leaves else v) for k, v in sd.items())
    if arch == "anomaly_unet":
        recon, amap = anomaly_unet_forward(work, image, training=True, storage=storage)
        losses = combined_loss(recon, amap, image, mask)
    else:
        # UNet cannot go through train_epoch (SURVEY.md §4 defect 2): config 1's step is
        # p = sigmoid(UNet(x)); loss = MSE(p, mask) + focal(p, mask)
        p = torch.sigmoid(unet_forward(work, image, training=True, storage=storage))
        rl = F.mse_loss(p, mask)
        sl = focal_loss(p, mask)
        losses = {"total_loss": rl + sl, "recon_loss": rl, "seg_loss": sl}
    losses["total_loss"].backward()
    grads = OrderedDict((k, leaves[k].grad) for k in leaves)
    if opt_state is not None:
        with torch.no_grad():
            for k in leaves:
                st = opt_state.setdefault(k, {"m": torch.zeros_like(sd[k]), "v": torch.zeros_like(sd[k])})
                adam_update(sd[k], grads[k], st["m"], st["v"], step, lr=lr, weight_decay=weight_decay)
    return {k: float(v.detach()) for k, v in losses.items()}, grads
