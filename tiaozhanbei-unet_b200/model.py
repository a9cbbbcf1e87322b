"""Drop-in replacement for the reference's `src/model.py` (UNet / AnomalyUNet and their blocks).

Same constructors, attribute paths, parameter shapes / dtypes / initialisers and state_dict keys
as /root/reference/src/model.py:6-210, so `src/train.py` and `src/test.py` run unchanged and
checkpoints round-trip with the reference.  The modules only HOLD parameters; `forward` hands the
whole network to engine.Engine, which launches the sm_100a kernels of include/unetk.h.  There is
no PyTorch-operator or CPU fallback: on anything but a CUDA sm_100 device forward raises.

`bilinear=True` decoders (src/model.py:47-49) use the bilinear x2 up-sampling kernels instead of the
transposed conv.  Input sizes that are not multiples of 16 take the reference's F.pad path
(src/model.py:57-61) inside the kernels: the up-sampled tensor is read as zero beyond its extent.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
from torch.autograd.function import once_differentiable

from .engine import Engine


# ---------------------------------------------------------------------------------------------
# parameter holders: same parameter names / shapes / init as the torch.nn layers of the reference
# ---------------------------------------------------------------------------------------------
class _Conv2dParams(nn.Module):
    """Stands in for nn.Conv2d(cin, cout, k, padding=k//2, bias=bias): weight OIHW (+bias)."""

    def __init__(self, in_channels, out_channels, kernel_size, bias):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, kernel_size, kernel_size))
        self.bias = nn.Parameter(torch.empty(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        # nn.Conv2d.reset_parameters: kaiming_uniform_(a=sqrt(5)); bias ~ U(+-1/sqrt(fan_in))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in = self.in_channels * self.kernel_size * self.kernel_size
            bound = 1 / math.sqrt(fan_in) if fan_in > 0 else 0
            nn.init.uniform_(self.bias, -bound, bound)

    def extra_repr(self):
        return f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, bias={self.bias is not None}"


class _ConvTranspose2dParams(nn.Module):
    """Stands in for nn.ConvTranspose2d(cin, cout, kernel_size=2, stride=2): weight [cin, cout, 2, 2] + bias."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = nn.Parameter(torch.empty(in_channels, out_channels, 2, 2))
        self.bias = nn.Parameter(torch.empty(out_channels))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        fan_in = out_channels * 4      # torch computes fan_in from weight.size(1) for transposed convs
        bound = 1 / math.sqrt(fan_in)
        nn.init.uniform_(self.bias, -bound, bound)

    def extra_repr(self):
        return f"{self.in_channels}, {self.out_channels}, kernel_size=2, stride=2"


class _BatchNorm2dParams(nn.Module):
    """Stands in for nn.BatchNorm2d(c): weight, bias, running_mean, running_var, num_batches_tracked."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))

    def extra_repr(self):
        return f"{self.num_features}, eps={self.eps}, momentum={self.momentum}"


class _Slot(nn.Module):
    """Parameter-less placeholder keeping nn.Sequential indices aligned with the reference
    (ReLU at double_conv.2/.5, MaxPool2d at maxpool_conv.0)."""

    def __init__(self, what):
        super().__init__()
        self.what = what

    def extra_repr(self):
        return self.what


def _no_standalone(name):
    raise NotImplementedError(
        f"{name} is a parameter container in this implementation: run it through UNet / AnomalyUNet, "
        "whose forward executes the fused sm_100a kernel plan (see engine.py)")


class DoubleConv(nn.Module):
    """(convolution => [BN] => ReLU) * 2 — src/model.py:6-23"""

    def __init__(self, in_channels, out_channels, mid_channels=None):
        super().__init__()
        if not mid_channels:
            mid_channels = out_channels
        self.double_conv = nn.Sequential(
            _Conv2dParams(in_channels, mid_channels, 3, bias=False),
            _BatchNorm2dParams(mid_channels),
            _Slot("ReLU(inplace=True) — fused"),
            _Conv2dParams(mid_channels, out_channels, 3, bias=False),
            _BatchNorm2dParams(out_channels),
            _Slot("ReLU(inplace=True) — fused"),
        )

    def forward(self, x):
        _no_standalone("DoubleConv")


class Down(nn.Module):
    """Downscaling with maxpool then double conv — src/model.py:26-37"""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.maxpool_conv = nn.Sequential(_Slot("MaxPool2d(2) — fused with the producer's BN+ReLU"),
                                          DoubleConv(in_channels, out_channels))

    def forward(self, x):
        _no_standalone("Down")


class Up(nn.Module):
    """Upscaling then double conv — src/model.py:40-66 (transposed-conv branch)"""

    def __init__(self, in_channels, out_channels, bilinear=True):
        super().__init__()
        self.bilinear = bool(bilinear)
        if bilinear:      # src/model.py:47-49: parameter-free up-sampling, the DoubleConv narrows through in // 2
            self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
            self.conv = DoubleConv(in_channels, out_channels, in_channels // 2)
        else:
            self.up = _ConvTranspose2dParams(in_channels, in_channels // 2)
            self.conv = DoubleConv(in_channels, out_channels)

    def forward(self, x1, x2):
        _no_standalone("Up")


class OutConv(nn.Module):
    """1x1 convolution with bias — src/model.py:69-75"""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = _Conv2dParams(in_channels, out_channels, 1, bias=True)

    def forward(self, x):
        _no_standalone("OutConv")


# ---------------------------------------------------------------------------------------------
# whole-network autograd node
# ---------------------------------------------------------------------------------------------
class _NetFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, x, *params):
        outs, plan, gen = engine.forward(x, training=True)
        ctx.engine, ctx.plan, ctx.gen = engine, plan, gen
        outs = tuple(outs)
        ctx.save_for_backward(*outs)      # the sigmoid heads need their outputs in backward
        return outs

    @staticmethod
    @once_differentiable
    def backward(ctx, *grad_outs):
        grads = ctx.engine.backward(ctx.plan, ctx.gen, grad_outs, ctx.saved_tensors)
        return (None, None) + tuple(grads)


class _EngineModule(nn.Module):
    """Shared forward dispatch of UNet / AnomalyUNet."""

    def _engine(self) -> Engine:
        eng = self.__dict__.get("_unetk_engine")
        if eng is None:
            eng = Engine(self, self._arch, self.n_channels, self._head_specs())
            self.__dict__["_unetk_engine"] = eng   # not a submodule / not in state_dict
        return eng

    def _run(self, x):
        eng = self._engine()
        p0 = next(self.parameters())
        if p0.device != x.device:
            raise RuntimeError(f"Input is on {x.device} but the model parameters are on {p0.device}")
        if self.training:
            if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
                eng.ensure(x.device)
                return _NetFunction.apply(eng, x, *eng.params)
            outs, _, _ = eng.forward(x, training=True)
            return tuple(outs)
        outs, _, _ = eng.forward(x, training=False)
        return tuple(outs)

    def load_state_dict(self, *args, **kwargs):
        r = super().load_state_dict(*args, **kwargs)
        eng = self.__dict__.get("_unetk_engine")
        if eng is not None:
            eng.mark_weights_dirty()
        return r


class UNet(_EngineModule):
    """src/model.py:78-108 — forward returns logits [N, n_classes, H, W] (no sigmoid)."""
    _arch = "unet"

    def __init__(self, n_channels=3, n_classes=1, bilinear=False):
        super().__init__()
        self.n_channels = n_channels
        self.n_classes = n_classes
        self.bilinear = bilinear
        self.inc = DoubleConv(n_channels, 64)
        self.down1 = Down(64, 128)
        self.down2 = Down(128, 256)
        self.down3 = Down(256, 512)
        factor = 2 if bilinear else 1
        self.down4 = Down(512, 1024 // factor)
        self.up1 = Up(1024, 512 // factor, bilinear)
        self.up2 = Up(512, 256 // factor, bilinear)
        self.up3 = Up(256, 128 // factor, bilinear)
        self.up4 = Up(128, 64, bilinear)
        self.outc = OutConv(64, n_classes)

    def _head_specs(self):
        return [("", self.n_classes, False)]

    def forward(self, x):
        return self._run(x)[0]


class AnomalyUNet(_EngineModule):
    """src/model.py:156-210 — shared encoder, reconstruction decoder (sigmoid, n_channels) and
    anomaly-segmentation decoder (sigmoid, 1 channel)."""
    _arch = "anomaly_unet"

    def __init__(self, n_channels=3, bilinear=False):
        super().__init__()
        self.n_channels = n_channels
        self.bilinear = bilinear
        self.inc = DoubleConv(n_channels, 64)
        self.down1 = Down(64, 128)
        self.down2 = Down(128, 256)
        self.down3 = Down(256, 512)
        factor = 2 if bilinear else 1
        self.down4 = Down(512, 1024 // factor)
        self.up1_recon = Up(1024, 512 // factor, bilinear)
        self.up2_recon = Up(512, 256 // factor, bilinear)
        self.up3_recon = Up(256, 128 // factor, bilinear)
        self.up4_recon = Up(128, 64, bilinear)
        self.outc_recon = OutConv(64, n_channels)
        self.up1_seg = Up(1024, 512 // factor, bilinear)
        self.up2_seg = Up(512, 256 // factor, bilinear)
        self.up3_seg = Up(256, 128 // factor, bilinear)
        self.up4_seg = Up(128, 64, bilinear)
        self.outc_seg = OutConv(64, 1)

    def _head_specs(self):
        return [("_recon", self.n_channels, True), ("_seg", 1, True)]

    def forward(self, x):
        reconstruction, anomaly_map = self._run(x)
        return reconstruction, anomaly_map


class SegmentationUNet(_EngineModule):
    """src/model.py:111-153 — UNet for multi-class segmentation (Gear / KolektorSDD pipelines): the UNet graph
    with nn.Dropout2d on the bottleneck and an n_classes-channel head; forward returns logits.  The state_dict
    is the UNet's (Dropout2d has no parameters)."""
    _arch = "unet"

    def __init__(self, n_channels=3, n_classes=4, bilinear=False, dropout=0.1):
        super().__init__()
        if n_classes > 4:
            raise NotImplementedError("the head kernels support up to 4 output channels")
        self.n_channels = n_channels
        self.n_classes = n_classes
        self.bilinear = bilinear
        self.inc = DoubleConv(n_channels, 64)
        self.down1 = Down(64, 128)
        self.down2 = Down(128, 256)
        self.down3 = Down(256, 512)
        factor = 2 if bilinear else 1
        self.down4 = Down(512, 1024 // factor)
        # like the reference: a Dropout2d module (or Identity) that holds p; the engine applies it
        self.dropout = nn.Dropout2d(dropout) if dropout > 0 else nn.Identity()
        self.up1 = Up(1024, 512 // factor, bilinear)
        self.up2 = Up(512, 256 // factor, bilinear)
        self.up3 = Up(256, 128 // factor, bilinear)
        self.up4 = Up(128, 64, bilinear)
        self.outc = OutConv(64, n_classes)

    def _head_specs(self):
        return [("", self.n_classes, False)]

    def _engine(self):
        eng = super()._engine()
        eng.bottleneck_dropout = float(self.dropout.p) if isinstance(self.dropout, nn.Dropout2d) else 0.0
        return eng

    def forward(self, x):
        return self._run(x)[0]
