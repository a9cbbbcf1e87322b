"""Flat-import shim: `from model import AnomalyUNet, UNet` (src/train.py:16,179; src/test.py:15)."""
from tiaozhanbei_unet_b200.model import (AnomalyUNet, DoubleConv, Down, OutConv, SegmentationUNet, UNet,  # noqa: F401
                                         Up)
