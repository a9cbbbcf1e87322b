"""Flat-import shim: `from train_utils import ...` (src/train.py:17-24)."""
from tiaozhanbei_unet_b200.train_utils import (CombinedLoss, SSIMLoss, get_optimizer, get_scheduler,  # noqa: F401
                                               train_epoch, validate_epoch)
