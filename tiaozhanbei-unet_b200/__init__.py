"""B200-native (sm_100a) U-Net / AnomalyUNet hot path behind the reference's Python surface.

Import name: ``tiaozhanbei_unet_b200`` (the repository directory is ``tiaozhanbei-unet_b200``;
``tiaozhanbei_unet_b200.py`` at the repo root maps one onto the other).

Sub-modules
    model        UNet / AnomalyUNet / DoubleConv / Down / Up / OutConv   (reference: src/model.py)
    train_utils  CombinedLoss / SSIMLoss / get_optimizer / train_epoch … (reference: src/train_utils.py)
    utils        compute_anomaly_score, AverageMeter, checkpoint I/O      (reference: src/utils.py)
    ops, _lib    tensor-level wrappers and the ctypes binding of include/unetk.h
    engine       static execution plan (buffers + kernel sequence) of one network
"""
__version__ = "0.1.0"
