"""ctypes binding of csrc/build/libunetk.so (the C ABI declared in include/unetk.h).

The product path fails loudly when the CUDA library is missing or a call returns an error code:
there is no PyTorch / CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "build", "libunetk.so")

P, I, L, F, D, Z = c_void_p, c_int, c_int64, c_float, c_double, c_size_t

# name -> (restype, argtypes); must mirror include/unetk.h exactly
SIGNATURES = {
    "unetk_version": (I, []),
    "unetk_last_error": (c_char_p, []),
    "unetk_check_device": (I, []),
    "unetk_memset_zero": (I, [P, Z, P]),
    "unetk_debug_stamp": (I, [P, P]),
    "unetk_conv3x3_fwd": (I, [P, I, P, I, P, P, P, P, P, I, I, I, I, I, I, I, P]),
    "unetk_conv3x3_dgrad": (I, [P, I, P, P, I, I, P, I, I, P, P, I, I, I, I, I, P]),
    "unetk_conv3x3_wgrad": (I, [P, P, I, P, I, P, I, I, I, I, I, I, P]),
    "unetk_first_conv_fwd": (I, [P, P, P, P, I, I, I, I, I, P]),
    "unetk_first_conv_wgrad": (I, [P, P, P, I, I, I, I, I, P]),
    "unetk_image_to_nhwc64": (I, [P, P, I, I, I, I, P]),
    "unetk_first_conv_pack": (I, [P, P, I, P]),
    "unetk_first_conv_wgrad_tc": (I, [P, P, P, P, I, I, I, I, I, P]),
    "unetk_convt2x2_fwd": (I, [P, I, P, P, P, I, I, I, I, P]),
    "unetk_convt2x2_dgrad": (I, [P, I, P, P, I, I, P, I, I, I, P]),
    "unetk_convt2x2_wgrad": (I, [P, I, P, I, P, I, I, I, P]),
    "unetk_bn_finalize": (I, [P, P, P, P, P, P, P, P, P, P, I, D, F, F, P]),
    "unetk_bn_fold": (I, [P, P, P, P, P, P, I, F, P]),
    "unetk_bnrelu_apply": (I, [P, P, P, P, L, I, P]),
    "unetk_bnrelu_pool": (I, [P, P, P, P, P, P, I, I, I, I, I, P]),
    "unetk_bnrelu_apply_fin": (I, [P, P, P, L, I, P]),
    "unetk_bnrelu_pool_fin": (I, [P, P, P, P, P, I, I, I, I, P]),
    "unetk_maxpool_bwd": (I, [P, P, P, I, I, I, I, I, P, P, P]),
    "unetk_bn_bwd_reduce": (I, [P, P, P, P, P, P, P, L, I, P]),
    "unetk_bn_bwd_apply": (I, [P, P, P, P, P, P, P, P, P, P, P, L, I, P]),
    "unetk_head_fwd": (I, [P, P, P, P, I, I, I, I, I, I, P]),
    "unetk_head_fwd_bn": (I, [P, P, P, P, P, I, I, I, I, I, I, P]),
    "unetk_head_loss_workspace_bytes": (Z, []),
    "unetk_head_loss_bwd": (I, [P, P, P, P, P, P, P, P, P, P, P, Z, P, I, I, I, I, I, I, F, F, F, F, P]),
    "unetk_head_bwd": (I, [P, P, P, P, P, P, P, I, I, I, I, I, I, P, P]),
    "unetk_combined_loss_workspace_bytes": (Z, []),
    "unetk_combined_loss": (I, [P, P, P, P, P, P, P, P, Z, L, L, F, F, F, F, P]),
    "unetk_scale_by_scalar": (I, [P, P, L, P]),
    "unetk_ssim_fwd": (I, [P, P, P, P, I, I, I, I, I, I, P]),
    "unetk_ssim_finalize": (I, [P, P, I, I, I, I, I, P]),
    "unetk_ssim_bwd": (I, [P, P, P, P, P, I, I, I, I, I, I, I, P]),
    "unetk_anomaly_score": (I, [P, P, P, I, I, I, I, I, P]),
    "unetk_seg_loss_workspace_bytes": (Z, [I, I]),
    "unetk_seg_loss_fwd": (I, [P, P, P, P, P, Z, I, I, I, I, L, F, F, F, F, F, I, P]),
    "unetk_seg_loss_bwd": (I, [P, P, P, P, P, P, I, I, I, I, L, F, F, F, F, F, I, P]),
    "unetk_channel_scale": (I, [P, P, I, I, I, I, P]),
    "unetk_argmax_channels": (I, [P, P, I, I, I, I, P]),
    "unetk_upsample2x_fwd": (I, [P, P, I, I, I, I, P]),
    "unetk_upsample2x_bwd": (I, [P, P, I, I, I, I, I, P]),
    "unetk_pixel_confusion": (I, [P, P, P, P, I, F, P, I, L, P]),
    "unetk_class_confusion": (I, [P, P, P, I, I, L, L, I, P, P, P]),
    "unetk_select_workspace_bytes": (Z, []),
    "unetk_select_kth": (I, [P, L, L, P, P, Z, P]),
    "unetk_resize_ksize": (I, [I, I]),
    "unetk_resize_coeffs": (I, [I, I, P, P]),
    "unetk_resize_u8": (I, [P, P, P, P, I, I, I, I, I, I, P, P, I, P, P, I, P, P, P]),
    "unetk_flip_rotate_u8": (I, [P, P, I, I, I, I, P, P]),
    "unetk_color_jitter_u8": (I, [P, I, I, I, P, P, P, P]),
    "unetk_u8_to_tensor": (I, [P, P, I, I, I, I, P, P, P]),
    "unetk_resize_nearest_table": (I, [I, I, P]),
    "unetk_resize_nearest_u8": (I, [P, P, P, I, I, I, I, I, P, P, P]),
    "unetk_adam_step": (I, [P, P, P, P, L, F, F, F, F, F, L, F, I, P]),
    "unetk_permute_batched": (I, [P, I, I, P, P, I, F, P]),
    "unetk_permute_chunk_elems": (I, []),
    "unetk_pack_conv_weights": (I, [P, I, I, P, P, P]),
    "unetk_unpack_conv_grads": (I, [P, I, I, I, P, P, F, P]),
    "unetk_fused_update": (I, [P, I, I, I, P, P, P, P, P, P, P, I, P]),
    "unetk_update_plain_chunk_elems": (I, []),
    "unetk_set_floats": (I, [P, P, I, P]),
    "unetk_nchw_to_nhwc": (I, [P, P, I, I, I, I, P]),
    "unetk_nhwc_to_nchw": (I, [P, P, I, I, I, I, P]),
}


class UnetkError(RuntimeError):
    pass


class BnCtx(ctypes.Structure):
    """Mirror of `unetk_bn_ctx` (include/unetk.h): fused BatchNorm-backward first pass."""
    _fields_ = [("y", c_void_p), ("scale", c_void_p), ("shift", c_void_p), ("mean", c_void_p), ("invstd", c_void_p),
                ("sums", c_void_p)]


class BnFinalizeArgs(ctypes.Structure):
    """Mirror of `unetk_bn_finalize_args` (include/unetk.h)."""
    _fields_ = [("stats", c_void_p), ("gamma", c_void_p), ("beta", c_void_p), ("running_mean", c_void_p),
                ("running_var", c_void_p), ("num_batches_tracked", c_void_p), ("scale", c_void_p), ("shift", c_void_p),
                ("mean", c_void_p), ("invstd", c_void_p), ("count", ctypes.c_double), ("inv_count", ctypes.c_double),
                ("unbias", ctypes.c_float), ("eps", ctypes.c_float),
                ("momentum", ctypes.c_float)]


_lib = None
_launches = 0  # number of C-ABI calls made by this process
_kernel_launches = 0  # ... of which launch at least one of our kernels (bench: gpu_launches)


def get():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if os.environ.get("UNETK_AUTOBUILD", "0") == "1":
            from . import build as _build
            _build.build()
        else:
            raise UnetkError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  This package has no CPU / PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so is stale: loud on purpose
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    msg = get().unetk_last_error()
    return msg.decode() if msg else ""


class Timeline:
    """Diagnostics (tools/timeline.py): a %globaltimer stamp before and after every C-ABI call, on the call's
    own stream.  Stamps are kernels, so they are captured into the engine's CUDA graphs and every replay
    rewrites the buffer: read `intervals()` after a synchronize."""

    def __init__(self, device, capacity=8192):
        import torch
        self.buf = torch.zeros(2 * capacity, dtype=torch.int64, device=device)
        self.capacity = capacity
        self.calls = []          # (name, stream handle, note)
        self.note = ""

    def stamp(self, name, stream, fn, args):
        i = len(self.calls)
        if i >= self.capacity:
            return fn(*args)
        self.calls.append((name, stream or 0, self.note))
        lib, base = get(), self.buf.data_ptr()
        lib.unetk_debug_stamp(base + 16 * i, stream)
        rc = fn(*args)
        lib.unetk_debug_stamp(base + 16 * i + 8, stream)
        return rc

    def intervals(self):
        """[(name, stream, note, start_ns, end_ns)] of the LAST execution of every recorded call."""
        t = self.buf.cpu().view(-1, 2).tolist()
        return [(n, s, note, t[i][0], t[i][1]) for i, (n, s, note) in enumerate(self.calls)]


timeline = None   # set to a Timeline to record (tools/timeline.py)


def call(name: str, *args) -> None:
    """Call an int-returning entry point and raise UnetkError on a non-zero code."""
    global _launches, _kernel_launches
    if timeline is not None:
        rc = timeline.stamp(name, args[-1], getattr(get(), name), args)
    else:
        rc = getattr(get(), name)(*args)
    _launches += 1
    if name != "unetk_memset_zero":
        _kernel_launches += 1
    if rc != 0:
        raise UnetkError(f"{name} failed with code {rc}: {last_error()}")


def call_host(name: str, *args) -> None:
    """Call a host-only entry point (no kernel launch, works without a device)."""
    rc = getattr(get(), name)(*args)
    if rc != 0:
        raise UnetkError(f"{name} failed with code {rc}: {last_error()}")


def add_replayed(calls: int, kernels: int) -> None:
    """A CUDA-graph replay launched `kernels` of our kernels (`calls` C-ABI calls were captured)."""
    global _launches, _kernel_launches
    _launches += calls
    _kernel_launches += kernels


def launch_count() -> int:
    return _launches


def kernel_launch_count() -> int:
    return _kernel_launches
