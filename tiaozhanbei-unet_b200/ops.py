"""Tensor-level wrappers around the C ABI (include/unetk.h).

PyTorch is plumbing here: it owns device memory and the current stream; every computation is a
call into libunetk.so.  All wrappers are asynchronous on torch's current CUDA stream.
Activations are NHWC bf16 tensors of shape [N, H, W, C]; images / outputs are NCHW fp32.
"""
from __future__ import annotations

import torch

from . import _lib

bf16 = torch.bfloat16
f32 = torch.float32


def _p(t):
    if t is None:
        return None
    return t.data_ptr()


def _s():
    # the CURRENT device's current stream: every public wrapper below runs under _on_tensor_device, which makes
    # the tensors' device current first
    return torch.cuda.current_stream().cuda_stream


def _chk(t, dtype, name):
    if not t.is_cuda:
        raise _lib.UnetkError(f"{name}: expected a CUDA tensor (this library has no CPU path)")
    if t.dtype != dtype:
        raise _lib.UnetkError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.UnetkError(f"{name}: expected a contiguous tensor")


# Optional instrumentation (bench.py roofline): CUDA events around every tensor-core conv launch.
_timing = None


def timing_enabled():
    return _timing is not None


def enable_conv_timing():
    global _timing
    _timing = []
    return _timing


def collect_conv_timing(rec):
    """-> (algorithmic FLOPs, summed kernel milliseconds, launches); disables the instrumentation."""
    global _timing
    _timing = None
    flops = sum(r[0] for r in rec)
    ms = sum(r[1].elapsed_time(r[2]) for r in rec)
    return flops, ms, len(rec)


def conv_timing_by_kernel(rec):
    """{kernel variant: (algorithmic FLOPs, milliseconds, launches)} of an instrumentation record (call
    before collect_conv_timing's result is discarded; events must have completed)."""
    out = {}
    for flops, e0, e1, tag in rec:
        f, ms, n = out.get(tag, (0.0, 0.0, 0))
        out[tag] = (f + flops, ms + e0.elapsed_time(e1), n + 1)
    return out


def _igemm_variant(ncols):
    """Name of the igemm_kernel<BN, SUBS, CTAS> instantiation launch_igemm (csrc/igemm.cu) picks for a GEMM
    with `ncols` output columns."""
    if ncols % 256 == 0:
        return "igemm_kernel<256,1,2>"
    return "igemm_kernel<128,2,2>" if ncols % 128 == 0 else "igemm_kernel<64,2,2>"


def _timed(flops, tag, name, *args):
    if _timing is None:
        _lib.call(name, *args)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.call(name, *args)
    e1.record()
    _timing.append((flops, e0, e1, tag))


def bn_ctx(y, scale, shift, mean, invstd, sums):
    """Host-side `unetk_bn_ctx` for the kernels that can fuse the BatchNorm-backward reduction (sums=None: only
    the layer's y / scale / shift, for head_bwd(a=None))."""
    c = _lib.BnCtx(_p(y), _p(scale), _p(shift), _p(mean), _p(invstd), _p(sums))
    c._keep = (y, scale, shift, mean, invstd, sums)
    return c


def _bn(bn):
    import ctypes
    return ctypes.byref(bn) if bn is not None else None


def zero_(t):
    _lib.call("unetk_memset_zero", _p(t), t.numel() * t.element_size(), _s())
    return t


# ---------------------------------------------------------------------------------------------
# convolutions
# ---------------------------------------------------------------------------------------------
def first_conv_pack(w, wf):
    """OIHW fp32 [64, cin, 3, 3] -> Wf[64][9][64] bf16, zero beyond cin."""
    _chk(w, f32, "w"); _chk(wf, bf16, "wf")
    _lib.call("unetk_first_conv_pack", _p(w), _p(wf), w.shape[1], _s())


def conv3x3_fwd(src1, src2, wf, y, stats=None, scale=None, shift=None, relu=False, algo_cin=None):
    """y = conv3x3(cat(src1, src2)) with packed weights wf [Cout][9][Cin]."""
    _chk(src1, bf16, "src1"); _chk(wf, bf16, "wf"); _chk(y, bf16, "y")
    n, h, w, c1 = src1.shape
    c2 = h2 = w2 = 0
    if src2 is not None:
        _chk(src2, bf16, "src2")
        _, h2, w2, c2 = src2.shape       # may be up to one row / column short of src1 (F.pad path, src/model.py:57-61)
    cout = y.shape[3]
    # algo_cin: the layer's real input channels when src1 is zero-padded (first conv): only algorithmic FLOPs count
    _timed(2.0 * n * h * w * cout * 9 * (algo_cin if algo_cin is not None else c1 + c2), _igemm_variant(cout), "unetk_conv3x3_fwd", _p(src1), c1, _p(src2), c2, _p(wf), _p(y), _p(stats), _p(scale),
              _p(shift), int(relu), n, h, w, cout, h2, w2, _s())
    return y


def conv3x3_dgrad(dy, wd, dx1, acc1=False, dx2=None, acc2=False, stats=None, bn=None):
    _chk(dy, bf16, "dy"); _chk(wd, bf16, "wd"); _chk(dx1, bf16, "dx1")
    n, h, w, cout = dy.shape
    c1 = dx1.shape[3]
    c2 = h2 = w2 = 0
    if dx2 is not None:
        _chk(dx2, bf16, "dx2")
        _, h2, w2, c2 = dx2.shape
    _timed(2.0 * n * h * w * cout * 9 * (c1 + c2), _igemm_variant(c1 + c2), "unetk_conv3x3_dgrad", _p(dy), cout, _p(wd), _p(dx1), c1, int(acc1), _p(dx2), c2,
           int(acc2), _p(stats), _bn(bn), n, h, w, h2, w2, _s())


def conv3x3_wgrad(dy, src1, src2, dw):
    """dw[co][tap][ci] += ... (fp32, caller zeroes)."""
    _chk(dy, bf16, "dy"); _chk(src1, bf16, "src1"); _chk(dw, f32, "dw")
    n, h, w, cout = dy.shape
    c1 = src1.shape[3]
    c2 = h2 = w2 = 0
    if src2 is not None:
        _chk(src2, bf16, "src2")
        _, h2, w2, c2 = src2.shape
    _timed(2.0 * n * h * w * cout * 9 * (c1 + c2), "wgrad_kernel", "unetk_conv3x3_wgrad", _p(dy), _p(src1), c1, _p(src2), c2, _p(dw), n, h, w, cout,
           h2, w2, _s())


def first_conv_fwd(x, w, y, stats=None):
    _chk(x, f32, "x"); _chk(w, f32, "w"); _chk(y, bf16, "y")
    n, cin, h, ww = x.shape
    _lib.call("unetk_first_conv_fwd", _p(x), _p(w), _p(y), _p(stats), n, cin, h, ww, y.shape[3], _s())
    return y


def first_conv_wgrad(x, dy, dw):
    _chk(x, f32, "x"); _chk(dy, bf16, "dy"); _chk(dw, f32, "dw")
    n, cin, h, ww = x.shape
    _lib.call("unetk_first_conv_wgrad", _p(x), _p(dy), _p(dw), n, cin, h, ww, dy.shape[3], _s())


def image_to_nhwc64(x, x64):
    """channels [0, 8) of the zero-initialised NHWC bf16 buffer x64[n, h, w, 64] <- NCHW fp32 image."""
    _chk(x, f32, "x"); _chk(x64, bf16, "x64")
    n, cin, h, ww = x.shape
    _lib.call("unetk_image_to_nhwc64", _p(x), _p(x64), n, cin, h, ww, _s())
    return x64


def first_conv_wgrad_tc(x64, dy, scratch, dw):
    """dw (OIHW [64, cin, 3, 3], overwritten) through the tcgen05 weight-gradient kernel on the
    channel-padded image; scratch: 64*9*64 floats."""
    _chk(x64, bf16, "x64"); _chk(dy, bf16, "dy"); _chk(scratch, f32, "scratch"); _chk(dw, f32, "dw")
    n, h, ww, _ = x64.shape
    _lib.call("unetk_first_conv_wgrad_tc", _p(x64), _p(dy), _p(scratch), _p(dw), n, dw.shape[1], h, ww, dy.shape[3], _s())


def convt2x2_fwd(a, wt, bias, up):
    _chk(a, bf16, "a"); _chk(wt, bf16, "wt"); _chk(up, bf16, "up")
    n, h, w, cin = a.shape
    _timed(2.0 * n * h * w * cin * 4 * up.shape[3], _igemm_variant(4 * up.shape[3]), "unetk_convt2x2_fwd", _p(a), cin, _p(wt), _p(bias), _p(up), n, h, w, up.shape[3], _s())
    return up


def convt2x2_dgrad(dup, wtd, da, acc=False, bn=None):
    _chk(dup, bf16, "dup"); _chk(wtd, bf16, "wtd"); _chk(da, bf16, "da")
    n, h, w, cin = da.shape
    _timed(2.0 * n * h * w * cin * 4 * dup.shape[3], _igemm_variant(cin), "unetk_convt2x2_dgrad", _p(dup), dup.shape[3], _p(wtd), _p(da), cin, int(acc), _bn(bn), n, h, w, _s())


def convt2x2_wgrad(a, dup, dw):
    _chk(a, bf16, "a"); _chk(dup, bf16, "dup"); _chk(dw, f32, "dw")
    n, h, w, cin = a.shape
    _timed(2.0 * n * h * w * cin * 4 * dup.shape[3], "wgrad_kernel", "unetk_convt2x2_wgrad", _p(a), cin, _p(dup), dup.shape[3], _p(dw), n, h, w, _s())


# ---------------------------------------------------------------------------------------------
# batch norm / relu / pool
# ---------------------------------------------------------------------------------------------
def bn_finalize(stats, gamma, beta, running_mean, running_var, nbt, scale, shift, mean, invstd,
                count, eps=1e-5, momentum=0.1):
    c = gamma.numel()
    _lib.call("unetk_bn_finalize", _p(stats), _p(gamma), _p(beta), _p(running_mean), _p(running_var),
              _p(nbt), _p(scale), _p(shift), _p(mean), _p(invstd), c, float(count), eps, momentum, _s())


def bn_fold(gamma, beta, running_mean, running_var, scale, shift, eps=1e-5):
    _lib.call("unetk_bn_fold", _p(gamma), _p(beta), _p(running_mean), _p(running_var), _p(scale),
              _p(shift), gamma.numel(), eps, _s())


def bnrelu_apply(y, scale, shift, a):
    _chk(y, bf16, "y"); _chk(a, bf16, "a")
    c = y.shape[-1]
    _lib.call("unetk_bnrelu_apply", _p(y), _p(scale), _p(shift), _p(a), y.numel() // c, c, _s())
    return a


def bn_finalize_args(stats, gamma, beta, rm, rv, nbt, scale, shift, mean, invstd, count, eps, momentum):
    """Host-side `unetk_bn_finalize_args` for bnrelu_apply_fin / bnrelu_pool_fin."""
    f = _lib.BnFinalizeArgs(_p(stats), _p(gamma), _p(beta), _p(rm), _p(rv), _p(nbt), _p(scale), _p(shift), _p(mean),
                            _p(invstd), float(count), 0.0, 0.0, float(eps), float(momentum))
    f._keep = (stats, gamma, beta, rm, rv, nbt, scale, shift, mean, invstd)
    return f


def bnrelu_apply_fin(y, fin, a):
    """a = relu(BatchNorm_train(y)) with the statistics finalize folded in (see bn_finalize_args)."""
    import ctypes
    _chk(y, bf16, "y"); _chk(a, bf16, "a")
    c = y.shape[-1]
    _lib.call("unetk_bnrelu_apply_fin", _p(y), ctypes.byref(fin), _p(a), y.numel() // c, c, _s())


def bnrelu_pool_fin(y, fin, a, pooled, idx):
    import ctypes
    _chk(y, bf16, "y"); _chk(a, bf16, "a"); _chk(pooled, bf16, "pooled")
    n, h, w, c = y.shape
    _lib.call("unetk_bnrelu_pool_fin", _p(y), ctypes.byref(fin), _p(a), _p(pooled), _p(idx), n, h, w, c, _s())


def bnrelu_pool(y, scale, shift, a, pooled, idx, apply=True):
    _chk(y, bf16, "y"); _chk(pooled, bf16, "pooled")
    n, h, w, c = y.shape
    _lib.call("unetk_bnrelu_pool", _p(y), _p(scale), _p(shift), _p(a), _p(pooled), _p(idx), n, h, w, c,
              int(apply), _s())


def maxpool_bwd(dpooled, idx, da, accumulate, bn=None, extra=None):
    _chk(dpooled, bf16, "dpooled"); _chk(da, bf16, "da")
    n, h, w, c = da.shape
    if extra is not None:
        _chk(extra, bf16, "extra")
    _lib.call("unetk_maxpool_bwd", _p(dpooled), _p(idx), _p(da), n, h, w, c, int(accumulate), _p(extra), _bn(bn), _s())


def bn_bwd_reduce(da, y, scale, shift, mean, invstd, sums):
    c = y.shape[-1]
    _lib.call("unetk_bn_bwd_reduce", _p(da), _p(y), _p(scale), _p(shift), _p(mean), _p(invstd), _p(sums),
              y.numel() // c, c, _s())


def bn_bwd_apply(da, y, scale, shift, mean, invstd, sums, dy, dgamma, dbeta, param_sums=None):
    c = y.shape[-1]
    _lib.call("unetk_bn_bwd_apply", _p(da), _p(y), _p(scale), _p(shift), _p(mean), _p(invstd), _p(sums), _p(param_sums),
              _p(dy), _p(dgamma), _p(dbeta), y.numel() // c, c, _s())


# ---------------------------------------------------------------------------------------------
# head / losses / scoring
# ---------------------------------------------------------------------------------------------
def head_fwd(a, w, b, out, sigmoid):
    _chk(a, bf16, "a"); _chk(out, f32, "out")
    n, h, ww, c = a.shape
    _lib.call("unetk_head_fwd", _p(a), _p(w), _p(b), _p(out), n, h, ww, c, out.shape[1], int(sigmoid), _s())
    return out


def head_fwd_bn(y, fin, w, b, out, sigmoid):
    """out = act(conv1x1(relu(BatchNorm_train(y)))) with the statistics finalize folded in (see bn_finalize_args);
    the activated tensor is not stored."""
    import ctypes
    _chk(y, bf16, "y"); _chk(out, f32, "out")
    n, h, ww, c = y.shape
    _lib.call("unetk_head_fwd_bn", _p(y), ctypes.byref(fin), _p(w), _p(b), _p(out), n, h, ww, c, out.shape[1], int(sigmoid), _s())
    return out


def head_loss_workspace(device):
    return torch.zeros(_lib.get().unetk_head_loss_workspace_bytes(), dtype=torch.uint8, device=device)


def head_loss_bwd(y, fin, w, b, target, out, da, dw, db, bn_sums, work, out3, which, rw, sw, alpha, gamma):
    """One decoder's output stage of a training step in one pass (include/unetk.h: unetk_head_loss_bwd)."""
    import ctypes
    _chk(y, bf16, "y"); _chk(target, f32, "target"); _chk(da, bf16, "da"); _chk(out3, f32, "out3")
    n, h, ww, c = y.shape
    n_out = w.shape[0]
    if tuple(target.shape) != (n, n_out, h, ww):
        raise ValueError(f"head_loss_bwd: target must be {(n, n_out, h, ww)}, got {tuple(target.shape)}")
    if out is not None:
        _chk(out, f32, "out")
    _lib.call("unetk_head_loss_bwd", _p(y), ctypes.byref(fin), _p(w), _p(b), _p(target), _p(out), _p(da), _p(dw), _p(db),
              _p(bn_sums), _p(work), work.numel(), _p(out3), int(which), n, h, ww, c, n_out, float(rw), float(sw),
              float(alpha), float(gamma), _s())


def head_bwd(a, w, dout, out, da, dw, db, sigmoid, bn=None):
    """a=None: the activation is recomputed from the BatchNorm context `bn` (forward used head_fwd_bn)."""
    _chk(dout, f32, "dout"); _chk(da, bf16, "da")
    if a is not None:
        _chk(a, bf16, "a")
    elif bn is None:
        raise _lib.UnetkError("head_bwd: without the stored activation a BatchNorm context is required")
    n, h, ww, c = da.shape
    _lib.call("unetk_head_bwd", _p(a), _p(w), _p(dout), _p(out), _p(da), _p(dw), _p(db), n, h, ww, c,
              dout.shape[1], int(sigmoid), _bn(bn), _s())


def channel_scale(x, mask):
    """x[n, h, w, c] *= mask[n, c] in place (Dropout2d on an NHWC bf16 activation)."""
    _chk(x, bf16, "x"); _chk(mask, f32, "mask")
    n, h, w, c = x.shape
    if mask.numel() != n * c:
        raise _lib.UnetkError(f"mask: expected {n * c} elements, got {mask.numel()}")
    _lib.call("unetk_channel_scale", _p(x), _p(mask), n, h, w, c, _s())
    return x


def seg_loss_workspace(n, c, device):
    return torch.empty(_lib.get().unetk_seg_loss_workspace_bytes(n, c) // 4, dtype=f32, device=device)


def seg_loss_fwd(logits, target, class_w, out4, work, ignore_index, ce_w, dice_w, focal_w, alpha, gamma, is_prob=False):
    _chk(logits, f32, "logits"); _chk(target, torch.int64, "target"); _chk(out4, f32, "out4")
    n, c, h, w = logits.shape
    _lib.call("unetk_seg_loss_fwd", _p(logits), _p(target), _p(class_w), _p(out4), _p(work), work.numel() * 4, n, c, h, w,
              int(ignore_index), float(ce_w), float(dice_w), float(focal_w), float(alpha), float(gamma), int(is_prob), _s())


def seg_loss_bwd(logits, target, class_w, work, gout, dlogits, ignore_index, ce_w, dice_w, focal_w, alpha, gamma, is_prob=False):
    _chk(logits, f32, "logits"); _chk(target, torch.int64, "target"); _chk(dlogits, f32, "dlogits")
    n, c, h, w = logits.shape
    _lib.call("unetk_seg_loss_bwd", _p(logits), _p(target), _p(class_w), _p(work), _p(gout), _p(dlogits), n, c, h, w,
              int(ignore_index), float(ce_w), float(dice_w), float(focal_w), float(alpha), float(gamma), int(is_prob), _s())


def upsample2x_fwd(x, out):
    """out[n, 2h, 2w, c] = bilinear x2 (align_corners=True) of x[n, h, w, c], NHWC bf16."""
    _chk(x, bf16, "x"); _chk(out, bf16, "out")
    n, h, w, c = x.shape
    if tuple(out.shape) != (n, 2 * h, 2 * w, c):
        raise _lib.UnetkError(f"out: expected shape {(n, 2 * h, 2 * w, c)}, got {tuple(out.shape)}")
    _lib.call("unetk_upsample2x_fwd", _p(x), _p(out), n, h, w, c, _s())
    return out


def upsample2x_bwd(dout, dx, accumulate=False):
    _chk(dout, bf16, "dout"); _chk(dx, bf16, "dx")
    n, h, w, c = dx.shape
    if tuple(dout.shape) != (n, 2 * h, 2 * w, c):
        raise _lib.UnetkError(f"dout: expected shape {(n, 2 * h, 2 * w, c)}, got {tuple(dout.shape)}")
    _lib.call("unetk_upsample2x_bwd", _p(dout), _p(dx), n, h, w, c, int(accumulate), _s())


def argmax_channels(logits, out):
    _chk(logits, f32, "logits"); _chk(out, torch.int64, "out")
    n, c, h, w = logits.shape
    _lib.call("unetk_argmax_channels", _p(logits), _p(out), n, c, h, w, _s())
    return out


def combined_loss_workspace(device):
    nbytes = _lib.get().unetk_combined_loss_workspace_bytes()
    return torch.zeros(nbytes, dtype=torch.uint8, device=device)


def combined_loss(recon, amap, image, mask, d_recon, d_amap, out3, work, rw, sw, alpha, gamma):
    for t, nm in ((recon, "reconstruction"), (amap, "anomaly_map"), (image, "original_image"), (mask, "true_mask")):
        _chk(t, f32, nm)
    _lib.call("unetk_combined_loss", _p(recon), _p(amap), _p(image), _p(mask), _p(d_recon), _p(d_amap),
              _p(out3), _p(work), work.numel(), recon.numel(), amap.numel(), rw, sw, alpha, gamma, _s())


def scale_by_scalar(y, scalar):
    _lib.call("unetk_scale_by_scalar", _p(y), _p(scalar), y.numel(), _s())


def ssim_fwd(img1, img2, sums, maps, window_size, need2):
    _chk(img1, f32, "img1"); _chk(img2, f32, "img2")
    n, c, h, w = img1.shape
    _lib.call("unetk_ssim_fwd", _p(img1), _p(img2), _p(sums), _p(maps), n, c, h, w, window_size, int(need2), _s())


def ssim_finalize(sums, out, shape, size_average):
    n, c, h, w = shape
    _lib.call("unetk_ssim_finalize", _p(sums), _p(out), n, c, h, w, int(size_average), _s())


def ssim_bwd(img1, img2, maps, gout, grad, window_size, which, per_image):
    n, c, h, w = img1.shape
    _lib.call("unetk_ssim_bwd", _p(img1), _p(img2), _p(maps), _p(gout), _p(grad), n, c, h, w, window_size,
              which, int(per_image), _s())


def anomaly_score(recon, orig, out, method=0):
    _chk(recon, f32, "reconstruction"); _chk(orig, f32, "original")
    n, c, h, w = recon.shape
    _lib.call("unetk_anomaly_score", _p(recon), _p(orig), _p(out), n, c, h, w, method, _s())
    return out


def pixel_confusion(pred, truth, include, thresholds, counts, truth_threshold=0.5):
    """counts[t] += (tn, fp, fn, tp) of (truth > truth_threshold, pred > thresholds[t]) over the images with
    include[i] != 0 (all images when include is None).  pred / truth: fp32 [n, ...] of equal shape."""
    import ctypes
    _chk(pred, f32, "pred"); _chk(truth, f32, "truth"); _chk(counts, torch.int64, "counts")
    if pred.shape != truth.shape:
        raise ValueError(f"pixel_confusion: shapes differ: {tuple(pred.shape)} vs {tuple(truth.shape)}")
    n = pred.shape[0] if pred.dim() > 0 else 0
    per = pred.numel() // n if n else 0
    nt = len(thresholds)
    if counts.numel() != nt * 4:
        raise ValueError(f"pixel_confusion: counts must hold {nt} x 4 values")
    if include is not None:
        _chk(include, torch.uint8, "include")
        if include.numel() != n:
            raise ValueError("pixel_confusion: one include flag per image")
    thr = (ctypes.c_float * max(nt, 1))(*[float(t) for t in thresholds])
    _lib.call("unetk_pixel_confusion", _p(pred), _p(truth), _p(include) if include is not None else None,
              ctypes.cast(thr, ctypes.c_void_p), nt, float(truth_threshold), _p(counts), n, per, _s())
    return counts


def class_confusion(pred, target, num_classes, ignore_index, cm, kept):
    """cm[t, p] += 1 / kept += pixels (see include/unetk.h); pred: NCHW fp32 logits or int64 [N, ...] indices."""
    _chk(target, torch.int64, "target"); _chk(cm, torch.int64, "cm"); _chk(kept, torch.int64, "kept")
    if cm.numel() != num_classes * num_classes or kept.numel() != 1:
        raise ValueError("class_confusion: cm must hold num_classes^2 values, kept one")
    n = target.shape[0] if target.dim() > 0 else 0
    per = target.numel() // n if n else 0
    has_ignore = ignore_index is not None
    ii = int(ignore_index) if has_ignore else 0
    if pred.dtype == f32:
        _chk(pred, f32, "pred")
        if pred.dim() != target.dim() + 1 or pred.shape[0] != n or pred.shape[1] != num_classes \
                or pred.numel() != n * num_classes * per:
            raise ValueError(f"class_confusion: logits {tuple(pred.shape)} do not match target {tuple(target.shape)} "
                             f"with {num_classes} classes")
        _lib.call("unetk_class_confusion", _p(pred), None, _p(target), n, num_classes, per, ii, int(has_ignore),
                  _p(cm), _p(kept), _s())
    else:
        _chk(pred, torch.int64, "pred")
        if pred.numel() != target.numel():
            raise ValueError("class_confusion: predictions and target differ in size")
        _lib.call("unetk_class_confusion", None, _p(pred), _p(target), n, num_classes, per, ii, int(has_ignore),
                  _p(cm), _p(kept), _s())
    return cm


def select_kth(x, k, out2=None, work=None):
    """(k-th, (k+1)-th) smallest values of the fp32 tensor x (0-based rank, NaN last) -> fp32 [2] on the device."""
    _chk(x, f32, "x")
    nbytes = _lib.get().unetk_select_workspace_bytes()
    if work is None:
        work = torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=x.device)
    if out2 is None:
        out2 = torch.empty(2, dtype=f32, device=x.device)
    _lib.call("unetk_select_kth", _p(x), x.numel(), int(k), _p(out2), _p(work), work.numel() * 8, _s())
    return out2


def resize_coeffs(in_size, out_size):
    """HOST: Pillow's bilinear tap table of one axis -> (bounds int32 [out, 2], coeffs int32 [out, ksize])."""
    import numpy as np
    lib = _lib.get()
    ksize = lib.unetk_resize_ksize(int(in_size), int(out_size))
    if ksize <= 0:
        raise ValueError(f"resize_coeffs: bad sizes {in_size} -> {out_size}")
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    coeffs = np.zeros((out_size, ksize), dtype=np.int32)
    _lib.call_host("unetk_resize_coeffs", int(in_size), int(out_size), bounds.ctypes.data, coeffs.ctypes.data)
    return bounds, coeffs


def resize_u8(src, tmp, out_u8, out_f32, hplan, vplan, mean=None, std=None):
    """src u8 [n, h, w, c] -> out_u8 [n, oh, ow, c] / out_f32 [n, c, oh, ow]; hplan / vplan: device (bounds, coeffs)."""
    import ctypes
    u8 = torch.uint8
    _chk(src, u8, "src"); _chk(tmp, u8, "tmp")
    n, h, w, c = src.shape
    (hb, hk), (vb, vk) = hplan, vplan
    for t, nm in ((hb, "hbounds"), (hk, "hcoeffs"), (vb, "vbounds"), (vk, "vcoeffs")):
        _chk(t, torch.int32, nm)
    ow, oh = hb.shape[0], vb.shape[0]
    if tmp.numel() < n * h * ow * c:
        raise ValueError("resize_u8: tmp too small")
    if out_u8 is not None:
        _chk(out_u8, u8, "out_u8")
        if tuple(out_u8.shape) != (n, oh, ow, c):
            raise ValueError(f"resize_u8: out_u8 must be {(n, oh, ow, c)}")
    if out_f32 is not None:
        _chk(out_f32, f32, "out_f32")
        if tuple(out_f32.shape) != (n, c, oh, ow):
            raise ValueError(f"resize_u8: out_f32 must be {(n, c, oh, ow)}")
    m = (ctypes.c_float * c)(*[float(v) for v in mean]) if mean is not None else None
    sd = (ctypes.c_float * c)(*[float(v) for v in std]) if std is not None else None
    _lib.call("unetk_resize_u8", _p(src), _p(tmp), _p(out_u8), _p(out_f32), n, h, w, c, oh, ow, _p(hb), _p(hk),
              hk.shape[1], _p(vb), _p(vk), vk.shape[1],
              ctypes.cast(m, ctypes.c_void_p) if m is not None else None,
              ctypes.cast(sd, ctypes.c_void_p) if sd is not None else None, _s())
    return out_f32 if out_f32 is not None else out_u8


def resize_nearest_table(in_size, out_size):
    """HOST: Pillow's nearest-neighbour source index per output index of one axis (int32 [out])."""
    import numpy as np
    table = np.zeros(out_size, dtype=np.int32)
    _lib.call_host("unetk_resize_nearest_table", int(in_size), int(out_size), table.ctypes.data)
    return table


def resize_nearest_u8(src, out_i64, out_u8, xtab, ytab):
    """u8 [n, h, w] -> int64 / u8 [n, oh, ow] through the device copies of the two index tables."""
    _chk(src, torch.uint8, "src"); _chk(xtab, torch.int32, "xtab"); _chk(ytab, torch.int32, "ytab")
    n, h, w = src.shape
    oh, ow = ytab.numel(), xtab.numel()
    for t, dt, nm in ((out_i64, torch.int64, "out_i64"), (out_u8, torch.uint8, "out_u8")):
        if t is not None:
            _chk(t, dt, nm)
            if tuple(t.shape) != (n, oh, ow):
                raise ValueError(f"resize_nearest_u8: {nm} must be {(n, oh, ow)}")
    _lib.call("unetk_resize_nearest_u8", _p(src), _p(out_i64), _p(out_u8), n, h, w, oh, ow, _p(xtab), _p(ytab), _s())
    return out_i64 if out_i64 is not None else out_u8


def flip_rotate_u8(src, dst, geom):
    """dst = rotate(flip(src)) per image; geom int32 [n, 8] on the device (see include/unetk.h)."""
    _chk(src, torch.uint8, "src"); _chk(dst, torch.uint8, "dst"); _chk(geom, torch.int32, "geom")
    n, h, w, c = src.shape
    if dst.shape != src.shape or geom.numel() != n * 8:
        raise ValueError("flip_rotate_u8: dst must have src's shape and geom 8 ints per image")
    _lib.call("unetk_flip_rotate_u8", _p(src), _p(dst), n, h, w, c, _p(geom), _s())
    return dst


def color_jitter_u8(img, stage_ops, factors, work=None):
    """ColorJitter's four adjustments in place on u8 [n, h, w, 3]; stage_ops int32 [n, 4], factors fp32 [n, 4]."""
    _chk(img, torch.uint8, "img"); _chk(stage_ops, torch.int32, "ops"); _chk(factors, f32, "factors")
    n, h, w, c = img.shape
    if c != 3 or stage_ops.numel() != n * 4 or factors.numel() != n * 4:
        raise ValueError("color_jitter_u8: RGB images, 4 ops and 4 factors per image")
    if work is None:
        work = torch.empty(4 * n, dtype=torch.int64, device=img.device)
    _lib.call("unetk_color_jitter_u8", _p(img), n, h, w, _p(stage_ops), _p(factors), _p(work), _s())
    return img


def u8_to_tensor(hwc, out, mean=None, std=None):
    """ToTensor + Normalize: u8 [n, h, w, c] -> fp32 [n, c, h, w]."""
    import ctypes
    _chk(hwc, torch.uint8, "hwc"); _chk(out, f32, "out")
    n, h, w, c = hwc.shape
    if tuple(out.shape) != (n, c, h, w):
        raise ValueError(f"u8_to_tensor: out must be {(n, c, h, w)}")
    m = (ctypes.c_float * c)(*[float(v) for v in mean]) if mean is not None else None
    sd = (ctypes.c_float * c)(*[float(v) for v in std]) if std is not None else None
    _lib.call("unetk_u8_to_tensor", _p(hwc), _p(out), n, h, w, c,
              ctypes.cast(m, ctypes.c_void_p) if m is not None else None,
              ctypes.cast(sd, ctypes.c_void_p) if sd is not None else None, _s())
    return out


# ---------------------------------------------------------------------------------------------
# optimizer / layout
# ---------------------------------------------------------------------------------------------
def adam_step(p, g, m, v, lr, beta1, beta2, eps, wd, step, grad_scale=1.0, mode=0):
    _lib.call("unetk_adam_step", _p(p), _p(g), _p(m), _p(v), p.numel(), lr, beta1, beta2, eps, wd, step,
              grad_scale, mode, _s())


def permute_batched(table, src, dst, scale=1.0):
    """table: packing.PermuteTable (device descriptor array)."""
    if table.n_desc == 0:          # e.g. bilinear decoders: no transposed-conv weights to re-lay-out
        return
    _lib.call("unetk_permute_batched", _p(table.device_descs), table.n_desc, table.n_chunks, _p(src), _p(dst),
              int(dst.dtype == bf16), scale, _s())


def pack_conv_weights(table, flat_p, wpack):
    if table.n_desc == 0:
        return
    _lib.call("unetk_pack_conv_weights", _p(table.device_descs), table.n_desc, table.n_blocks, _p(flat_p), _p(wpack), _s())


def unpack_conv_grads(table, gstage, flat_g, scale=1.0):
    _lib.call("unetk_unpack_conv_grads", _p(table.device_descs), table.n_desc, table.n_blocks, table.max_cin,
              _p(gstage), _p(flat_g), scale, _s())


def fused_update(table, p, g, m, v, gstage, wpack, hyper, mode, max_blocks=0):
    """table: packing.UpdateTable.  Staged gradients -> g, optimizer update of p / m / v, new weights -> wpack, zeroes
    over the staging ranges read; hyper: device float32[>= 9] (see include/unetk.h: unetk_fused_update)."""
    if table.n_desc == 0:
        return
    for t, name in ((p, "p"), (g, "g"), (m, "m"), (gstage, "gstage"), (hyper, "hyper")):
        _chk(t, f32, name)
    if v is not None:
        _chk(v, f32, "v")
    _chk(wpack, bf16, "wpack")
    if not (p.numel() == g.numel() == m.numel() == gstage.numel() and p.numel() >= table.max_end) or hyper.numel() < 9:
        raise _lib.UnetkError("fused_update: flat buffers of different sizes / shorter than the table's tensors")
    _lib.call("unetk_fused_update", _p(table.device_descs), table.n_desc, table.n_blocks, max_blocks, _p(p), _p(g), _p(m), _p(v),
              _p(gstage), _p(wpack), _p(hyper), mode, _s())


def set_floats(dst, values):
    """dst[:len(values)] = values (at most 16 floats, passed by value inside the launch: no host buffer to keep alive)."""
    import ctypes
    _chk(dst, f32, "dst")
    n = len(values)
    if not 0 < n <= 16 or dst.numel() < n:
        raise _lib.UnetkError("set_floats: 1..16 values into a buffer that holds them")
    arr = (ctypes.c_float * n)(*[float(x) for x in values])
    _lib.call("unetk_set_floats", _p(dst), ctypes.cast(arr, ctypes.c_void_p), n, _s())


def nchw_to_nhwc(x, out):
    _chk(x, f32, "x"); _chk(out, bf16, "out")
    n, c, h, w = x.shape
    _lib.call("unetk_nchw_to_nhwc", _p(x), _p(out), n, c, h, w, _s())
    return out


def nhwc_to_nchw(a, out):
    _chk(a, bf16, "a"); _chk(out, f32, "out")
    n, h, w, c = a.shape
    _lib.call("unetk_nhwc_to_nchw", _p(a), _p(out), n, c, h, w, _s())
    return out


# ---------------------------------------------------------------------------------------------
# device guard
# ---------------------------------------------------------------------------------------------
def _on_tensor_device(fn):
    """Run `fn` with its first tensor argument's device current: _s() and the library's own device checks look
    at the current device, so a model on cuda:1 under a caller whose current device is cuda:0 would otherwise
    launch on device 0's stream with device-1 pointers."""
    import functools

    @functools.wraps(fn)
    def guarded(*args, **kwargs):
        for a in args:
            if isinstance(a, torch.Tensor):
                if a.is_cuda and a.device.index != torch.cuda.current_device():
                    with torch.cuda.device(a.device):
                        return fn(*args, **kwargs)
                break
        return fn(*args, **kwargs)

    return guarded


for _name, _fn in list(globals().items()):
    if callable(_fn) and not _name.startswith("_") and getattr(_fn, "__module__", None) == __name__ \
            and _name not in ("timing_enabled", "enable_conv_timing", "collect_conv_timing", "conv_timing_by_kernel",
                              "resize_coeffs", "resize_nearest_table", "bn_ctx", "bn_finalize_args"):
        globals()[_name] = _on_tensor_device(_fn)
del _name, _fn
