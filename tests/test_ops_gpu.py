"""Operator-level parity of the CUDA kernels (through the C ABI) against plain fp32 PyTorch ops
of the same operator on the same seeded inputs.  Inputs are rounded to bf16 first so that the only
differences are accumulation order and the bf16 rounding of the stored result.

Tolerances (stated per test): conv-type outputs are stored in bf16 -> relative 2^-8 per element;
we allow |d| <= 2e-2 * max|ref| (K up to 9216 fp32 accumulation, bf16 store)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _mods():
    from tiaozhanbei_unet_b200 import _lib, ops, packing
    return _lib, ops, packing


def _pack(ops, packing, w, recipe, device, dtype=torch.bfloat16):
    _lib = ops._lib
    tab = packing.PermuteTable(_lib.get().unetk_permute_chunk_elems())
    dims, strides, flip = recipe
    tab.add(0, 0, dims, strides, flip)
    tab.finalize(device)
    out = torch.empty(dims[0] * dims[1] * dims[2], dtype=dtype, device=device)
    ops.permute_batched(tab, w.contiguous().view(-1), out)
    return out


def _nhwc(x):  # NCHW fp32 -> NHWC bf16
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def _nchw(a):  # NHWC bf16 -> NCHW fp32
    return a.float().permute(0, 3, 1, 2).contiguous()


def _close(got, ref, tol, what):
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-12
    assert err <= tol * scale, f"{what}: max abs err {err:.4e} vs scale {scale:.4e} (tol {tol})"


CONV_CASES = [
    # n, h, w, c1, c2, cout
    (2, 16, 16, 64, 0, 64),
    (2, 16, 16, 128, 0, 128),
    (1, 32, 32, 64, 64, 64),      # dual source (decoder concat)
    (2, 16, 24, 128, 0, 256),     # partial tiles in x, two n-blocks
    (1, 8, 8, 256, 256, 512),
    (8, 64, 64, 64, 0, 128),      # more tiles than SMs: persistent loop + double-buffered TMEM
    (3, 4, 4, 64, 0, 64),         # tile spans several images
    (1, 32, 88, 64, 0, 64),       # KolektorSDD-like width (not a power of two)
]


@pytest.mark.parametrize("n,h,w,c1,c2,cout", CONV_CASES)
def test_conv3x3_fwd(cuda_device, n, h, w, c1, c2, cout):
    _lib, ops, packing = _mods()
    g = torch.Generator(device="cpu").manual_seed(1234)
    cin = c1 + c2
    x = torch.randn(n, cin, h, w, generator=g).to(cuda_device)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / (3.0 * cin ** 0.5)).to(cuda_device)
    xb = _nhwc(x)
    wq = wt.to(torch.bfloat16).float()
    ref = F.conv2d(_nchw(xb), wq, padding=1)
    wf = _pack(ops, packing, wt, packing.conv_fwd_pack(cout, cin), cuda_device)
    y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device=cuda_device)
    stats = torch.zeros(2 * cout, dtype=torch.float32, device=cuda_device)
    s1 = xb[..., :c1].contiguous()
    s2 = xb[..., c1:].contiguous() if c2 else None
    ops.conv3x3_fwd(s1, s2, wf, y, stats=stats)
    torch.cuda.synchronize()
    _close(_nchw(y), ref, 2e-2, "conv3x3_fwd")
    # BatchNorm statistics of the fp32 accumulators
    _close(stats[:cout], ref.sum(dim=(0, 2, 3)), 2e-3 * (n * h * w) ** 0.5, "sum")
    _close(stats[cout:], (ref * ref).sum(dim=(0, 2, 3)), 5e-3, "sumsq")


def test_conv3x3_fwd_affine_relu(cuda_device):
    _lib, ops, packing = _mods()
    g = torch.Generator(device="cpu").manual_seed(7)
    n, h, w, cin, cout = 2, 16, 16, 64, 128
    x = torch.randn(n, cin, h, w, generator=g).to(cuda_device)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / 24.0).to(cuda_device)
    scale = (torch.rand(cout, generator=g) + 0.5).to(cuda_device)
    shift = torch.randn(cout, generator=g).to(cuda_device) * 0.1
    xb = _nhwc(x)
    ref = F.relu(F.conv2d(_nchw(xb), wt.to(torch.bfloat16).float(), padding=1) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1))
    wf = _pack(ops, packing, wt, packing.conv_fwd_pack(cout, cin), cuda_device)
    y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device=cuda_device)
    ops.conv3x3_fwd(xb, None, wf, y, scale=scale, shift=shift, relu=True)
    torch.cuda.synchronize()
    _close(_nchw(y), ref, 2e-2, "conv3x3_fwd affine")


@pytest.mark.parametrize("n,h,w,c1,c2,cout", CONV_CASES)
def test_conv3x3_dgrad(cuda_device, n, h, w, c1, c2, cout):
    _lib, ops, packing = _mods()
    g = torch.Generator(device="cpu").manual_seed(99)
    cin = c1 + c2
    dy = torch.randn(n, cout, h, w, generator=g).to(cuda_device)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / (3.0 * cout ** 0.5)).to(cuda_device)
    dyb = _nhwc(dy)
    wq = wt.to(torch.bfloat16).float()
    ref = F.conv_transpose2d(_nchw(dyb), wq, padding=1)   # = dgrad of conv2d(pad=1)
    wd = _pack(ops, packing, wt, packing.conv_dgrad_pack(cout, cin), cuda_device)
    prev1 = torch.randn(n, h, w, c1, generator=g).to(cuda_device).to(torch.bfloat16)
    dx1 = prev1.clone()
    dx2 = torch.empty(n, h, w, c2, dtype=torch.bfloat16, device=cuda_device) if c2 else None
    stats = torch.zeros(c2, dtype=torch.float32, device=cuda_device) if c2 else None
    ops.conv3x3_dgrad(dyb, wd, dx1, acc1=True, dx2=dx2, acc2=False, stats=stats)
    torch.cuda.synchronize()
    _close(_nchw(dx1), ref[:, :c1] + _nchw(prev1), 2e-2, "dgrad dx1 (accumulate)")
    if c2:
        _close(_nchw(dx2), ref[:, c1:], 2e-2, "dgrad dx2")
        _close(stats, ref[:, c1:].sum(dim=(0, 2, 3)), 2e-3 * (n * h * w) ** 0.5, "dgrad column sums")
    else:
        # fused first pass of the BatchNorm backward of the producing layer (single destination)
        yb = _nhwc(torch.randn(n, c1, h, w, generator=g).to(cuda_device))
        sc = (torch.rand(c1, generator=g) + 0.5).to(cuda_device); sh = torch.randn(c1, generator=g).to(cuda_device) * 0.3
        mu = torch.randn(c1, generator=g).to(cuda_device) * 0.2; iv = (torch.rand(c1, generator=g) + 0.5).to(cuda_device)
        fused = torch.zeros(2 * c1, device=cuda_device); ref_sums = torch.zeros(2 * c1, device=cuda_device)
        dx3 = prev1.clone()
        ops.conv3x3_dgrad(dyb, wd, dx3, acc1=True, bn=ops.bn_ctx(yb, sc, sh, mu, iv, fused))
        ops.bn_bwd_reduce(dx3, yb, sc, sh, mu, iv, ref_sums)
        torch.cuda.synchronize()
        assert torch.equal(dx3, dx1)
        _close(fused, ref_sums, 1e-4, "dgrad fused BN reduction")


@pytest.mark.parametrize("n,h,w,c1,c2,cout", CONV_CASES)
def test_conv3x3_wgrad(cuda_device, n, h, w, c1, c2, cout):
    _lib, ops, packing = _mods()
    g = torch.Generator(device="cpu").manual_seed(5)
    cin = c1 + c2
    x = torch.randn(n, cin, h, w, generator=g).to(cuda_device)
    dy = torch.randn(n, cout, h, w, generator=g).to(cuda_device)
    xb, dyb = _nhwc(x), _nhwc(dy)
    ref = torch.nn.grad.conv2d_weight(_nchw(xb), (cout, cin, 3, 3), _nchw(dyb), padding=1)
    dwk = torch.zeros(cout * 9 * cin, dtype=torch.float32, device=cuda_device)
    s1 = xb[..., :c1].contiguous()
    s2 = xb[..., c1:].contiguous() if c2 else None
    ops.conv3x3_wgrad(dyb, s1, s2, dwk)
    got = _pack(ops, packing, dwk, packing.conv_grad_unpack(cout, cin), cuda_device, dtype=torch.float32).view(cout, cin, 3, 3)
    torch.cuda.synchronize()
    _close(got, ref, 2e-3, "conv3x3_wgrad")


CONVT_CASES = [(2, 8, 8, 128, 64), (1, 16, 16, 256, 128), (2, 4, 4, 1024, 512), (1, 2, 2, 128, 64), (1, 32, 44, 128, 64)]


@pytest.mark.parametrize("n,h,w,cin,cout", CONVT_CASES)
def test_convt2x2(cuda_device, n, h, w, cin, cout):
    _lib, ops, packing = _mods()
    g = torch.Generator(device="cpu").manual_seed(11)
    a = torch.randn(n, cin, h, w, generator=g).to(cuda_device)
    wt = (torch.randn(cin, cout, 2, 2, generator=g) / cin ** 0.5).to(cuda_device)
    bias = torch.randn(cout, generator=g).to(cuda_device)
    ab = _nhwc(a)
    wq = wt.to(torch.bfloat16).float()
    ref = F.conv_transpose2d(_nchw(ab), wq, bias, stride=2)
    wp = _pack(ops, packing, wt, packing.convt_fwd_pack(cin, cout), cuda_device)
    up = torch.empty(n, 2 * h, 2 * w, cout, dtype=torch.bfloat16, device=cuda_device)
    ops.convt2x2_fwd(ab, wp, bias, up)
    torch.cuda.synchronize()
    _close(_nchw(up), ref, 2e-2, "convt2x2_fwd")
    # dgrad
    dup = torch.randn(n, cout, 2 * h, 2 * w, generator=g).to(cuda_device)
    dupb = _nhwc(dup)
    ref_da = F.conv2d(_nchw(dupb), wq, stride=2)
    wpd = _pack(ops, packing, wt, packing.convt_dgrad_pack(cin, cout), cuda_device)
    da = torch.empty(n, h, w, cin, dtype=torch.bfloat16, device=cuda_device)
    ops.convt2x2_dgrad(dupb, wpd, da, acc=False)
    torch.cuda.synchronize()
    _close(_nchw(da), ref_da, 2e-2, "convt2x2_dgrad")
    yb = _nhwc(torch.randn(n, cin, h, w, generator=g).to(cuda_device))
    sc = (torch.rand(cin, generator=g) + 0.5).to(cuda_device); sh = torch.randn(cin, generator=g).to(cuda_device) * 0.3
    mu = torch.randn(cin, generator=g).to(cuda_device) * 0.2; iv = (torch.rand(cin, generator=g) + 0.5).to(cuda_device)
    fused = torch.zeros(2 * cin, device=cuda_device); ref_sums = torch.zeros(2 * cin, device=cuda_device)
    da2 = torch.empty_like(da)
    ops.convt2x2_dgrad(dupb, wpd, da2, acc=False, bn=ops.bn_ctx(yb, sc, sh, mu, iv, fused))
    ops.bn_bwd_reduce(da2, yb, sc, sh, mu, iv, ref_sums)
    torch.cuda.synchronize()
    assert torch.equal(da2, da)
    _close(fused, ref_sums, 1e-4, "convt2x2_dgrad fused BN reduction")
    # wgrad: d/dW of conv_transpose2d == conv2d_weight with roles swapped
    ref_dw = torch.nn.grad.conv2d_weight(_nchw(dupb), (cin, cout, 2, 2), _nchw(ab), stride=2)
    dwk = torch.zeros(cout * 4 * cin, dtype=torch.float32, device=cuda_device)
    ops.convt2x2_wgrad(ab, dupb, dwk)
    got = _pack(ops, packing, dwk, packing.convt_grad_unpack(cin, cout), cuda_device, dtype=torch.float32).view(cin, cout, 2, 2)
    torch.cuda.synchronize()
    _close(got, ref_dw, 2e-3, "convt2x2_wgrad")


@pytest.mark.parametrize("n,cin,h,w", [(2, 3, 16, 16), (1, 3, 40, 24), (2, 1, 8, 8)])
def test_first_conv(cuda_device, n, cin, h, w):
    _lib, ops, packing = _mods()
    g = torch.Generator(device="cpu").manual_seed(3)
    x = torch.randn(n, cin, h, w, generator=g).to(cuda_device)
    wt = (torch.randn(64, cin, 3, 3, generator=g) / 5.0).to(cuda_device)
    ref = F.conv2d(x, wt, padding=1)
    y = torch.empty(n, h, w, 64, dtype=torch.bfloat16, device=cuda_device)
    stats = torch.zeros(128, dtype=torch.float32, device=cuda_device)
    ops.first_conv_fwd(x, wt, y, stats)
    torch.cuda.synchronize()
    _close(_nchw(y), ref, 1e-2, "first_conv_fwd")
    _close(stats[:64], ref.sum(dim=(0, 2, 3)), 1e-3 * (n * h * w) ** 0.5, "first conv sum")
    _close(stats[64:], (ref * ref).sum(dim=(0, 2, 3)), 1e-3, "first conv sumsq")
    dy = torch.randn(n, 64, h, w, generator=g).to(cuda_device)
    dyb = _nhwc(dy)
    ref_dw = torch.nn.grad.conv2d_weight(x, (64, cin, 3, 3), _nchw(dyb), padding=1)
    dw = torch.empty(64, cin, 3, 3, dtype=torch.float32, device=cuda_device)
    ops.first_conv_wgrad(x, dyb, dw)
    torch.cuda.synchronize()
    _close(dw, ref_dw, 1e-4, "first_conv_wgrad")
    # tensor-core route: channel-padded bf16 image + the 64->64 nine-tap weight-gradient kernel
    x64 = torch.zeros(n, h, w, 64, dtype=torch.bfloat16, device=cuda_device)
    ops.image_to_nhwc64(x, x64)
    torch.cuda.synchronize()
    assert torch.equal(x64[..., :cin].float(), x.to(torch.bfloat16).float().permute(0, 2, 3, 1))
    assert float(x64[..., cin:].abs().max()) == 0.0
    ref_dw_bf = torch.nn.grad.conv2d_weight(x.to(torch.bfloat16).float(), (64, cin, 3, 3), _nchw(dyb), padding=1)
    scratch = torch.empty(64 * 9 * 64, dtype=torch.float32, device=cuda_device)
    dw2 = torch.full((64, cin, 3, 3), 7.0, dtype=torch.float32, device=cuda_device)
    ops.first_conv_wgrad_tc(x64, dyb, scratch, dw2)
    torch.cuda.synchronize()
    _close(dw2, ref_dw_bf, 1e-4, "first_conv_wgrad_tc (bf16 image)")
    _close(dw2, ref_dw, 1e-2, "first_conv_wgrad_tc vs fp32 image")


@pytest.mark.parametrize("n,h,w,c", [(2, 8, 8, 64), (1, 16, 12, 128), (3, 4, 6, 1024)])
def test_bn_relu_pool_roundtrip(cuda_device, n, h, w, c):
    """bn_finalize + bnrelu_apply/pool forward and the two-pass backward against torch autograd of
    BatchNorm2d(train) -> ReLU -> MaxPool2d(2)."""
    _lib, ops, packing = _mods()
    g = torch.Generator(device="cpu").manual_seed(21)
    dev = cuda_device
    yb = _nhwc((torch.randn(n, c, h, w, generator=g) * 2 + 0.5).to(dev))
    gamma = (torch.rand(c, generator=g) + 0.5).to(dev)
    beta = (torch.randn(c, generator=g) * 0.2).to(dev)
    yref = _nchw(yb).requires_grad_(True)
    bn = torch.nn.BatchNorm2d(c).to(dev).train()
    with torch.no_grad():
        bn.weight.copy_(gamma); bn.bias.copy_(beta)
    a_ref = F.relu(bn(yref))
    p_ref, idx_ref = F.max_pool2d(a_ref, 2, return_indices=True)
    # ours
    stats = torch.stack([_nchw(yb).sum(dim=(0, 2, 3)), (_nchw(yb) ** 2).sum(dim=(0, 2, 3))]).contiguous().view(-1)
    rm = torch.zeros(c, device=dev); rv = torch.ones(c, device=dev); nbt = torch.zeros((), dtype=torch.int64, device=dev)
    scale = torch.empty(c, device=dev); shift = torch.empty(c, device=dev); mean = torch.empty(c, device=dev); invstd = torch.empty(c, device=dev)
    ops.bn_finalize(stats, gamma, beta, rm, rv, nbt, scale, shift, mean, invstd, count=n * h * w)
    a = torch.empty_like(yb)
    pooled = torch.empty(n, h // 2, w // 2, c, dtype=torch.bfloat16, device=dev)
    idx = torch.empty(n, h // 2, w // 2, c, dtype=torch.uint8, device=dev)
    ops.bnrelu_pool(yb, scale, shift, a, pooled, idx, apply=True)
    a2 = torch.empty_like(yb)
    ops.bnrelu_apply(yb, scale, shift, a2)
    torch.cuda.synchronize()
    assert int(nbt.item()) == 1
    _close(rm, bn.running_mean, 1e-5, "running_mean")
    _close(rv, bn.running_var, 1e-4, "running_var")
    _close(_nchw(a), a_ref.detach(), 1e-2, "bnrelu (pool variant)")
    assert torch.equal(a, a2)
    _close(_nchw(pooled), p_ref.detach(), 1e-2, "pooled")
    # the same with the finalize folded into the kernels: bit-identical outputs, constants and running stats
    for pool_variant in (False, True):
        rm2 = torch.zeros(c, device=dev); rv2 = torch.ones(c, device=dev); nbt2 = torch.zeros((), dtype=torch.int64, device=dev)
        sc2 = torch.empty(c, device=dev); sh2 = torch.empty(c, device=dev); mu2 = torch.empty(c, device=dev); iv2 = torch.empty(c, device=dev)
        fin = ops.bn_finalize_args(stats, gamma, beta, rm2, rv2, nbt2, sc2, sh2, mu2, iv2, n * h * w, 1e-5, 0.1)
        a3 = torch.empty_like(yb)
        if pool_variant:
            pooled3 = torch.empty_like(pooled); idx3 = torch.empty_like(idx)
            ops.bnrelu_pool_fin(yb, fin, a3, pooled3, idx3)
        else:
            ops.bnrelu_apply_fin(yb, fin, a3)
        torch.cuda.synchronize()
        assert torch.equal(a3, a) and int(nbt2.item()) == 1
        for got, want in ((sc2, scale), (sh2, shift), (mu2, mean), (iv2, invstd), (rm2, rm), (rv2, rv)):
            assert torch.equal(got, want)
        if pool_variant:
            assert torch.equal(pooled3, pooled) and torch.equal(idx3, idx)
    # backward: upstream gradient on the pooled tensor
    dp = torch.randn(n, c, h // 2, w // 2, generator=g).to(dev)
    dpb = _nhwc(dp)
    p_ref.backward(_nchw(dpb))
    da = torch.empty_like(yb)
    ops.maxpool_bwd(dpb, idx, da, accumulate=False)
    sums = torch.zeros(2 * c, device=dev)
    ops.bn_bwd_reduce(da, yb, scale, shift, mean, invstd, sums)
    dy = torch.empty_like(yb); dgamma = torch.empty(c, device=dev); dbeta = torch.empty(c, device=dev)
    ops.bn_bwd_apply(da, yb, scale, shift, mean, invstd, sums, dy, dgamma, dbeta)
    torch.cuda.synchronize()
    _close(dgamma, bn.weight.grad, 2e-2, "dgamma")
    _close(dbeta, bn.bias.grad, 2e-2, "dbeta")
    _close(_nchw(dy), yref.grad, 3e-2, "bn backward dy")
    # accumulate variant
    base = torch.randn(n, h, w, c, generator=g).to(dev).to(torch.bfloat16)
    da2 = base.clone()
    ops.maxpool_bwd(dpb, idx, da2, accumulate=True)
    torch.cuda.synchronize()
    _close(da2.float(), base.float() + da.float(), 1e-2, "maxpool_bwd accumulate")
    # a second gradient of the same activation added on the fly (parallel decoders)
    extra = _nhwc(torch.randn(n, c, h, w, generator=g).to(dev))
    da4 = base.clone()
    ops.maxpool_bwd(dpb, idx, da4, accumulate=True, extra=extra)
    torch.cuda.synchronize()
    _close(da4.float(), base.float() + da.float() + extra.float(), 1.5e-2, "maxpool_bwd accumulate + extra")
    # fused first pass of the BatchNorm backward == the stand-alone reduction on the finished gradient
    da3 = base.clone()
    fused = torch.zeros(2 * c, device=dev)
    ops.maxpool_bwd(dpb, idx, da3, accumulate=True, bn=ops.bn_ctx(yb, scale, shift, mean, invstd, fused))
    ref_sums = torch.zeros(2 * c, device=dev)
    ops.bn_bwd_reduce(da3, yb, scale, shift, mean, invstd, ref_sums)
    torch.cuda.synchronize()
    assert torch.equal(da3, da2)
    _close(fused, ref_sums, 1e-4, "maxpool_bwd fused BN reduction")


@pytest.mark.parametrize("n_out,sigmoid", [(3, True), (1, True), (1, False)])
def test_head(cuda_device, n_out, sigmoid):
    _lib, ops, packing = _mods()
    g = torch.Generator(device="cpu").manual_seed(31)
    dev = cuda_device
    n, h, w = 2, 12, 20
    a = _nhwc(torch.randn(n, 64, h, w, generator=g).to(dev))
    wt = (torch.randn(n_out, 64, generator=g) / 8).to(dev)
    b = torch.randn(n_out, generator=g).to(dev)
    aref = _nchw(a).requires_grad_(True)
    wref = wt.clone().view(n_out, 64, 1, 1).requires_grad_(True)
    bref = b.clone().requires_grad_(True)
    o_ref = F.conv2d(aref, wref, bref)
    if sigmoid:
        o_ref = torch.sigmoid(o_ref)
    out = torch.empty(n, n_out, h, w, device=dev)
    ops.head_fwd(a, wt, b, out, sigmoid)
    torch.cuda.synchronize()
    _close(out, o_ref.detach(), 1e-5, "head_fwd")
    dout = torch.randn(n, n_out, h, w, generator=g).to(dev)
    o_ref.backward(dout)
    da = torch.empty_like(a); dw = torch.empty(n_out, 64, device=dev); db = torch.empty(n_out, device=dev)
    ops.head_bwd(a, wt, dout, out, da, dw, db, sigmoid)
    torch.cuda.synchronize()
    _close(_nchw(da), aref.grad, 1e-2, "head_bwd da")
    _close(dw, wref.grad.view(n_out, 64), 1e-4, "head_bwd dw")
    _close(db, bref.grad, 1e-4, "head_bwd db")
    # fused first pass of the BatchNorm backward of the layer that produced the head's input:
    # a2 = relu(y*scale + shift) (the kernel recovers xhat from the activation it already holds)
    yb = _nhwc(torch.randn(n, 64, h, w, generator=g).to(dev))
    sc = (torch.rand(64, generator=g) + 0.5).to(dev); sh = torch.randn(64, generator=g).to(dev) * 0.3
    iv = (torch.rand(64, generator=g) + 0.5).to(dev)
    mu = -sh / sc + torch.randn(64, generator=g).to(dev) * 0.05     # consistent BatchNorm constants (beta small)
    a2 = torch.empty_like(yb)
    ops.bnrelu_apply(yb, sc, sh, a2)
    for scale_case in (sc, torch.cat([sc[:5] * 0, sc[5:]])):     # second case: gamma == 0 channels (direct route)
        if scale_case is not sc:
            ops.bnrelu_apply(yb, scale_case, sh, a2)
        fused = torch.zeros(128, device=dev); ref_sums = torch.zeros(128, device=dev)
        da2 = torch.empty_like(a); da3 = torch.empty_like(a)
        ops.head_bwd(a2, wt, dout, out, da3, dw, db, sigmoid)
        ops.head_bwd(a2, wt, dout, out, da2, dw, db, sigmoid, bn=ops.bn_ctx(yb, scale_case, sh, mu, iv, fused))
        ops.bn_bwd_reduce(da2, yb, scale_case, sh, mu, iv, ref_sums)
        torch.cuda.synchronize()
        assert torch.equal(da2, da3)
        _close(fused, ref_sums, 5e-3, "head_bwd fused BN reduction")
        # a=None: the activation recomputed from y (forward went through head_fwd_bn) -- bit for bit the same
        da4 = torch.empty_like(a); dw4 = torch.empty_like(dw); db4 = torch.empty_like(db); fused4 = torch.zeros(128, device=dev)
        ops.head_bwd(a2, wt, dout, out, da3, dw, db, sigmoid)
        ops.head_bwd(None, wt, dout, out, da4, dw4, db4, sigmoid, bn=ops.bn_ctx(yb, scale_case, sh, mu, iv, fused4))
        torch.cuda.synchronize()
        assert torch.equal(da4, da3)
        _close(dw4, dw, 1e-4, "head_bwd(a=None) dw")       # (fp32 atomics: the order differs between launches)
        _close(db4, db, 1e-4, "head_bwd(a=None) db")
        _close(fused4, ref_sums, 5e-3, "head_bwd(a=None) fused BN reduction")
        da5 = torch.empty_like(a)
        ops.head_bwd(None, wt, dout, out, da5, dw4, db4, sigmoid, bn=ops.bn_ctx(yb, scale_case, sh, mu, iv, None))
        torch.cuda.synchronize()
        assert torch.equal(da5, da3)

    # head_fwd_bn == bnrelu_apply_fin + head_fwd, bit for bit (outputs, saved BatchNorm constants, running statistics)
    count = n * h * w
    y32 = _nchw(yb)
    stats = torch.cat([y32.sum((0, 2, 3)), (y32 * y32).sum((0, 2, 3))]).contiguous()
    gamma = (torch.rand(64, generator=g) + 0.5).to(dev); beta = (torch.randn(64, generator=g) * 0.2).to(dev)
    res = []
    for fused_head in (False, True):
        rm, rv, nbt = torch.zeros(64, device=dev), torch.ones(64, device=dev), torch.zeros((), dtype=torch.int64, device=dev)
        aux = [torch.empty(64, device=dev) for _ in range(4)]
        fin = ops.bn_finalize_args(stats, gamma, beta, rm, rv, nbt, *aux, count, 1e-5, 0.1)
        o = torch.empty(n, n_out, h, w, device=dev)
        if fused_head:
            ops.head_fwd_bn(yb, fin, wt, b, o, sigmoid)
        else:
            act = torch.empty_like(yb)
            ops.bnrelu_apply_fin(yb, fin, act)
            ops.head_fwd(act, wt, b, o, sigmoid)
        torch.cuda.synchronize()
        res.append([o, rm, rv, nbt] + aux)
    for t0, t1 in zip(*res):
        assert torch.equal(t0, t1)
    assert int(res[1][3]) == 1


def test_adam_and_layout(cuda_device):
    _lib, ops, packing = _mods()
    dev = cuda_device
    g = torch.Generator(device="cpu").manual_seed(41)
    n = 10007
    p0 = torch.randn(n, generator=g).to(dev)
    for mode, make in ((0, lambda p: torch.optim.Adam([p], lr=1e-3, weight_decay=1e-4)),
                       (1, lambda p: torch.optim.AdamW([p], lr=1e-3, weight_decay=1e-2)),
                       (2, lambda p: torch.optim.SGD([p], lr=1e-2, momentum=0.9, weight_decay=1e-4))):
        pref = p0.clone().requires_grad_(True)
        opt = make(pref)
        p = torch.zeros(n + 5, device=dev)[:n]  # keep base 16B aligned
        p.copy_(p0)
        m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev)
        wd = opt.param_groups[0]["weight_decay"]; lr = opt.param_groups[0]["lr"]
        for step in range(1, 4):
            grad = torch.randn(n, generator=g).to(dev)
            pref.grad = grad.clone()
            opt.step()
            ops.adam_step(p, grad, m, v, lr, 0.9, 0.999, 1e-8, wd, step, 1.0, mode)
        torch.cuda.synchronize()
        _close(p, pref.detach(), 1e-6, f"optimizer mode {mode}")
    x = torch.randn(2, 5, 7, 9, generator=g).to(dev)
    nh = torch.empty(2, 7, 9, 5, dtype=torch.bfloat16, device=dev)
    ops.nchw_to_nhwc(x, nh)
    back = torch.empty_like(x)
    ops.nhwc_to_nchw(nh, back)
    torch.cuda.synchronize()
    assert torch.equal(nh, x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16))
    assert torch.equal(back, x.to(torch.bfloat16).float())


@pytest.mark.parametrize("n,h,w", [(2, 12, 20), (1, 33, 7), (3, 64, 64)])
def test_head_loss_bwd_is_the_unfused_sequence(cuda_device, n, h, w):
    """unetk_head_loss_bwd (one pass: BatchNorm + ReLU + 1x1 head + sigmoid + loss term + head backward + BatchNorm-
    backward sums) against the kernels it replaces, launched one by one: head_fwd_bn -> combined_loss -> head_bwd.
    Same arithmetic element for element except for the order of the 64-term dot product of the 1x1 conv (eight lanes
    per pixel here, one thread per pixel in head_fwd): outputs agree to the last bits of fp32, dA to a rare bf16
    rounding flip, the atomically accumulated sums to fp32 noise; the BatchNorm constants and running statistics are
    bit-identical."""
    _lib, ops, packing = _mods()
    dev = cuda_device
    g = torch.Generator(device="cpu").manual_seed(77)
    rw, sw, alpha, gamma = 0.7, 1.3, 0.25, 2.0
    image = torch.rand(n, 3, h, w, generator=g).to(dev)
    mask = (torch.rand(n, 1, h, w, generator=g) > 0.9).float().to(dev)
    count = n * h * w
    ys, fins, heads, buf = [], [], [], []
    for d, n_out in enumerate((3, 1)):
        yb = _nhwc((torch.randn(n, 64, h, w, generator=g) * 1.5 + 0.2).to(dev))
        y32 = _nchw(yb)
        stats = torch.cat([y32.sum((0, 2, 3)), (y32 * y32).sum((0, 2, 3))]).contiguous()
        gamma_ = (torch.rand(64, generator=g) + 0.5).to(dev); beta = (torch.randn(64, generator=g) * 0.2).to(dev)
        if d == 1:
            gamma_[3] = 0.0          # a degenerate channel (scale == 0) takes the direct route through y
        ys.append((yb, stats, gamma_, beta))
        heads.append(((torch.randn(n_out, 64, generator=g) / 8).to(dev), torch.randn(n_out, generator=g).to(dev)))

    def make_fin(d):
        yb, stats, gamma_, beta = ys[d]
        rm, rv, nbt = torch.zeros(64, device=dev), torch.ones(64, device=dev), torch.zeros((), dtype=torch.int64, device=dev)
        aux = [torch.empty(64, device=dev) for _ in range(4)]
        return ops.bn_finalize_args(stats, gamma_, beta, rm, rv, nbt, *aux, count, 1e-5, 0.1), (rm, rv, nbt, aux)

    # unfused reference sequence
    outs, ref = [], []
    for d in range(2):
        fin, keep = make_fin(d)
        o = torch.empty(n, heads[d][0].shape[0], h, w, device=dev)
        ops.head_fwd_bn(ys[d][0], fin, heads[d][0], heads[d][1], o, True)
        outs.append(o); ref.append(keep)
    d_r, d_a = torch.empty_like(outs[0]), torch.empty_like(outs[1])
    out3_ref = torch.empty(3, device=dev)
    ops.combined_loss(outs[0], outs[1], image, mask, d_r, d_a, out3_ref, ops.combined_loss_workspace(dev), rw, sw, alpha, gamma)
    ref_b = []
    for d, dout in enumerate((d_r, d_a)):
        rm, rv, nbt, aux = ref[d]
        da = torch.empty_like(ys[d][0]); dw = torch.empty_like(heads[d][0]); db = torch.empty_like(heads[d][1])
        sums = torch.zeros(128, device=dev)
        ops.head_bwd(None, heads[d][0], dout, outs[d], da, dw, db, True, bn=ops.bn_ctx(ys[d][0], aux[0], aux[1], aux[2], aux[3], sums))
        ref_b.append((da, dw, db, sums))
    # fused, twice (the workspace cleans itself up)
    work = ops.head_loss_workspace(dev)
    for rep in range(2):
        out3 = torch.zeros(3, device=dev)
        got = []
        for d in range(2):
            fin, keep = make_fin(d)
            o = torch.empty_like(outs[d]) if rep == 0 else None
            da = torch.empty_like(ys[d][0]); dw = torch.zeros_like(heads[d][0]); db = torch.zeros_like(heads[d][1])
            sums = torch.zeros(128, device=dev)
            ops.head_loss_bwd(ys[d][0], fin, heads[d][0], heads[d][1], image if d == 0 else mask, o, da, dw, db, sums, work, out3,
                              d, rw, sw, alpha, gamma)
            got.append((o, da, dw, db, sums, keep))
        torch.cuda.synchronize()
        for d in range(2):
            o, da, dw, db, sums, (rm, rv, nbt, aux) = got[d]
            if o is not None:
                assert torch.allclose(o, outs[d], rtol=0, atol=2e-6)
            _close(_nchw(da), _nchw(ref_b[d][0]), 8e-3, "fused dA")      # (one bf16 ulp of the largest element)
            assert (da != ref_b[d][0]).float().mean().item() < 2e-2
            _close(dw, ref_b[d][1], 1e-4, "fused dw")
            _close(db, ref_b[d][2], 1e-4, "fused db")
            _close(sums, ref_b[d][3], 5e-3, "fused BN sums")
            assert int(nbt) == 1 and torch.equal(rm, ref[d][0]) and torch.equal(rv, ref[d][1])
            for t0, t1 in zip(aux, ref[d][3]):
                assert torch.equal(t0, t1)
        assert torch.allclose(out3, out3_ref, rtol=1e-5, atol=1e-7), (out3, out3_ref)
        assert abs(float(out3[0]) - (rw * float(out3[1]) + sw * float(out3[2]))) <= 1e-6 * abs(float(out3[0]))


@pytest.mark.parametrize("mode", ["1", "2"])
def test_tma_store_epilogue_is_bit_identical(cuda_device, mode, monkeypatch):
    """UNETK_TMA_STORE: the epilogue stages its bf16 tile in swizzled shared memory and writes it with bulk tensor
    stores (clipped at the image border / the second destination's extent; pixel shuffle through a 5-D map for the
    transposed conv) instead of per-thread stores.  Same values, bit for bit, on full and ragged tiles."""
    _lib, ops, packing = _mods()
    dev = cuda_device
    g = torch.Generator(device="cpu").manual_seed(5)

    def bf(*shape):
        return (torch.randn(*shape, generator=g) * 0.5).to(dev).to(torch.bfloat16)

    def run_all():
        out = []
        for n, h, w, c1, c2, cout in [(2, 16, 16, 64, 0, 64), (1, 20, 28, 64, 64, 128), (2, 16, 24, 128, 0, 256)]:
            s1 = bf(n, h, w, c1)
            s2 = bf(n, h - 1, w - 1, c2) if c2 else None           # the second source one row / column short (F.pad path)
            wf, wd = bf(cout * 9 * (c1 + c2)), bf((c1 + c2) * 9 * cout)
            y = torch.zeros(n, h, w, cout, dtype=torch.bfloat16, device=dev)
            stats = torch.zeros(2 * cout, device=dev)
            ops.conv3x3_fwd(s1, s2, wf, y, stats=stats)
            dx1 = torch.zeros(n, h, w, c1, dtype=torch.bfloat16, device=dev)
            dx2 = torch.zeros(n, h - 1, w - 1, c2, dtype=torch.bfloat16, device=dev) if c2 else None
            ops.conv3x3_dgrad(y, wd, dx1, False, dx2, False)
            out += [y, stats, dx1] + ([dx2] if c2 else [])
        for n, h, w, cin, cout in [(2, 8, 8, 128, 64), (1, 12, 20, 256, 128)]:
            a = bf(n, h, w, cin)
            wt = bf(4 * cout * cin)
            bias = torch.randn(cout, generator=g).to(dev)
            up = torch.zeros(n, 2 * h, 2 * w, cout, dtype=torch.bfloat16, device=dev)
            ops.convt2x2_fwd(a, wt, bias, up)
            out.append(up)
        torch.cuda.synchronize()
        return out

    monkeypatch.setenv("UNETK_TMA_STORE", "0")
    g.manual_seed(5)
    ref = run_all()
    monkeypatch.setenv("UNETK_TMA_STORE", mode)
    g.manual_seed(5)
    got = run_all()
    for i, (a, b) in enumerate(zip(ref, got)):
        if a.dtype == torch.float32:          # BatchNorm sums: atomics
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-4), i
        else:
            assert torch.equal(a, b), i
