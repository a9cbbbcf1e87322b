"""Evaluation-time input transform on the device (deterministic half) through the C ABI:
bit-exact against torchvision + Pillow (the libraries src/dataset.py:141-152 calls) and the numpy oracle."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import preprocess_oracle as P  # noqa: E402
from test_preprocess_oracle import SIZES, images, reference_transforms  # noqa: E402


@pytest.mark.parametrize("h,w,size", SIZES)
def test_eval_transforms_are_bit_exact(cuda_device, h, w, size):
    pytest.importorskip("torchvision")
    from PIL import Image
    from tiaozhanbei_unet_b200 import transforms as T
    batch = [images(h * 1000 + w + i, h, w) for i in range(3)]
    imgs = np.stack([b[0] for b in batch])
    masks = np.stack([b[1] for b in batch])
    it, mt = reference_transforms(size)
    want_i = torch.stack([it(Image.fromarray(b[0])) for b in batch])
    want_m = torch.stack([mt(Image.fromarray(b[1])) for b in batch])
    got_i = T.eval_image_transform(torch.from_numpy(imgs).to(cuda_device), size)
    got_m = T.eval_mask_transform(torch.from_numpy(masks).to(cuda_device), size)
    assert got_i.is_cuda and got_i.dtype == torch.float32 and got_i.shape == (3, 3, size, size)
    assert got_m.shape == (3, 1, size, size)
    assert torch.equal(got_i.cpu(), want_i), (got_i.cpu() - want_i).abs().max()
    assert torch.equal(got_m.cpu(), want_m)
    # the resized bytes themselves, against Pillow and the oracle
    u8 = T.resize_u8(torch.from_numpy(imgs).to(cuda_device), size).cpu().numpy()
    for i, b in enumerate(batch):
        np.testing.assert_array_equal(u8[i], np.asarray(Image.fromarray(b[0]).resize((size, size), Image.BILINEAR)))
    np.testing.assert_array_equal(u8[0], P.resize(imgs[0], size, size))
    m8 = T.resize_u8(torch.from_numpy(masks).to(cuda_device), size).cpu().numpy()
    np.testing.assert_array_equal(m8[1], P.resize(masks[1], size, size))


def test_single_pil_image_and_get_transforms(cuda_device):
    from PIL import Image
    from tiaozhanbei_unet_b200 import transforms as T
    img, mask = images(11, 90, 120)
    it, mt = reference_transforms(64)
    image_transform, mask_transform = T.get_transforms(64, is_train=False)
    got_i, got_m = image_transform(Image.fromarray(img)), mask_transform(Image.fromarray(mask))
    assert got_i.shape == (3, 64, 64) and got_m.shape == (1, 64, 64) and got_i.is_cuda
    assert torch.equal(got_i.cpu(), it(Image.fromarray(img)))
    assert torch.equal(got_m.cpu(), mt(Image.fromarray(mask)))
    # non-square target, bytes only
    out = T.resize_u8(img, (32, 48)).cpu().numpy()
    np.testing.assert_array_equal(out, np.asarray(Image.fromarray(img).resize((48, 32), Image.BILINEAR)))
    with pytest.raises(TypeError):
        T.eval_image_transform(torch.zeros(2, 8, 8, 3, device=cuda_device), 4)
    with pytest.raises(ValueError):
        T.eval_image_transform(torch.zeros(2, 8, 8, 4, dtype=torch.uint8, device=cuda_device), 4)


def test_transformed_batch_feeds_the_model(cuda_device):
    """uint8 MVTec-sized images -> device transform -> UNet eval -> sigmoid map: the transform's output is what the
    model takes (NCHW fp32), and equals the model on the torchvision-transformed batch bit for bit."""
    from PIL import Image
    from tiaozhanbei_unet_b200 import transforms as T
    from tiaozhanbei_unet_b200.model import UNet
    torch.manual_seed(0)
    model = UNet(3, 1).to(cuda_device).eval()
    batch = [images(50 + i, 300, 300)[0] for i in range(2)]
    it, _ = reference_transforms(64)
    x_ref = torch.stack([it(Image.fromarray(b)) for b in batch]).to(cuda_device)
    x_dev = T.eval_image_transform(np.stack(batch), 64, device=cuda_device)
    assert torch.equal(x_ref, x_dev)
    with torch.no_grad():
        assert torch.equal(torch.sigmoid(model(x_dev)), torch.sigmoid(model(x_ref)))


def test_transform_throughput(cuda_device):
    """Printed for the log (HBM-bound byte work): 16 MVTec-sized images (1024 x 1024 RGB) -> 256 x 256 normalised."""
    from tiaozhanbei_unet_b200 import transforms as T
    x = torch.randint(0, 256, (16, 1024, 1024, 3), dtype=torch.uint8, device=cuda_device)
    T.eval_image_transform(x, 256)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        T.eval_image_transform(x, 256)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    moved = x.numel() + 2 * 16 * 1024 * 256 * 3 + 16 * 3 * 256 * 256 * 4
    print(f"\neval_image_transform 16 x 1024^2 -> 256^2: {ms:.3f} ms, {16 / ms * 1e3:.0f} img/s, {moved / ms / 1e6:.0f} GB/s")
    assert 16 / ms * 1e3 > 5000


# ---------------------------------------------------------------------------------------------
# training augmentations (src/dataset.py:131-138)
# ---------------------------------------------------------------------------------------------
def test_flip_rotate_is_bit_exact(cuda_device):
    from tiaozhanbei_unet_b200 import ops, transforms as T
    rng = np.random.default_rng(2)
    for h, w in ((64, 64), (33, 47), (256, 256)):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        cases = [(f, a) for f in (False, True) for a in list(rng.uniform(-10, 10, 4)) + [0.0, 180.0, 90.0, 270.0, 45.0, -123.4]]
        src = torch.from_numpy(np.stack([img] * len(cases))).to(cuda_device)
        geom = torch.tensor([T._geometry(f, a, w, h) for f, a in cases], dtype=torch.int32, device=cuda_device)
        out = ops.flip_rotate_u8(src, torch.empty_like(src), geom).cpu().numpy()
        for i, (f, a) in enumerate(cases):
            want = P.rotate_nearest(img[:, ::-1].copy() if f else img, a)
            np.testing.assert_array_equal(out[i], want, err_msg=f"{h}x{w} flip {f} angle {a}")


def test_colour_jitter_stages_are_bit_exact(cuda_device):
    from test_preprocess_oracle import colour_cube
    from tiaozhanbei_unet_b200 import ops
    dev = cuda_device
    # hue over all 2^24 colours (RGB -> HSV -> +delta -> RGB), two deltas
    cube = colour_cube()
    for delta in (0, 37, 243):
        x = torch.from_numpy(cube)[None].to(dev).clone()
        ops.color_jitter_u8(x, torch.tensor([[3, -1, -1, -1]], dtype=torch.int32, device=dev),
                            torch.tensor([[float(delta), 0, 0, 0]], dtype=torch.float32, device=dev))
        hsv = P.rgb_to_hsv(cube)
        hsv[..., 0] = (hsv[..., 0].astype(np.int32) + delta).astype(np.uint8)
        np.testing.assert_array_equal(x[0].cpu().numpy(), P.hsv_to_rgb(hsv), err_msg=f"hue delta {delta}")
    # the three blends, interpolating and extrapolating factors, each alone
    rng = np.random.default_rng(4)
    img = rng.integers(0, 256, (96, 80, 3), dtype=np.uint8)
    img[:20] = (img[:20] // 16)              # dark rows: extrapolation clamps at 0 / truncation matters
    factors = [0.0, 1.0, 0.9, 1.1, 0.3, 1.7, 0.9999999, 1.0000001, float(np.float32(0.95))]
    for op in (0, 1, 2):
        x = torch.from_numpy(np.stack([img] * len(factors))).to(dev)
        ops.color_jitter_u8(x, torch.tensor([[op, -1, -1, -1]] * len(factors), dtype=torch.int32, device=dev),
                            torch.tensor([[f, 0, 0, 0] for f in factors], dtype=torch.float32, device=dev))
        for i, f in enumerate(factors):
            want = P.color_jitter(img, [op], (f, f, f, 0.0))
            np.testing.assert_array_equal(x[i].cpu().numpy(), want, err_msg=f"op {op} factor {f}")
    # all 24 orders on one image
    import itertools
    orders = list(itertools.permutations(range(4)))
    fac = (0.93, 1.08, 1.04, -0.031)
    x = torch.from_numpy(np.stack([img] * 24)).to(dev)
    per_op = {0: fac[0], 1: fac[1], 2: fac[2], 3: float(P.hue_delta(fac[3]))}
    ops.color_jitter_u8(x, torch.tensor(orders, dtype=torch.int32, device=dev),
                        torch.tensor([[per_op[o] for o in order] for order in orders], dtype=torch.float32, device=dev))
    for i, order in enumerate(orders):
        np.testing.assert_array_equal(x[i].cpu().numpy(), P.color_jitter(img, list(order), fac), err_msg=str(order))


@pytest.mark.parametrize("h,w,size", [(300, 300, 64), (100, 130, 128), (900, 900, 256)])
def test_train_transform_is_torchvisions(cuda_device, h, w, size):
    """Same seed -> the batch the reference's single-process loader would produce, bit for bit."""
    pytest.importorskip("torchvision")
    from PIL import Image
    from test_preprocess_oracle import reference_train_transform
    from tiaozhanbei_unet_b200 import transforms as T
    tf = reference_train_transform(size)
    batch = [images(h + i, h, w)[0] for i in range(6)]
    for seed in range(3):
        torch.manual_seed(seed)
        want = torch.stack([tf(Image.fromarray(b)) for b in batch])
        torch.manual_seed(seed)
        got = T.train_image_transform(torch.from_numpy(np.stack(batch)).to(cuda_device), size)
        assert got.shape == want.shape and got.dtype == torch.float32
        assert torch.equal(got.cpu(), want), (seed, (got.cpu() != want).sum().item())
    image_transform, mask_transform = T.get_transforms(size, is_train=True)
    torch.manual_seed(5)
    one = image_transform(Image.fromarray(batch[0]))
    torch.manual_seed(5)
    assert torch.equal(one.cpu(), tf(Image.fromarray(batch[0])))


def test_train_transform_throughput(cuda_device):
    from tiaozhanbei_unet_b200 import transforms as T
    x = torch.randint(0, 256, (16, 1024, 1024, 3), dtype=torch.uint8, device=cuda_device)
    params = T.draw_train_params(16)
    T.train_image_transform(x, 256, params)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        T.train_image_transform(x, 256, params)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"\ntrain_image_transform 16 x 1024^2 -> 256^2 (flip, rotate, 4 jitter stages): {ms:.3f} ms, {16 / ms * 1e3:.0f} img/s")
    assert 16 / ms * 1e3 > 2000


def test_label_mask_transform_is_bit_exact(cuda_device):
    from test_preprocess_oracle import NEAREST_SIZES
    from tiaozhanbei_unet_b200 import transforms as T
    for h, w, oh, ow in NEAREST_SIZES:
        masks = np.random.default_rng(h * w).integers(0, 4, (3, h, w), dtype=np.uint8)
        got = T.label_mask_transform(torch.from_numpy(masks).to(cuda_device), (oh, ow))
        assert got.dtype == torch.int64 and got.shape == (3, oh, ow)
        for i in range(3):
            np.testing.assert_array_equal(got[i].cpu().numpy(), P.label_mask_transform(masks[i], oh, ow))
    one = T.label_mask_transform(masks[0], (oh, ow), device=cuda_device)
    assert one.shape == (oh, ow)


def test_gear_train_transform_settings(cuda_device):
    """The Gear pipeline's ranges (src/gear_dataset.py:236-243) on a non-square target, against torchvision."""
    pytest.importorskip("torchvision")
    import torchvision.transforms as transforms
    from PIL import Image
    from tiaozhanbei_unet_b200 import transforms as T
    size = (96, 128)
    tf = transforms.Compose([
        transforms.Resize(size), transforms.RandomHorizontalFlip(p=0.5), transforms.RandomRotation(10),
        transforms.ColorJitter(brightness=0.2, contrast=0.2, saturation=0.2, hue=0.1), transforms.ToTensor(),
        transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    batch = [images(70 + i, 200, 310)[0] for i in range(5)]
    for seed in (0, 1):
        torch.manual_seed(seed)
        want = torch.stack([tf(Image.fromarray(b)) for b in batch])
        torch.manual_seed(seed)
        got = T.train_image_transform(torch.from_numpy(np.stack(batch)).to(cuda_device), size, degrees=10,
                                      brightness=0.2, contrast=0.2, saturation=0.2, hue=0.1)
        assert torch.equal(got.cpu(), want), (seed, (got.cpu() != want).sum().item())
