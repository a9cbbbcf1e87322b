"""Evaluation-time input transform (deterministic half) on the CPU: the numpy oracle is
pinned against the libraries the reference calls (torchvision transforms on PIL images, src/dataset.py:141-152),
and the library's HOST tap-table function against the oracle."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import preprocess_oracle as P  # noqa: E402

SIZES = [(900, 900, 256), (1024, 1024, 256), (100, 130, 256), (40, 40, 32), (700, 1024, 512), (256, 256, 256),
         (300, 256, 256), (33, 47, 16), (5, 5, 9), (64, 48, 64)]


def images(seed, h, w):
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    img[: h // 3] = (img[: h // 3].astype(np.int32) * 7 // 8 + 32).astype(np.uint8)      # some structure, full range
    mask = (rng.random((h, w)) > 0.8).astype(np.uint8)                                   # 0 / 1, as src/dataset.py:104-106
    return img, mask


def reference_transforms(size):
    """get_transforms(image_size, is_train=False) as written in src/dataset.py:141-152."""
    import torchvision.transforms as transforms
    image_transform = transforms.Compose([
        transforms.Resize((size, size)),
        transforms.ToTensor(),
        transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225]),
    ])
    mask_transform = transforms.Compose([transforms.Resize((size, size)), transforms.ToTensor()])
    return image_transform, mask_transform


@pytest.mark.parametrize("h,w,size", SIZES)
def test_oracle_is_pillow_and_torchvision(h, w, size):
    pytest.importorskip("torchvision")
    from PIL import Image
    img, mask = images(h * 1000 + w, h, w)
    np.testing.assert_array_equal(P.resize(img, size, size), np.asarray(Image.fromarray(img).resize((size, size), Image.BILINEAR)))
    np.testing.assert_array_equal(P.resize(mask, size, size), np.asarray(Image.fromarray(mask).resize((size, size), Image.BILINEAR)))
    it, mt = reference_transforms(size)
    want_i = it(Image.fromarray(img)).numpy()
    want_m = mt(Image.fromarray(mask)).numpy()
    got_i, got_m = P.eval_image_transform(img, size), P.eval_mask_transform(mask, size)
    assert got_i.dtype == want_i.dtype == np.float32 and got_i.shape == (3, size, size) and got_m.shape == (1, size, size)
    np.testing.assert_array_equal(got_i, want_i)        # bit-exact
    np.testing.assert_array_equal(got_m, want_m)


def test_non_square_resize_matches_pillow():
    from PIL import Image
    img, _ = images(7, 61, 83)
    for oh, ow in ((32, 48), (61, 20), (100, 83), (7, 200)):
        np.testing.assert_array_equal(P.resize(img, oh, ow), np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR)))


@pytest.mark.parametrize("a,b", [(900, 256), (1024, 256), (100, 256), (256, 256), (5, 9), (4096, 64), (1, 3), (7, 1), (1408, 512)])
def test_library_tap_tables_are_the_oracles(a, b):
    """unetk_resize_ksize / unetk_resize_coeffs are host functions of the C ABI: callable without a device."""
    from tiaozhanbei_unet_b200 import _lib, ops
    bounds, taps = ops.resize_coeffs(a, b)
    ob, ot = P.coeffs(a, b)
    np.testing.assert_array_equal(bounds, ob)
    np.testing.assert_array_equal(taps, ot)
    assert taps.shape[1] == P.ksize(a, b) == _lib.get().unetk_resize_ksize(a, b)
    assert (bounds[:, 0] >= 0).all() and (bounds[:, 0] + bounds[:, 1] <= a).all()
    # weights sum to 1 in fixed point up to the rounding of each tap
    assert np.abs(taps.sum(1) - (1 << P.PRECISION_BITS)).max() <= taps.shape[1]


def test_transforms_have_no_cpu_path():
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    from tiaozhanbei_unet_b200 import transforms as T
    with pytest.raises(Exception):
        T.eval_image_transform(np.zeros((8, 8, 3), np.uint8), 4)
    with pytest.raises(TypeError):
        T.eval_image_transform(np.zeros((8, 8, 3), np.float32), 4)


# ---------------------------------------------------------------------------------------------
# training augmentations (src/dataset.py:131-138)
# ---------------------------------------------------------------------------------------------
def reference_train_transform(size):
    """get_transforms(image_size, is_train=True)[0] as written in src/dataset.py:131-139."""
    import torchvision.transforms as transforms
    return transforms.Compose([
        transforms.Resize((size, size)),
        transforms.RandomHorizontalFlip(p=0.5),
        transforms.RandomRotation(degrees=10),
        transforms.ColorJitter(brightness=0.1, contrast=0.1, saturation=0.1, hue=0.05),
        transforms.ToTensor(),
        transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225]),
    ])


def colour_cube():
    """All 2^24 RGB (or HSV) triples as a 4096 x 4096 image."""
    c = np.arange(256, dtype=np.uint8)
    return np.stack(np.meshgrid(c, c, c, indexing="ij"), -1).reshape(4096, 4096, 3)


def test_colour_conversions_are_pillows_over_all_colours():
    from PIL import Image
    cube = colour_cube()
    np.testing.assert_array_equal(P.luma(cube), np.asarray(Image.fromarray(cube).convert("L")))
    np.testing.assert_array_equal(P.rgb_to_hsv(cube), np.asarray(Image.fromarray(cube).convert("HSV")))
    np.testing.assert_array_equal(P.hsv_to_rgb(cube), np.asarray(Image.fromarray(cube, "HSV").convert("RGB")))


def test_blend_and_rotation_are_pillows():
    from PIL import Image
    a = np.repeat(np.arange(256, dtype=np.uint8)[:, None], 256, 1)
    b = np.ascontiguousarray(a.T)
    rng = np.random.default_rng(0)
    for alpha in list(rng.uniform(0.0, 2.0, 40)) + [0.0, 1.0, 0.9, 1.1, 0.5, 1.0000001, 0.9999999]:
        np.testing.assert_array_equal(P.blend(a, b, alpha), np.asarray(Image.blend(Image.fromarray(a), Image.fromarray(b), alpha)))
    for h, w in ((64, 64), (33, 47), (256, 256)):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        for angle in list(rng.uniform(-10, 10, 12)) + [10.0, -10.0, 1e-3, 45.0, 90.0, 180.0, 270.0, 360.0, 123.4]:
            want = np.asarray(Image.fromarray(img).rotate(angle, Image.NEAREST, False, None, fillcolor=(0, 0, 0)))
            np.testing.assert_array_equal(P.rotate_nearest(img, angle), want, err_msg=f"{h}x{w} angle {angle}")


@pytest.mark.parametrize("h,w,size", [(300, 300, 64), (100, 130, 128), (40, 40, 32)])
def test_train_transform_oracle_is_torchvisions(h, w, size):
    """Same seed, same draws, same pixels: the whole is_train=True pipeline against torchvision on PIL images."""
    pytest.importorskip("torchvision")
    from PIL import Image
    tf = reference_train_transform(size)
    batch = [images(h + i, h, w)[0] for i in range(5)]
    for seed in range(3):
        torch.manual_seed(seed)
        want = [tf(Image.fromarray(b)).numpy() for b in batch]
        torch.manual_seed(seed)
        params = P.draw_train_params(len(batch))
        for b, p, wnt in zip(batch, params, want):
            np.testing.assert_array_equal(P.train_image_transform(b, size, p), wnt)
        assert len({tuple(p["order"]) for p in params}) > 1 or seed      # the op order really varies


def test_product_host_parameters_are_the_oracles():
    """transforms.draw_train_params / _geometry (product host code) against the oracle's restatement."""
    from tiaozhanbei_unet_b200 import transforms as T
    torch.manual_seed(11)
    a = T.draw_train_params(8)
    torch.manual_seed(11)
    assert a == P.draw_train_params(8)
    for p in a:
        g = T._geometry(p["flip"], p["angle"], 256, 192)
        assert g[0] == 1 and g[1] == int(p["flip"])
        assert tuple(g[2:]) == P.affine_fixed_coeffs(P.rotate_matrix(p["angle"] % 360.0, 256, 192))
    assert T._geometry(False, 180.0, 8, 8)[0] == 2 and T._geometry(False, 90.0, 8, 8)[0] == 3
    assert T._geometry(True, 270.0, 8, 8)[:2] == [4, 1] and T._geometry(False, 360.0, 8, 8)[0] == 0
    assert T._geometry(False, 90.0, 8, 6)[0] == 1          # 90 degrees on a non-square image is a real affine


NEAREST_SIZES = [(900, 900, 512, 512), (1024, 1024, 256, 256), (100, 130, 256, 256), (500, 1240, 512, 1408),
                 (704, 256, 512, 512), (33, 47, 64, 16), (5, 5, 9, 3), (777, 1333, 511, 313)]


@pytest.mark.parametrize("h,w,oh,ow", NEAREST_SIZES)
def test_label_mask_transform_is_torchvisions(h, w, oh, ow):
    """src/gear_dataset.py:255-258 target transform: oracle vs torchvision, library host tables vs oracle."""
    pytest.importorskip("torchvision")
    import torchvision.transforms as transforms
    from PIL import Image
    from tiaozhanbei_unet_b200 import ops
    mask = np.random.default_rng(h + w).integers(0, 4, (h, w), dtype=np.uint8)
    tf = transforms.Compose([transforms.Resize((oh, ow), interpolation=Image.NEAREST)])
    want = torch.from_numpy(np.array(tf(Image.fromarray(mask, mode="L")))).long().numpy()
    np.testing.assert_array_equal(P.label_mask_transform(mask, oh, ow), want)
    np.testing.assert_array_equal(ops.resize_nearest_table(w, ow), P.nearest_table(w, ow))
    np.testing.assert_array_equal(ops.resize_nearest_table(h, oh), P.nearest_table(h, oh))
