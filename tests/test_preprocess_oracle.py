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
    with pytest.raises(NotImplementedError):
        T.get_transforms(256, is_train=True)
    with pytest.raises(TypeError):
        T.eval_image_transform(np.zeros((8, 8, 3), np.float32), 4)
