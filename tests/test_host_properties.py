"""Property tests (hypothesis) of the product's HOST arithmetic — the parts of the device paths that run on the
CPU and must agree with numpy / Pillow for every size, not only the parametrised ones: percentile index and
interpolation, Pillow's bilinear tap tables and nearest index tables (host entry points of the C ABI), the
rotation's fixed-point coefficients."""
import os
import sys

import numpy as np
import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import preprocess_oracle as P  # noqa: E402

SETTINGS = dict(max_examples=60, deadline=None, derandomize=True, database=None)      # same examples every run


@settings(**SETTINGS)
@given(n=st.integers(1, 5000), q=st.floats(0, 100), seed=st.integers(0, 2 ** 16))
def test_percentile_host_half_equals_numpy(n, q, seed):
    from tiaozhanbei_unet_b200.utils import percentile_lerp, percentile_rank
    a = (np.random.default_rng(seed).random(n, dtype=np.float32) - np.float32(0.3)) * np.float32(7)
    srt = np.sort(a)
    lo, hi, gamma = percentile_rank(n, q)
    got, want = percentile_lerp(srt[lo], srt[hi], gamma), np.percentile(a, q)
    assert got == want and got.dtype == want.dtype


@settings(**SETTINGS)
@given(a=st.integers(1, 3000), b=st.integers(1, 1500))
def test_resize_tables_equal_the_oracles(a, b):
    from tiaozhanbei_unet_b200 import ops
    bounds, taps = ops.resize_coeffs(a, b)
    ob, ot = P.coeffs(a, b)
    assert np.array_equal(bounds, ob) and np.array_equal(taps, ot)
    assert np.array_equal(ops.resize_nearest_table(a, b), P.nearest_table(a, b))


@settings(max_examples=25, deadline=None, derandomize=True, database=None)
@given(h=st.integers(1, 40), w=st.integers(1, 40), oh=st.integers(1, 40), ow=st.integers(1, 40), seed=st.integers(0, 999))
def test_resize_oracle_equals_pillow_for_any_size(h, w, oh, ow, seed):
    from PIL import Image
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    assert np.array_equal(P.resize(img, oh, ow), np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR)))
    mask = rng.integers(0, 4, (h, w), dtype=np.uint8)
    want = np.asarray(Image.fromarray(mask).resize((ow, oh), Image.NEAREST)).astype(np.int64)
    assert np.array_equal(P.label_mask_transform(mask, oh, ow), want)


@settings(**SETTINGS)
@given(angle=st.floats(-359.9, 359.9), w=st.integers(2, 600), h=st.integers(2, 600), flip=st.booleans())
def test_rotation_coefficients_equal_the_oracles(angle, w, h, flip):
    from tiaozhanbei_unet_b200 import transforms as T
    a = angle % 360.0
    g = T._geometry(flip, angle, w, h)
    assert g[1] == int(flip)
    if a == 0:
        assert g[0] == 0
    elif a == 180:
        assert g[0] == 2
    elif a in (90, 270) and w == h:
        assert g[0] in (3, 4)
    else:
        assert g[0] == 1 and tuple(g[2:]) == P.affine_fixed_coeffs(P.rotate_matrix(a, w, h))


@settings(max_examples=20, deadline=None, derandomize=True, database=None)
@given(angle=st.floats(-180, 180), seed=st.integers(0, 999), h=st.integers(2, 48), w=st.integers(2, 48))
def test_rotation_oracle_equals_pillow_for_any_angle(angle, seed, h, w):
    from PIL import Image
    img = np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)
    # plus angles whose sine rounds to zero (Pillow's scale-only path) around 0 and 180 degrees
    for a in (angle, 1e-20, -1e-20, 2e-293, 180.0 + 1e-13, 180.0 - 1e-13, 360.0 - 1e-13):
        want = np.asarray(Image.fromarray(img).rotate(a, Image.NEAREST, False, None, fillcolor=(0, 0, 0)))
        assert np.array_equal(P.rotate_nearest(img, a), want), a
