"""Pins oracle/cv_resize.py (the numpy restatement of OpenCV's uint8 INTER_AREA / INTER_LINEAR arithmetic) against the
real cv2, bit for bit, over the shapes the reference produces (crops of a 512^2 slice -> 125 x 250, then 250^2 -> 224^2,
T1H:355-358, 485-488) and the corner cases: integer scales (fast path, 2x2 and others), non-integer shrinking,
up-sampling in one or both dimensions, tiny crops."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from oracle import cv_resize as R


def _img(rng, h, w):
    base = rng.integers(0, 256, (h, w)).astype(np.uint8)
    if rng.random() < 0.5:               # smooth content too (CLAHE'd CT slices are smooth)
        base = cv2.GaussianBlur(base, (0, 0), 2.0)
    return base


AREA_CASES = [(300, 200), (250, 125), (500, 250), (750, 375), (500, 375), (251, 126), (333, 177), (512, 512), (260, 140),
              (249, 124), (100, 60), (400, 90), (120, 300), (20, 9), (250, 126), (501, 251), (1000, 250)]


@pytest.mark.parametrize("sh,sw", AREA_CASES)
def test_inter_area_to_125x250_is_bit_exact(sh, sw):
    rng = np.random.default_rng(sh * 1000 + sw)
    src = _img(rng, sh, sw)
    want = cv2.resize(src, dsize=(125, 250), interpolation=cv2.INTER_AREA)
    got = R.resize_area_u8(src, 125, 250)
    assert np.array_equal(got, want), (np.abs(got.astype(int) - want.astype(int)).max(), int((got != want).sum()))


@pytest.mark.parametrize("s,d", [(250, 224), (250, 256), (250, 512), (250, 125), (250, 96), (250, 250), (37, 224), (512, 224)])
def test_inter_linear_square_is_bit_exact(s, d):
    rng = np.random.default_rng(s * 1000 + d)
    src = _img(rng, s, s)
    want = cv2.resize(src, dsize=(d, d), interpolation=cv2.INTER_LINEAR)
    got = R.resize_linear_u8(src, d, d)
    assert np.array_equal(got, want), (np.abs(got.astype(int) - want.astype(int)).max(), int((got != want).sum()))


def test_inter_linear_rectangular_is_bit_exact():
    rng = np.random.default_rng(7)
    src = _img(rng, 180, 333)
    want = cv2.resize(src, dsize=(224, 100), interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(R.resize_linear_u8(src, 224, 100), want)


@pytest.mark.parametrize("sh,sw", [(630, 630), (512, 512), (1024, 1024), (1536, 1024), (700, 520), (160, 144), (300, 600), (513, 512)])
def test_inter_area_float64_to_512_is_bit_exact(sh, sw):
    """the NIfTI ingest resizes raw float64 slices to 512 x 512 with INTER_AREA (T1H:335)"""
    rng = np.random.default_rng(sh + sw)
    src = rng.normal(-300.0, 400.0, (sh, sw))
    want = cv2.resize(src, dsize=(64 if sh > 1000 else 512, 64 if sh > 1000 else 512), interpolation=cv2.INTER_AREA)
    got = R.resize_area_f64(src, want.shape[1], want.shape[0])
    assert got.dtype == np.float64 and np.array_equal(got, want), float(np.abs(got - want).max())
