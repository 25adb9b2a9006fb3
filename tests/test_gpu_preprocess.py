"""P1-P3 device kernels against the reference's own OpenCV calls (cv2 is present on the box and is its own oracle,
SURVEY.md 8c): CLAHE bit-exact; crop + INTER_AREA resize + hconcat + INTER_LINEAR resize within 1 LSB."""
import importlib

import cv2
import numpy as np
import pytest

from conftest import PKG

pytestmark = pytest.mark.gpu


def _slices(n, size, seed):
    S = importlib.import_module(PKG + ".synthetic")
    x, lung = S.make_slices(n, size, seed=seed, task="lung")
    return x[..., 0], lung[..., 0]


@pytest.mark.parametrize("shape,clip,tiles", [((512, 512), 3.0, 8), ((256, 384), 2.0, 8), ((64, 64), 4.0, 4), ((512, 512), 40.0, 8)])
def test_clahe_bit_exact_vs_cv2(shape, clip, tiles):
    PP = importlib.import_module(PKG + ".preprocess")
    rng = np.random.default_rng(1)
    imgs = [np.uint8(_slices(1, 512, 3)[0][0][:shape[0], :shape[1]] * 255), rng.integers(0, 256, shape, dtype=np.uint8),
            np.zeros(shape, np.uint8), np.full(shape, 200, np.uint8)]
    batch = np.stack(imgs)
    got = PP.clahe_enhancer(batch, clip_limit=clip, tiles=tiles)
    cl = cv2.createCLAHE(clipLimit=clip, tileGridSize=(tiles, tiles))
    for k, im in enumerate(imgs):
        want = cl.apply(im)
        assert np.array_equal(got[k], want), "image %d: %d pixels differ" % (k, int((got[k] != want).sum()))


def test_clahe_reference_call_shape():
    """clahe_enhancer(test_img, demo) as the reference calls it: float [0,1] image in, uint8 out (T1H:163-170)"""
    PP = importlib.import_module(PKG + ".preprocess")
    x, _ = _slices(1, 512, 5)
    out = PP.clahe_enhancer(x[0], 0)
    want = cv2.createCLAHE(clipLimit=3.0, tileGridSize=(8, 8)).apply(np.uint8(x[0] * 255))
    assert out.dtype == np.uint8 and np.array_equal(out, want)
    lib = importlib.import_module(PKG + "._lib")
    with pytest.raises(lib.B2UError):
        PP.clahe_enhancer(np.zeros((100, 100)), tiles=8)          # not divisible by the tile grid


def test_crop_resize_vs_cv2():
    PP = importlib.import_module(PKG + ".preprocess")
    x, lung = _slices(6, 512, 7)
    cts = np.uint8(x * 255)
    masks = np.uint8((lung > 0.5) * 255)
    boxes = [PP.cropper_boxes(m) for m in masks]
    boxes[3] = [40, 60, 100, 200, 300, 50, 90, 260]            # narrower than 125: OpenCV's up-scaling rule in x
    got, mid = PP.crop_resize(cts, boxes)
    assert got.shape == (6, 224, 224, 1) and got.dtype == np.float32
    for k in range(6):
        x0, y0, w, h, p, q, r, s = boxes[k]
        c1 = cv2.resize(cts[k][y0:y0 + h, x0:x0 + w], dsize=(125, 250), interpolation=cv2.INTER_AREA)
        c2 = cv2.resize(cts[k][q:q + s, p:p + r], dsize=(125, 250), interpolation=cv2.INTER_AREA)
        fused = np.concatenate((c1, c2), axis=1)
        d = np.abs(mid[k].astype(int) - fused.astype(int))
        assert d.max() <= 1, "250x250 stage differs by %d" % d.max()
        final = np.uint8(cv2.resize(mid[k], dsize=(224, 224), interpolation=cv2.INTER_LINEAR)) / 255.0
        assert np.abs(got[k, :, :, 0] - final).max() <= 1.0 / 255 + 1e-6
