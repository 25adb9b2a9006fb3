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
        assert np.array_equal(mid[k], fused), "250x250 stage: %d pixels differ" % int((mid[k] != fused).sum())   # integer work: bit-exact
        final = (np.uint8(cv2.resize(fused, dsize=(224, 224), interpolation=cv2.INTER_LINEAR)) / 255.0).astype(np.float32)
        assert np.array_equal(got[k, :, :, 0], final)


def test_nifti_case_to_network_input_vs_reference_lines(tmp_path):
    """file -> (N,224,224,1) through nifti.preprocess_case (reader + host slice pipeline + device CLAHE / crop / resize)
    against the reference's own sequence of calls restated with cv2 (T1H:310-368, 485-488, 678-686)."""
    from test_nifti_cpu import reference_read_nii_demo, synthetic_case
    N = importlib.import_module(PKG + ".nifti")
    PP = importlib.import_module(PKG + ".preprocess")
    ct, lung, inf = synthetic_case(s=20, h=160, w=144)
    N.save_nii(str(tmp_path / "ct.nii.gz"), np.round(ct).astype(np.int16))
    N.save_nii(str(tmp_path / "lung.nii"), lung.astype(np.uint8))
    N.save_nii(str(tmp_path / "inf.nii"), inf.astype(np.uint8), byteorder=">")
    x, y, boxes = N.preprocess_case(str(tmp_path / "ct.nii.gz"), str(tmp_path / "lung.nii"), str(tmp_path / "inf.nii"))
    lungs = reference_read_nii_demo(lung.astype(np.uint8).astype(np.float64))
    cts = reference_read_nii_demo(np.round(ct).astype(np.int16).astype(np.float64))
    infs = reference_read_nii_demo(inf.astype(np.uint8).astype(np.float64))
    assert x.shape == (len(cts), 224, 224, 1) and y.shape == x.shape and x.dtype == np.float32
    clahe = cv2.createCLAHE(clipLimit=3.0, tileGridSize=(8, 8))
    for k in range(len(cts)):
        m = lungs[k].copy()
        m[m > 0] = 1
        want_box = PP.cropper_boxes(np.uint8(m))
        assert boxes[k].tolist() == want_box
        a, b, c, d, e, f, g, h = want_box
        for got, src in ((x, clahe.apply(np.uint8(cts[k] * 255))), (y, np.uint8(np.nan_to_num(infs[k]) * 255))):
            i1 = cv2.resize(src[b:b + d, a:a + c], dsize=(125, 250), interpolation=cv2.INTER_AREA)
            i2 = cv2.resize(src[f:f + h, e:e + g], dsize=(125, 250), interpolation=cv2.INTER_AREA)
            fused = np.concatenate((i1, i2), axis=1)
            final = np.uint8(cv2.resize(fused, dsize=(224, 224), interpolation=cv2.INTER_LINEAR)) / 255.0
            assert np.array_equal(got[k, :, :, 0], final.astype(np.float32))       # both resize stages are bit-exact


AREA_CASES = [(300, 200), (250, 125), (500, 250), (750, 375), (500, 375), (251, 126), (333, 177), (512, 512), (260, 140),
              (249, 124), (100, 60), (400, 90), (120, 300), (20, 9), (250, 126), (501, 251), (1000, 250)]


def test_resize_u8_is_bit_exact_against_cv2():
    """b2u_resize_u8 against cv2.resize over OpenCV's three INTER_AREA regimes (integer scales incl. 2x2, fractional
    shrinking, up-sampling in one or both dimensions) and the fixed-point INTER_LINEAR, on noise and on smooth images"""
    PP = importlib.import_module(PKG + ".preprocess")
    for sh, sw in AREA_CASES:
        rng = np.random.default_rng(sh * 1000 + sw)
        src = rng.integers(0, 256, (2, sh, sw)).astype(np.uint8)
        src[1] = cv2.GaussianBlur(src[1], (0, 0), 2.0)
        got = PP.resize(src, (125, 250), PP.INTER_AREA)
        for k in range(2):
            want = cv2.resize(src[k], dsize=(125, 250), interpolation=cv2.INTER_AREA)
            assert np.array_equal(got[k], want), ("area", sh, sw, k, int((got[k] != want).sum()))
    got = PP.resize(np.full((630, 630), 7, np.uint8), (512, 512), PP.INTER_AREA)
    assert got.shape == (512, 512) and (got == 7).all()
    for s_, d_ in ((250, 224), (250, 256), (250, 512), (250, 96), (37, 224), (512, 224)):
        rng = np.random.default_rng(s_ * 1000 + d_)
        src = rng.integers(0, 256, (s_, s_)).astype(np.uint8)
        want = cv2.resize(src, dsize=(d_, d_), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(PP.resize(src, (d_, d_), PP.INTER_LINEAR), want), ("linear", s_, d_)
    lib = importlib.import_module(PKG + "._lib")
    with pytest.raises(lib.B2UError):
        PP.resize(np.zeros((8, 8), np.uint8), (4, 4), 2)          # INTER_CUBIC: not part of the reference's path


@pytest.mark.parametrize("h,w", [(630, 630), (512, 512), (1024, 1024), (160, 144), (700, 520), (768, 512)])
def test_volume_slices_on_device_equals_the_reference_cv2_sequence(h, w):
    """N3: rot90 / slice window on the host, INTER_AREA to 512 x 512 on float64 + per-slice min-max on the GPU: exactly the
    arrays the reference's read_nii loop builds with cv2.resize and numpy (T1H:288-297, 335-337), NaN slices included"""
    N = importlib.import_module(PKG + ".nifti")
    rng = np.random.default_rng(h + w)
    vol = rng.normal(-400.0, 350.0, (h, w, 10))
    vol[:, :, 4] = 3.0                                                # a constant slice: 0/0 -> NaN in both
    want = N.volume_slices(vol, 512, device=False)
    got = N.volume_slices(vol, 512, device=True)
    assert got.shape == want.shape == (6, 512, 512) and got.dtype == np.float64
    # (a constant slice is 0/0 = NaN where the resize reproduces the constant exactly -- copy / integer scales --
    # and finite noise where the area weights do not sum to exactly one: either way both sides must agree)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    if (h, w) in ((512, 512), (1024, 1024)):
        assert np.isnan(want[2]).all()
    assert np.array_equal(np.nan_to_num(got, nan=-1.0), np.nan_to_num(want, nan=-1.0))
