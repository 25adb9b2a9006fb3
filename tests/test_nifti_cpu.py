"""NIfTI ingest on the CPU: the from-scratch NIfTI-1 reader against hand-packed files (the format's
published header layout: sizeof_hdr 348, dim @40, datatype/bitpix @70, vox_offset/scl_slope/scl_inter @108, magic
@344; Fortran voxel order) and the slice pipeline against an inline restatement of the reference's own lines
(task1_preprocessing_plus_unet_with_comments.py:286-297, 331-345)."""
import gzip
import importlib
import struct

import cv2
import numpy as np
import pytest

from conftest import PKG

N = importlib.import_module(PKG + ".nifti")


def hand_packed(bo, code, bitpix, shape, payload, slope=0.0, inter=0.0, vox_offset=352.0):
    h = bytearray(int(vox_offset))
    struct.pack_into(bo + "i", h, 0, 348)
    struct.pack_into(bo + "8h", h, 40, len(shape), *(list(shape) + [1] * (7 - len(shape))))
    struct.pack_into(bo + "2h", h, 70, code, bitpix)
    struct.pack_into(bo + "3f", h, 108, vox_offset, slope, inter)
    h[344:348] = b"n+1\x00"
    return bytes(h) + payload


def test_reader_known_answers(tmp_path):
    # 2 x 3 x 2 int16 volume, big-endian, first index fastest: voxel (i,j,k) = 100*k + 10*j + i
    vals = [100 * k + 10 * j + i for k in range(2) for j in range(3) for i in range(2)]
    p = tmp_path / "be.nii"
    p.write_bytes(hand_packed(">", 4, 16, (2, 3, 2), struct.pack(">12h", *vals)))
    a, hdr = N.load_nii(str(p))
    assert a.dtype == np.float64 and a.shape == (2, 3, 2) and hdr["byteorder"] == ">"
    assert a[1, 2, 0] == 21 and a[0, 1, 1] == 110 and a[1, 0, 1] == 101
    # uint8 with scaling, little-endian, gzip, voxels after a 16-byte header extension
    p2 = tmp_path / "le.nii.gz"
    p2.write_bytes(gzip.compress(hand_packed("<", 2, 8, (4, 2), bytes(range(8)), slope=2.0, inter=-1024.0, vox_offset=368.0)))
    b, hdr2 = N.load_nii(str(p2))
    assert b.shape == (4, 2) and b[3, 1] == 7 * 2.0 - 1024.0 and b[0, 0] == -1024.0 and hdr2["vox_offset"] == 368
    # float32, slope 0 = "no scaling"
    p3 = tmp_path / "f.nii"
    p3.write_bytes(hand_packed("<", 16, 32, (3,), struct.pack("<3f", 1.5, -2.25, 1e6)))
    c, _ = N.load_nii(str(p3))
    assert c.tolist() == [1.5, -2.25, 1e6]


def test_reader_rejects_bad_files(tmp_path):
    p = tmp_path / "x.nii"
    p.write_bytes(b"\x00" * 100)
    with pytest.raises(ValueError):
        N.load_nii(str(p))
    p.write_bytes(hand_packed("<", 4, 16, (4, 4), b"\x00" * 8))           # truncated voxel block
    with pytest.raises(ValueError):
        N.load_nii(str(p))
    bad = bytearray(hand_packed("<", 4, 16, (2,), b"\x00" * 4))
    bad[344:348] = b"ni1\x00"                                             # header/image pair: not supported
    p.write_bytes(bytes(bad))
    with pytest.raises(ValueError):
        N.load_nii(str(p))


@pytest.mark.parametrize("dtype,bo,gz", [("u1", "<", False), ("i2", ">", True), ("f4", "<", True), ("f8", ">", False), ("u2", "<", False)])
def test_writer_reader_roundtrip(tmp_path, dtype, bo, gz):
    rng = np.random.default_rng(3)
    a = (rng.random((5, 7, 3)) * 200).astype(dtype)
    p = tmp_path / ("v.nii.gz" if gz else "v.nii")
    N.save_nii(str(p), a, byteorder=bo)
    b, hdr = N.load_nii(str(p))
    assert hdr["byteorder"] == bo and np.array_equal(b, a.astype(np.float64))


def reference_read_nii_demo(array, img_size=512):
    """T1H:286-297 verbatim semantics (array = ct_scan.get_fdata())"""
    array = np.rot90(np.array(array))
    slices = array.shape[2]
    array = array[:, :, round(slices * 0.2):round(slices * 0.8)]
    array = np.reshape(np.rollaxis(array, 2), (array.shape[2], array.shape[0], array.shape[1], 1))
    data = []
    for img_no in range(0, array.shape[0]):
        img = cv2.resize(array[img_no], dsize=(img_size, img_size), interpolation=cv2.INTER_AREA)
        xmax, xmin = img.max(), img.min()
        img = (img - xmin) / (xmax - xmin)
        data.append(img)
    return data


def synthetic_case(s=20, h=96, w=80, seed=5):
    """CT-like volume + two-blob lung mask + infection mask, in the file orientation (before rot90)"""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    lung = np.zeros((h, w, s), np.float64)
    ct = np.zeros((h, w, s), np.float64)
    inf = np.zeros((h, w, s), np.float64)
    for k in range(s):
        r = 0.8 + 0.2 * np.sin(k / s * np.pi)
        b1 = ((yy - h * 0.30) / (h * 0.16 * r)) ** 2 + ((xx - w * 0.5) / (w * 0.30 * r)) ** 2 < 1
        b2 = ((yy - h * 0.72) / (h * 0.13 * r)) ** 2 + ((xx - w * 0.5) / (w * 0.26 * r)) ** 2 < 1
        lung[:, :, k] = (b1 | b2) * (1 + (k % 2))             # labels 1 / 2 like the dataset's lung masks
        ct[:, :, k] = -1000 + 900 * (~(b1 | b2)) + 60 * rng.standard_normal((h, w))
        inf[:, :, k] = b1 & (((yy - h * 0.3) ** 2 + (xx - w * 0.45) ** 2) < (h * 0.07) ** 2)
    return ct, lung, inf


def test_volume_slices_match_reference_lines():
    ct, lung, _ = synthetic_case()
    for vol in (ct, lung):
        want = reference_read_nii_demo(vol, 128)
        got = N.volume_slices(vol, 128)
        assert got.shape == (len(want), 128, 128) and len(want) == 12            # 20 slices -> [4, 16)
        assert all(np.array_equal(got[k], want[k]) for k in range(len(want)))


def test_lung_boxes_match_cropper():
    _, lung, _ = synthetic_case()
    lung[:, :, 6] = 0                                                # a constant slice inside the window is skipped
    sl = N.volume_slices(lung, 256)
    kept, boxes = N.lung_boxes(sl)
    assert 2 not in kept and len(kept) == 11 and boxes.shape == (11, 8)
    for k, bx in zip(kept, boxes):
        img = sl[k].copy()
        img[img > 0] = 1
        test_img = np.uint8(img * 255)                               # T1H:213-214
        contours, _ = cv2.findContours(test_img, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
        areas = [cv2.contourArea(c) for c in contours]
        order = np.argsort(areas)
        want = list(cv2.boundingRect(contours[order[-1]])) + list(cv2.boundingRect(contours[order[-2]]))
        assert bx.tolist() == want
