"""NIfTI-1 ingest for the runners: what `nib.load(path).get_fdata()` + `read_nii` /
`read_nii_demo` do in the reference (task1_preprocessing_plus_unet_with_comments.py:281-297, 310-376), without nibabel
(absent offline) and with the per-pixel work of the CT / mask pipeline on the GPU kernels of preprocess.py.

  load_nii(path)                 -> float64 array, header dict      == nib.load(path).get_fdata()   (T1H:286-287)
  volume_slices(array)           -> (S, 512, 512) float64 in [0,1]  rot90, middle 60 % of the slices, 512^2 INTER_AREA
                                                                     resize, per-slice min-max       (T1H:288-297)
  lung_boxes(lung_slices)        -> (kept slice indices, boxes)      binarise + cropper()            (T1H:331-345)
  preprocess_case(ct, lung, inf) -> x (N,new_dim,new_dim,1), y (same), boxes                          (T1H:347-368, 485-488)

File format (NIfTI-1.1, single-file `.nii` / `.nii.gz`): 348-byte header, `sizeof_hdr` = 348 in the file's byte order,
`dim[8]` int16 at 40, `datatype` / `bitpix` int16 at 70 / 72, `vox_offset` / `scl_slope` / `scl_inter` float32 at
108 / 112 / 116, magic "n+1\\0" at 344; voxels follow at `vox_offset` in Fortran order (first index fastest).
`get_fdata()` semantics: float64, `raw * scl_slope + scl_inter` when the slope is finite and non-zero.
"""
import gzip
import struct

import numpy as np

_DTYPES = {2: "u1", 4: "i2", 8: "i4", 16: "f4", 64: "f8", 256: "i1", 512: "u2", 768: "u4", 1024: "i8", 1280: "u8"}
IMG_SIZE = 512          # T1H:151  img_size


def _read_all(path):
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    return raw


def load_nii(path):
    """Reads a single-file NIfTI-1 image; returns (float64 ndarray shaped dim[1..dim[0]], header dict)."""
    raw = _read_all(path)
    if len(raw) < 348:
        raise ValueError("%s: not a NIfTI file (%d bytes)" % (path, len(raw)))
    if struct.unpack("<i", raw[:4])[0] == 348:
        bo = "<"
    elif struct.unpack(">i", raw[:4])[0] == 348:
        bo = ">"
    else:
        raise ValueError("%s: sizeof_hdr is not 348 (NIfTI-2 / Analyze pairs are not supported)" % path)
    magic = raw[344:348]
    if magic[:3] != b"n+1":
        raise ValueError("%s: magic %r is not a single-file NIfTI-1 ('n+1')" % (path, magic))
    dim = struct.unpack(bo + "8h", raw[40:56])
    ndim = dim[0]
    if not 1 <= ndim <= 7:
        raise ValueError("%s: bad dim[0] = %d" % (path, ndim))
    shape = tuple(int(d) for d in dim[1:1 + ndim])
    datatype, bitpix = struct.unpack(bo + "2h", raw[70:74])
    if datatype not in _DTYPES:
        raise ValueError("%s: unsupported NIfTI datatype %d" % (path, datatype))
    pixdim = struct.unpack(bo + "8f", raw[76:108])
    vox_offset, slope, inter = struct.unpack(bo + "3f", raw[108:120])
    dt = np.dtype(bo + _DTYPES[datatype])
    if dt.itemsize * 8 != bitpix:
        raise ValueError("%s: bitpix %d does not match datatype %d" % (path, bitpix, datatype))
    off = int(vox_offset) if vox_offset >= 352 else 352
    count = int(np.prod(shape))
    if len(raw) < off + count * dt.itemsize:
        raise ValueError("%s: truncated (need %d voxel bytes, have %d)" % (path, count * dt.itemsize, len(raw) - off))
    data = np.frombuffer(raw, dtype=dt, count=count, offset=off).reshape(shape, order="F")
    out = data.astype(np.float64)
    if np.isfinite(slope) and slope != 0.0 and not (slope == 1.0 and inter == 0.0):
        out = out * float(slope) + float(inter)
    return out, dict(shape=shape, datatype=datatype, pixdim=pixdim[1:1 + ndim], scl_slope=slope, scl_inter=inter,
                     byteorder=bo, vox_offset=off)


def save_nii(path, array, scl_slope=0.0, scl_inter=0.0, byteorder="<"):
    """Minimal single-file NIfTI-1 writer (tests and synthetic fixtures): `array` dtype decides the datatype."""
    a = np.asarray(array)
    code = {v: k for k, v in _DTYPES.items()}.get(a.dtype.str[1:])
    if code is None:
        raise ValueError("unsupported dtype %s" % a.dtype)
    hdr = bytearray(352)
    struct.pack_into(byteorder + "i", hdr, 0, 348)
    dim = [a.ndim] + list(a.shape) + [1] * (7 - a.ndim)
    struct.pack_into(byteorder + "8h", hdr, 40, *dim)
    struct.pack_into(byteorder + "2h", hdr, 70, code, a.dtype.itemsize * 8)
    struct.pack_into(byteorder + "8f", hdr, 76, 1.0, *([1.0] * 7))
    struct.pack_into(byteorder + "3f", hdr, 108, 352.0, scl_slope, scl_inter)
    hdr[344:348] = b"n+1\x00"
    body = a.astype(a.dtype.newbyteorder(byteorder)).tobytes(order="F")
    blob = bytes(hdr) + body
    if str(path).endswith(".gz"):
        blob = gzip.compress(blob, compresslevel=1)
    with open(path, "wb") as f:
        f.write(blob)


def volume_slices(array, img_size=IMG_SIZE, device=False):
    """T1H:288-297: rot90, keep slices [round(0.2 S), round(0.8 S)), resize each to img_size^2 with INTER_AREA and
    min-max normalise it.  Returns (S', img_size, img_size) float64; constant slices come out as NaN exactly like the
    reference's 0/0 (read_nii skips them for the lung volume before this point matters).
    device=True runs the resize + normalisation on the GPU (b2u_resize_area_f64, bit-exact against cv2 / numpy: this is
    what preprocess_case uses); device=False is the reference's own cv2 call sequence on the host."""
    a = np.rot90(np.array(array))
    s = a.shape[2]
    a = a[:, :, round(s * 0.2):round(s * 0.8)]
    if device:
        from .preprocess import resize_area_normalize
        return resize_area_normalize(np.ascontiguousarray(np.rollaxis(a, 2), dtype=np.float64), img_size)
    import cv2
    a = np.reshape(np.rollaxis(a, 2), (a.shape[2], a.shape[0], a.shape[1], 1))
    out = np.empty((a.shape[0], img_size, img_size), np.float64)
    for k in range(a.shape[0]):
        img = cv2.resize(a[k], dsize=(img_size, img_size), interpolation=cv2.INTER_AREA)
        xmax, xmin = img.max(), img.min()
        with np.errstate(invalid="ignore", divide="ignore"):
            out[k] = (img - xmin) / (xmax - xmin)
    return out


def lung_boxes(lung_slices):
    """T1H:331-345: constant lung slices are skipped, the others binarised (img > 0 -> 1) and passed to cropper();
    returns (indices of the kept slices, (n, 8) int array of [x, y, w, h, p, q, r, s]).  NOTE the reference appends the
    boxes to global lists and later indexes them with the CT slice number, so a skipped lung slice shifts every later
    box by one (T1H:347 `img_no < len(all_points1)`); `preprocess_case` keeps that indexing."""
    from .preprocess import cropper_boxes
    kept, boxes = [], []
    for k, img in enumerate(lung_slices):
        if not np.isfinite(img).all() or np.unique(img).size == 1:
            continue
        m = np.zeros(img.shape, np.uint8)
        m[img > 0] = 1
        boxes.append(cropper_boxes(m))
        kept.append(k)
    return np.asarray(kept, np.int64), np.asarray(boxes, np.int32).reshape(len(boxes), 8)


def preprocess_case(ct_path, lung_path, infection_path=None, new_dim=224, img_size=IMG_SIZE):
    """One patient, file to network input: what read_nii('lungs') / ('cts') / ('infections') followed by the final
    resize and /255 produce (T1H:390-393, 485-488, 678-686).  The 512^2 area resize + min-max of the raw float64 slices,
    CLAHE, crop, area / bilinear resize and scaling all run on the GPU (preprocess.py), bit-exact against the
    reference's cv2 / numpy calls; only the contour tracing of `cropper` (cv2.findContours) stays on the host.
    Returns (x, y, boxes): x, y float32 (N, new_dim, new_dim, 1); y is None without an infection mask."""
    from . import preprocess as PP
    lung, _ = load_nii(lung_path)
    ct, _ = load_nii(ct_path)
    _, boxes = lung_boxes(volume_slices(lung, img_size, device=True))
    cts = volume_slices(ct, img_size, device=True)
    n = min(len(cts), len(boxes))                               # T1H:347: slice i uses box i while i < len(boxes)
    enhanced = PP.clahe_enhancer(np.nan_to_num(cts[:n]))        # np.uint8(img * 255) -> CLAHE, uint8 (T1H:169-170)
    x, _ = PP.crop_resize(enhanced, boxes[:n], new_dim=new_dim)
    y = None
    if infection_path is not None:
        inf, _ = load_nii(infection_path)
        infs = np.uint8(np.nan_to_num(volume_slices(inf, img_size, device=True)[:n]) * 255)          # T1H:363
        y, _ = PP.crop_resize(infs, boxes[:n], new_dim=new_dim)
    return x, y, boxes[:n]
