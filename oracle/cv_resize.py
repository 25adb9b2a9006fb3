"""TEST INFRASTRUCTURE (oracle) -- OpenCV's uint8 resize arithmetic restated in numpy, bit for bit.

The reference resizes with cv2 (OpenCV, an un-vendored dependency; 4.1.x on Colab in 2020, 4.13 here):
  * crops -> (125, 250) with INTER_AREA   /root/reference/Scripts/task1_preprocessing_plus_unet_with_comments.py:236-241, 355-358
  * 250 x 250 -> new_dim x new_dim with INTER_LINEAR                                                          :485-488
OpenCV's algorithms for CV_8UC1 (modules/imgproc/src/resize.cpp) are restated here so that the CUDA kernels
(csrc/preprocess.cu) can be held to np.array_equal:
  * INTER_AREA, both scales >= 1, both integer  -> resizeAreaFast_: 2x2 is (a+b+c+d+2)>>2, else cvRound(sum * (1.f/area))
  * INTER_AREA, both scales >= 1, otherwise      -> resizeArea_ with DecimateAlpha tables (float accumulation, table order)
  * INTER_AREA with a scale < 1 (up-sampling)    -> the INTER_LINEAR fixed-point path with area-style coefficients
  * INTER_LINEAR                                  -> 11-bit coefficients (saturate_cast<short>(f * 2048)), horizontal pass
                                                    in int, vertical ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2 >> 2
This file is PINNED against the real cv2 in tests/test_oracle_cv_resize.py (cv2 is importable on the build and GPU
boxes), i.e. unlike the Keras oracle its parity is pinned.  Only tests/ may import it.
"""
import math

import numpy as np

F = np.float32


def _rint(v):
    """cvRound: round half to even"""
    return int(np.rint(v))


def _sat_u8(i):
    return 0 if i < 0 else (255 if i > 255 else i)


def _area_tab(ssize, dsize, scale):
    """computeResizeAreaTab: per destination index the list of (source index, float32 alpha), in table order"""
    tab = []
    for dx in range(dsize):
        fsx1 = dx * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1, sx2 = math.ceil(fsx1), math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        ent = []
        if sx1 - fsx1 > 1e-3:
            ent.append((sx1 - 1, F((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            ent.append((sx, F(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            ent.append((sx2, F(min(min(fsx2 - sx2, 1.0), cell) / cell)))
        tab.append(ent)
    return tab


def _linear_coeffs(ssize, dsize, area_mode):
    """xofs / ialpha of cv::resize's generic path (ksize = 2), 11-bit fixed point"""
    inv_scale = dsize / ssize
    scale = 1.0 / inv_scale
    ofs, a0, a1 = [], [], []
    for dx in range(dsize):
        if not area_mode:
            fx = F((dx + 0.5) * scale - 0.5)
            sx = int(math.floor(fx))
            fx = F(fx - F(sx))
        else:
            sx = int(math.floor(dx * scale))
            fx = F((dx + 1) - (sx + 1) * inv_scale)
            fx = F(0) if fx <= 0 else F(fx - F(math.floor(fx)))
        ofs.append(sx)
        a0.append(_rint(F(F(1.0) - fx) * F(2048)))
        a1.append(_rint(fx * F(2048)))
    return ofs, a0, a1


def _linear_fixed(src, dw, dh, area_mode):
    sh, sw = src.shape
    xo, xa0, xa1 = _linear_coeffs(sw, dw, area_mode)
    yo, ya0, ya1 = _linear_coeffs(sh, dh, area_mode)
    s = src.astype(np.int64)
    rows = np.empty((sh, dw), np.int64)
    for dx in range(dw):
        sx = xo[dx]
        if sx < 0:                       # (sx < 0: fx = 0, sx = 0)
            rows[:, dx] = s[:, 0] * 2048
        elif sx >= sw - 1:               # (dx >= xmax: D = S[xofs] * ONE)
            rows[:, dx] = s[:, sw - 1] * 2048
        else:
            rows[:, dx] = s[:, sx] * xa0[dx] + s[:, sx + 1] * xa1[dx]
    out = np.empty((dh, dw), np.uint8)
    for dy in range(dh):
        r0 = rows[min(max(yo[dy], 0), sh - 1)]
        r1 = rows[min(max(yo[dy] + 1, 0), sh - 1)]
        v = (((ya0[dy] * (r0 >> 4)) >> 16) + ((ya1[dy] * (r1 >> 4)) >> 16) + 2) >> 2
        out[dy] = np.clip(v, 0, 255).astype(np.uint8)      # (uchar(...) of a value that is in range by construction)
    return out


def resize_linear_u8(src, dw, dh):
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR) for a 2-D uint8 image"""
    return _linear_fixed(np.asarray(src, np.uint8), dw, dh, False)


def resize_area_u8(src, dw, dh):
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_AREA) for a 2-D uint8 image"""
    src = np.asarray(src, np.uint8)
    sh, sw = src.shape
    scale_x, scale_y = 1.0 / (dw / sw), 1.0 / (dh / sh)
    if not (scale_x >= 1 and scale_y >= 1):
        return _linear_fixed(src, dw, dh, True)
    ix, iy = _rint(scale_x), _rint(scale_y)
    eps = np.finfo(np.float64).eps
    out = np.empty((dh, dw), np.uint8)
    if abs(scale_x - ix) < eps and abs(scale_y - iy) < eps:
        s = src.astype(np.int64)[:dh * iy, :dw * ix].reshape(dh, iy, dw, ix).sum(axis=(1, 3))
        if ix == 2 and iy == 2:
            return ((s + 2) >> 2).astype(np.uint8)
        sc = F(1.0) / F(ix * iy)
        for dy in range(dh):
            for dx in range(dw):
                out[dy, dx] = _sat_u8(_rint(F(s[dy, dx]) * sc))
        return out
    xtab, ytab = _area_tab(sw, dw, scale_x), _area_tab(sh, dh, scale_y)
    sf = src.astype(np.float32)
    for dy in range(dh):
        acc = None
        for sy, beta in ytab[dy]:
            buf = np.zeros(dw, np.float32)
            for dx in range(dw):
                b = F(0)
                for sx, alpha in xtab[dx]:
                    b = F(b + F(sf[sy, sx] * alpha))
                buf[dx] = b
            acc = (beta * buf).astype(np.float32) if acc is None else (acc + (beta * buf).astype(np.float32)).astype(np.float32)
        out[dy] = np.clip(np.rint(acc), 0, 255).astype(np.uint8)
    return out


# ---- CV_64F INTER_AREA (the 512 x 512 stage of the NIfTI ingest works on get_fdata() float64 slices, T1H:335) ------
def resize_area_f64(src, dw, dh):
    """cv2.resize(src, (dw, dh), interpolation=cv2.INTER_AREA) for a 2-D float64 image: the same three regimes as uint8
    with double accumulators (WT = double), float32 table / interpolation coefficients, and no rounding at the end."""
    src = np.asarray(src, np.float64)
    sh, sw = src.shape
    if (sh, sw) == (dh, dw):
        return src.copy()                                   # cv::resize: same size -> plain copy
    inv_x, inv_y = dw / sw, dh / sh
    scale_x, scale_y = 1.0 / inv_x, 1.0 / inv_y
    out = np.empty((dh, dw), np.float64)
    if scale_x >= 1 and scale_y >= 1:
        ix, iy = _rint(scale_x), _rint(scale_y)
        eps = np.finfo(np.float64).eps
        if abs(scale_x - ix) < eps and abs(scale_y - iy) < eps:
            sc = float(F(1.0) / F(ix * iy))                 # float scale, promoted to double in sum * scale
            for dy in range(dh):
                for dx in range(dw):
                    # ofs[] order: row by row inside the cell; CV_ENABLE_UNROLLED adds the values in groups of four
                    vals = [src[dy * iy + yy, dx * ix + xx] for yy in range(iy) for xx in range(ix)]
                    s, k = 0.0, 0
                    while k <= len(vals) - 4:
                        s += ((vals[k] + vals[k + 1]) + vals[k + 2]) + vals[k + 3]
                        k += 4
                    for v in vals[k:]:
                        s += v
                    out[dy, dx] = s * sc
            return out
        xtab, ytab = _area_tab(sw, dw, scale_x), _area_tab(sh, dh, scale_y)
        for dy in range(dh):
            acc = None
            for sy, beta in ytab[dy]:
                buf = np.zeros(dw, np.float64)
                for dx in range(dw):
                    b = 0.0
                    for sx, alpha in xtab[dx]:
                        b = b + src[sy, sx] * float(alpha)
                    buf[dx] = b
                acc = float(beta) * buf if acc is None else acc + float(beta) * buf
            out[dy] = acc
        return out
    # up-sampling in a dimension: bilinear with float32 "area" coefficients, double arithmetic
    def coeffs(ssize, dsize, inv_scale, scale):
        res = []
        for d in range(dsize):
            s = int(math.floor(d * scale))
            f = F((d + 1) - (s + 1) * inv_scale)
            f = F(0) if f <= 0 else F(f - F(math.floor(f)))
            if s < 0:
                f, s = F(0), 0
            if s >= ssize - 1:
                f, s = F(0), ssize - 1
            res.append((s, float(F(1.0) - f), float(f)))
        return res
    cx, cy = coeffs(sw, dw, inv_x, scale_x), coeffs(sh, dh, inv_y, scale_y)
    rows = np.empty((sh, dw), np.float64)
    for dx, (sx, a0, a1) in enumerate(cx):
        rows[:, dx] = src[:, sx] * a0 + src[:, min(sx + 1, sw - 1)] * a1 if sx < sw - 1 else src[:, sx] * 1.0
    for dy, (sy, b0, b1) in enumerate(cy):
        out[dy] = rows[sy] * b0 + rows[min(sy + 1, sh - 1)] * b1
    return out
