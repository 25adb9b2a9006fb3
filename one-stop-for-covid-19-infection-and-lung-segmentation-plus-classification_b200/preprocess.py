"""Device preprocessing behind the reference's function names.

  clahe_enhancer(img)            T1H:163-194  cv2.createCLAHE(3.0,(8,8)).apply          -> b2u_clahe_u8
  cropper(mask) -> boxes         T1H:211-233  cv2.findContours + two largest boundingRects (host, cv2)
  crop_resize(imgs, boxes)       T1H:236-241, 347-368, 485-488, 678-686                 -> b2u_crop_resize
      crop each lung box, resize to (125 x 250) INTER_AREA, hconcat -> 250 x 250, resize to new_dim INTER_LINEAR,
      uint8, /255 -> float32 (N, new_dim, new_dim, 1)
  resize(imgs, dsize, interp)    cv2.resize(img, dsize, interpolation=cv2.INTER_AREA | cv2.INTER_LINEAR), uint8  -> b2u_resize_u8

The resize kernels restate OpenCV's CV_8UC1 arithmetic exactly (area tables with float accumulation in table order, 11-bit
fixed-point bilinear): their outputs are np.array_equal to cv2's (tests/test_gpu_preprocess.py, oracle/cv_resize.py).

Contour tracing (cv2.findContours, RETR_TREE + contourArea ranking) is serial border following and stays on
the host exactly as the reference calls it (polygon area != pixel count, so a connected-component
restatement could rank contours differently); only the per-pixel work runs on the GPU.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def clahe_enhancer(test_img, demo=0, clip_limit=3.0, tiles=8):
    """(H,W) or (N,H,W) float images in [0,1] -> uint8 CLAHE-enhanced images (same leading shape)."""
    a = np.asarray(test_img)
    single = a.ndim == 2
    if single:
        a = a[None]
    u8 = np.ascontiguousarray(np.uint8(a * 255)) if a.dtype != np.uint8 else np.ascontiguousarray(a)
    n, h, w = u8.shape
    l = _lib.lib()
    d_in = torch.from_numpy(u8).cuda()
    d_out = torch.empty_like(d_in)
    ws = torch.empty(n * tiles * tiles * 256, dtype=torch.uint8, device="cuda")
    _lib.check(l.b2u_clahe_u8(C.c_void_p(d_in.data_ptr()), C.c_void_p(d_out.data_ptr()), n, h, w, float(clip_limit),
                              int(tiles), C.c_void_p(ws.data_ptr()), ws.numel(), _stream_ptr()), "clahe_u8")
    out = d_out.cpu().numpy()
    return out[0] if single else out


def cropper_boxes(lung_mask_u8):
    """the two largest-area contours' bounding boxes of a binary uint8 mask, as the reference's cropper()
    computes them (T1H:219-233). Returns [x,y,w,h, p,q,r,s]."""
    import cv2
    contours, _ = cv2.findContours(lung_mask_u8, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
    if len(contours) < 2:
        raise ValueError("cropper needs at least two contours (got %d)" % len(contours))
    areas = [cv2.contourArea(c) for c in contours]
    order = np.argsort(areas)
    b1 = cv2.boundingRect(contours[order[-1]])
    b2 = cv2.boundingRect(contours[order[-2]])
    return list(b1) + list(b2)


def crop_resize(images_u8, boxes, half_w=125, out_h=250, new_dim=224):
    """images (N,H,W) uint8, boxes (N,8) int -> float32 (N,new_dim,new_dim,1) in [0,1] and the uint8 250x250 stage."""
    u8 = np.ascontiguousarray(images_u8, dtype=np.uint8)
    bx = np.ascontiguousarray(boxes, dtype=np.int32).reshape(len(u8), 8)
    n, h, w = u8.shape
    l = _lib.lib()
    d_in, d_bx = torch.from_numpy(u8).cuda(), torch.from_numpy(bx).cuda()
    mid = torch.empty(n, out_h, 2 * half_w, dtype=torch.uint8, device="cuda")
    out = torch.empty(n, new_dim, new_dim, dtype=torch.float32, device="cuda")
    _lib.check(l.b2u_crop_resize(C.c_void_p(d_in.data_ptr()), n, h, w, C.c_void_p(d_bx.data_ptr()), half_w, out_h, new_dim,
                                 C.c_void_p(mid.data_ptr()), C.c_void_p(out.data_ptr()), _stream_ptr()), "crop_resize")
    return out.cpu().numpy()[..., None], mid.cpu().numpy()


INTER_LINEAR, INTER_AREA = 1, 3           # OpenCV's enum values


def resize(images_u8, dsize, interpolation=INTER_AREA):
    """cv2.resize(img, dsize=(width, height), interpolation=...) for (H,W) or (N,H,W) uint8 images, on the GPU."""
    a = np.asarray(images_u8)
    if a.dtype != np.uint8:
        raise ValueError("resize: uint8 images only (got %s)" % a.dtype)
    single = a.ndim == 2
    if single:
        a = a[None]
    a = np.ascontiguousarray(a)
    n, h, w = a.shape
    dw, dh = int(dsize[0]), int(dsize[1])
    d_in = torch.from_numpy(a).cuda()
    d_out = torch.empty(n, dh, dw, dtype=torch.uint8, device="cuda")
    _lib.check(_lib.lib().b2u_resize_u8(C.c_void_p(d_in.data_ptr()), n, h, w, C.c_void_p(d_out.data_ptr()), dh, dw,
                                        int(interpolation), _stream_ptr()), "resize_u8")
    out = d_out.cpu().numpy()
    return out[0] if single else out


def resize_area_normalize(slices_f64, img_size=512, normalize=True):
    """read_nii's first stage on the GPU (T1H:335-337): cv2.resize(img, (img_size, img_size), INTER_AREA) on float64
    slices (S,H,W), then (img - min) / (max - min) per slice.  Bit-exact against cv2 + numpy; returns float64."""
    a = np.ascontiguousarray(slices_f64, dtype=np.float64)
    if a.ndim != 3:
        raise ValueError("resize_area_normalize: (S,H,W) slices expected")
    n, h, w = a.shape
    d_in = torch.from_numpy(a).cuda()
    d_out = torch.empty(n, img_size, img_size, dtype=torch.float64, device="cuda")
    _lib.check(_lib.lib().b2u_resize_area_f64(C.c_void_p(d_in.data_ptr()), n, h, w, C.c_void_p(d_out.data_ptr()), img_size,
                                              img_size, 1 if normalize else 0, _stream_ptr()), "resize_area_f64")
    return d_out.cpu().numpy()
