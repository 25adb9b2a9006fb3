"""tcgen05 implicit-GEMM kernels (csrc/conv_tc.cu) against tests/emulator.py on U-Net-shaped problems:
all K-slab widths (16/32/64 channels -> 32B/64B/128B swizzle), several N tiles (J = 512), ragged
image sizes (tiles clipped by TMA zero-fill), concat-slice strides, mask / accumulate / BN-statistics
epilogues, and the transposed-conv scatter / sub-grid gather variants."""
import importlib

import numpy as np
import pytest

import emulator as E
from conftest import PKG
from helpers import P
from test_gpu_ops import Img, compare

pytestmark = pytest.mark.gpu
dt = P.F16


def test_tensor_path_is_live():
    lib = importlib.import_module(PKG + "._lib")
    assert lib.lib().b2u_tensor_path_available() == 1, "tcgen05 path not compiled in / not an sm_100 device"


TC_CONV = [  # n, h, w, cin, cout
    (1, 16, 16, 64, 64), (2, 32, 32, 32, 32), (1, 32, 32, 256, 512), (1, 16, 16, 512, 256), (2, 24, 40, 128, 128),
    (1, 14, 14, 256, 512), (1, 28, 28, 96, 32), (1, 56, 56, 16, 16), (3, 8, 8, 192, 64), (1, 64, 64, 64, 32),
    (1, 16, 16, 32, 80), (1, 16, 16, 32, 48), (1, 16, 16, 48, 32), (1, 16, 16, 80, 32),      # 16-channel K slabs, odd N tiles
]


@pytest.mark.parametrize("n,h,w,cin,cout", TC_CONV)
def test_tc_conv3x3_fwd(n, h, w, cin, cout):
    img = Img(21)
    x = img.view(n, h, w, cin, dt, ld=2 * cin, c0=cin, fill="uniform")          # right half of a concat buffer
    y = img.view(n, h, w, cout, dt, ld=cout + 16, c0=8, fill=None)
    wt = img.farr(img.par, 9 * cin * cout, scale=(2.0 / (9 * cin)) ** 0.5)
    b = img.farr(img.par, cout, scale=0.1)
    stats = img.zero.alloc(2 * cout * 8)
    for act in (1, 2):
        ops = [P.Op(P.OP_CONV3X3_FWD, dt, [x.ref, wt, b, y.ref, stats], [x.ld, cin, act, y.ld, cout, n, h, w])]
        compare(ops, img, dt, tol=3e-3)


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 32, 32, 16, 16), (1, 28, 28, 32, 32), (3, 14, 14, 64, 64), (1, 24, 40, 128, 256),
                                            (1, 16, 16, 512, 512), (2, 56, 56, 16, 32)])
def test_tc_conv3x3_fwd_with_post_activation_affine(n, h, w, cin, cout):
    """inference plans: Conv2D(relu | elu) -> BatchNormalization as one kernel, y = scale * act(conv + b) + shift (p[7], p[8]),
    here with the preceding BN finalize op producing scale / shift from moving statistics, strided input and output"""
    img = Img(121)
    x = img.view(n, h, w, cin, dt, ld=2 * cin, c0=cin, fill="uniform")
    y = img.view(n, h, w, cout, dt, ld=cout + 16, c0=8, fill=None)
    wt = img.farr(img.par, 9 * cin * cout, scale=(2.0 / (9 * cin)) ** 0.5)
    b = img.farr(img.par, cout, scale=0.1)
    gamma, beta = img.farr(img.par, cout, fill="pos"), img.farr(img.par, cout, scale=0.2)
    mm, mv = img.farr(img.par, cout, scale=0.2), img.farr(img.par, cout, fill="pos")
    sc, sh, mean, inv = (img.f32.alloc(cout * 4) for _ in range(4))
    for act in (1, 2):
        ops = [P.Op(P.OP_BN_FINALIZE, 0, [None, gamma, beta, mm, mv, sc, sh, mean, inv], [n * h * w, 0, cout], [0.99, 1e-3]),
               P.Op(P.OP_CONV3X3_FWD, dt, [x.ref, wt, b, y.ref, None, None, None, sc, sh],
                    [x.ld, cin, act, y.ld, cout, n, h, w, 0, 0])]
        compare(ops, img, dt, tol=3e-3)


@pytest.mark.parametrize("n,h,w,cin,cout", TC_CONV)
def test_tc_conv3x3_dgrad(n, h, w, cin, cout):
    img = Img(22)
    dy = img.view(n, h, w, cout, dt, ld=cout + 8, scale=0.5)
    dx = img.view(n, h, w, cin, dt, ld=2 * cin, c0=0, scale=0.3)
    mask = img.view(n, h, w, cin, dt)
    wt = img.farr(img.par, 9 * cin * cout, scale=(2.0 / (9 * cin)) ** 0.5)
    for acc, mact in ((0, 1), (1, 2), (0, 0)):
        ops = [P.Op(P.OP_CONV3X3_DGRAD, dt, [dy.ref, wt, dx.ref, mask.ref if mact else None],
                    [dy.ld, cout, dx.ld, cin, mask.ld, mact, acc, n, h, w])]
        compare(ops, img, dt, tol=4e-3)


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 8, 8, 64, 32), (1, 5, 7, 32, 16), (1, 4, 4, 512, 256), (1, 32, 32, 128, 64),
                                            (2, 16, 16, 256, 128)])
def test_tc_convt(n, h, w, cin, cout):
    img = Img(23)
    x = img.view(n, h, w, cin, dt, fill="uniform")
    y = img.view(n, 2 * h, 2 * w, cout, dt, ld=2 * cout, c0=0, fill=None)
    dy = img.view(n, 2 * h, 2 * w, cout, dt, ld=2 * cout, c0=cout, scale=0.5)
    dx = img.view(n, h, w, cin, dt, scale=0.2)
    wt = img.farr(img.par, 4 * cout * cin, scale=(1.0 / cin) ** 0.5)
    b = img.farr(img.par, cout, scale=0.1)
    for acc in (0, 1):
        ops = [P.Op(P.OP_CONVT_FWD, dt, [x.ref, wt, b, y.ref, None], [x.ld, cin, y.ld, cout, n, h, w, 0]),
               P.Op(P.OP_CONVT_DGRAD, dt, [dy.ref, wt, dx.ref, x.ref], [dy.ld, cout, dx.ld, cin, x.ld, 1, acc, n, h, w])]
        compare(ops, img, dt, tol=4e-3)


def test_tc_streamed_weights_two_subtiles_many_tiles():
    """wide layer with streamed weights: 32 x 8 CTA tiles (two sub-tiles per weight tile), more tiles than SMs, both
    accumulator stages, forward with BN statistics and data gradient with mask + column sums"""
    n, h, w, cin, cout = 4, 128, 128, 128, 128
    img = Img(25)
    x = img.view(n, h, w, cin, dt, fill="uniform")
    y = img.view(n, h, w, cout, dt, ld=2 * cout, c0=cout, fill=None)
    wt = img.farr(img.par, 9 * cin * cout, scale=(2.0 / (9 * cin)) ** 0.5)
    b = img.farr(img.par, cout, scale=0.1)
    stats = img.zero.alloc(2 * cout * 8)
    ops = [P.Op(P.OP_CONV3X3_FWD, dt, [x.ref, wt, b, y.ref, stats], [x.ld, cin, 1, y.ld, cout, n, h, w])]
    compare(ops, img, dt, tol=3e-3)
    img = Img(26)
    dy = img.view(n, h, w, cout, dt, scale=0.5)
    dx = img.view(n, h, w, cin, dt, fill=None)
    mask = img.view(n, h, w, cin, dt)
    wt = img.farr(img.par, 9 * cin * cout, scale=(2.0 / (9 * cin)) ** 0.5)
    db = img.farr(img.gr, cin, scale=0.01)
    ops = [P.Op(P.OP_CONV3X3_DGRAD, dt, [dy.ref, wt, dx.ref, mask.ref, db], [dy.ld, cout, dx.ld, cin, mask.ld, 1, 0, n, h, w])]
    compare(ops, img, dt, tol=4e-3)


def test_tc_large_layer_many_tiles():
    """more tiles than SMs: exercises the persistent loop, the smem ring wrap-around and both TMEM stages"""
    n, h, w, cin, cout = 2, 128, 128, 32, 32
    img = Img(24)
    x = img.view(n, h, w, cin, dt, fill="uniform")
    y = img.view(n, h, w, cout, dt, fill=None)
    wt = img.farr(img.par, 9 * cin * cout, scale=(2.0 / (9 * cin)) ** 0.5)
    b = img.farr(img.par, cout, scale=0.1)
    stats = img.zero.alloc(2 * cout * 8)
    ops = [P.Op(P.OP_CONV3X3_FWD, dt, [x.ref, wt, b, y.ref, stats], [x.ld, cin, 1, y.ld, cout, n, h, w])]
    compare(ops, img, dt, tol=3e-3)


TC_WGRAD = [  # n, h, w, cin, cout
    (2, 16, 16, 32, 32), (1, 32, 32, 64, 64), (2, 16, 32, 64, 128), (1, 16, 16, 128, 128), (1, 24, 40, 128, 256),
    (1, 8, 8, 512, 512), (1, 16, 16, 512, 256), (1, 20, 12, 32, 64), (2, 40, 24, 16, 16), (1, 14, 14, 256, 512),
    (2, 64, 64, 32, 32),
]


@pytest.mark.parametrize("n,h,w,cin,cout", TC_WGRAD)
def test_tc_conv3x3_wgrad(n, h, w, cin, cout):
    img = Img(31)
    x = img.view(n, h, w, cin, dt, ld=2 * cin, c0=cin, fill="uniform")
    dy = img.view(n, h, w, cout, dt, ld=cout + 8, scale=0.5)
    dw = img.farr(img.gr, 9 * cin * cout, scale=0.01)
    db = img.farr(img.gr, cout, scale=0.01)
    ops = [P.Op(P.OP_CONV3X3_WGRAD, dt, [x.ref, dy.ref, dw, db], [x.ld, cin, dy.ld, cout, n, h, w])]
    compare(ops, img, dt, tol=3e-3)


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 8, 8, 64, 32), (1, 16, 16, 128, 64), (1, 4, 4, 512, 256), (1, 32, 32, 256, 128),
                                            (2, 10, 6, 64, 32)])
def test_tc_convt_wgrad(n, h, w, cin, cout):
    img = Img(32)
    x = img.view(n, h, w, cin, dt, fill="uniform")
    dy = img.view(n, 2 * h, 2 * w, cout, dt, ld=2 * cout, c0=0, scale=0.5)
    dw = img.farr(img.gr, 4 * cout * cin, scale=0.01)
    db = img.farr(img.gr, cout, scale=0.01)
    ops = [P.Op(P.OP_CONVT_WGRAD, dt, [x.ref, dy.ref, dw, db], [x.ld, cin, dy.ld, cout, n, h, w])]
    compare(ops, img, dt, tol=3e-3)


@pytest.mark.parametrize("dhm", [1, 0])
@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 16, 16, 32, 32), (1, 32, 32, 64, 32), (1, 32, 32, 64, 64), (1, 20, 12, 32, 64),
                                            (3, 21, 13, 32, 32), (1, 40, 72, 64, 32), (1, 32, 32, 128, 64), (1, 16, 24, 96, 32),
                                            (2, 17, 9, 160, 64), (4, 128, 128, 32, 32), (2, 64, 64, 64, 64)])
def test_tc_conv3x3_wgrad_thin_outputs_dh_merged(n, h, w, cin, cout, dhm):
    """Cout = 32 / 64: the three tap rows merged in N (one MMA per K step against a dy tile with a row halo, wgrad_tc.cu
    WhParams::dhm) against the per-row accumulators; ragged tiles (TMA zero fill is the padding), several images, 32- and
    64-channel slabs, concat-slice strides; integer data so that both variants and the emulator agree to fp32 summation
    order (tolerance 1e-5 instead of the fp16-product tolerance)."""
    lib = importlib.import_module(PKG + "._lib").lib()
    old = lib.b2u_set_option(b"wgrad_dhm", dhm)
    try:
        img = Img(131)
        x = img.view(n, h, w, cin, dt, ld=2 * cin, c0=cin, fill="int")
        dy = img.view(n, h, w, cout, dt, ld=cout + 8, fill="int", scale=0.25)
        dw = img.farr(img.gr, 9 * cin * cout, scale=0.01)
        db = img.farr(img.gr, cout, scale=0.01)
        ops = [P.Op(P.OP_CONV3X3_WGRAD, dt, [x.ref, dy.ref, dw, db], [x.ld, cin, dy.ld, cout, n, h, w])]
        compare(ops, img, dt, tol=1e-5)
    finally:
        lib.b2u_set_option(b"wgrad_dhm", old)


@pytest.mark.parametrize("n,h,w,cin,cout", [(1, 32, 32, 96, 32), (1, 16, 24, 192, 64), (2, 16, 16, 128, 32)])
def test_tc_wgrad_unetpp_concat_widths(n, h, w, cin, cout):
    """U-Net++ level-1/2 'a' convs see 96- and 192-channel concat inputs (UPP:905, 919)"""
    img = Img(33)
    x = img.view(n, h, w, cin, dt, fill="uniform")
    dy = img.view(n, h, w, cout, dt, scale=0.5)
    dw = img.farr(img.gr, 9 * cin * cout, scale=0.01)
    db = img.farr(img.gr, cout, scale=0.01)
    ops = [P.Op(P.OP_CONV3X3_WGRAD, dt, [x.ref, dy.ref, dw, db], [x.ld, cin, dy.ld, cout, n, h, w])]
    compare(ops, img, dt, tol=3e-3)


# ---- producer-side bias gradients: `colsum` outputs (trailing optional pointers of the backward ops) -----------
COLSUM_CONV = [  # n, h, w, cin (columns written), cout (reduction)
    (2, 32, 32, 32, 32), (1, 32, 40, 64, 32), (1, 16, 16, 64, 64), (2, 24, 40, 128, 128), (1, 16, 16, 256, 512),
    (1, 14, 14, 512, 256), (1, 56, 56, 16, 16),
]


@pytest.mark.parametrize("n,h,w,cin,cout", COLSUM_CONV)
def test_tc_conv3x3_dgrad_colsum(n, h, w, cin, cout):
    img = Img(41)
    dy = img.view(n, h, w, cout, dt, ld=cout + 8, scale=0.5)
    dx = img.view(n, h, w, cin, dt, ld=2 * cin, c0=cin, fill=None)
    mask = img.view(n, h, w, cin, dt)
    wt = img.farr(img.par, 9 * cin * cout, scale=(2.0 / (9 * cin)) ** 0.5)
    db = img.farr(img.gr, cin, scale=0.01)
    ops = [P.Op(P.OP_CONV3X3_DGRAD, dt, [dy.ref, wt, dx.ref, mask.ref, db],
                [dy.ld, cout, dx.ld, cin, mask.ld, 1, 0, n, h, w])]
    compare(ops, img, dt, tol=4e-3)


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 8, 8, 64, 32), (1, 32, 32, 128, 64), (2, 16, 16, 256, 128), (1, 8, 8, 512, 256),
                                            (1, 10, 6, 32, 16)])
def test_tc_convt_dgrad_colsum(n, h, w, cin, cout):
    img = Img(42)
    x = img.view(n, h, w, cin, dt, fill="normal")
    dy = img.view(n, 2 * h, 2 * w, cout, dt, ld=2 * cout, c0=cout, scale=0.5)
    dx = img.view(n, h, w, cin, dt, fill=None)
    wt = img.farr(img.par, 4 * cout * cin, scale=(1.0 / cin) ** 0.5)
    db = img.farr(img.gr, cin, scale=0.01)
    ops = [P.Op(P.OP_CONVT_DGRAD, dt, [dy.ref, wt, dx.ref, x.ref, db], [dy.ld, cout, dx.ld, cin, x.ld, 1, 0, n, h, w])]
    compare(ops, img, dt, tol=4e-3)


@pytest.mark.parametrize("c,npix", [(32, 5000), (64, 1111), (256, 300)])
def test_bn_bwd_apply_colsum(c, npix):
    img = Img(43)
    dy = img.view(1, 1, npix, c, dt, scale=0.5)
    x = img.view(1, 1, npix, c, dt, ld=2 * c, c0=c)
    dx = img.view(1, 1, npix, c, dt, fill=None)
    mask = img.view(1, 1, npix, c, dt)
    gamma = img.farr(img.par, c, fill="pos")
    mean = img.farr(img.f32, c, scale=0.1)
    inv = img.farr(img.f32, c, fill="pos")
    sums = img.farr(img.zero, 2 * c, scale=3.0, dtype=np.float64)
    dgamma, dbeta, db = img.farr(img.gr, c, fill=None), img.farr(img.gr, c, fill=None), img.farr(img.gr, c, scale=0.01)
    ops = [P.Op(P.OP_BN_BWD_APPLY, dt, [dy.ref, x.ref, dx.ref, gamma, mean, inv, sums, dgamma, dbeta, mask.ref, db],
                [dy.ld, x.ld, dx.ld, c, npix, mask.ld, 1, npix])]
    compare(ops, img, dt, tol=4e-3)


@pytest.mark.parametrize("c,npix", [(32, 4096), (16, 999), (64, 2500)])
def test_head_bwd_colsum(c, npix):
    img = Img(44)
    x = img.view(1, 1, npix, c, dt, fill="uniform")
    dx = img.view(1, 1, npix, c, dt, ld=c + 8, fill=None)
    wt, b = img.farr(img.par, c, scale=0.5), img.farr(img.par, 1, scale=0.1)
    prob = img.f32.alloc(npix * 4)
    tgt = img.farr(img.f32, npix, fill="uniform")
    out = img.f32.alloc(16)
    sums = img.zero.alloc(32)
    dw, db1, db = img.farr(img.gr, c, fill="zero"), img.farr(img.gr, 4, fill="zero"), img.farr(img.gr, c, scale=0.01)
    step = P.Ref("step", 0)
    ops = [P.Op(P.OP_HEAD_FWD, dt, [x.ref, wt, b, prob], [x.ld, c, npix]),
           P.Op(P.OP_BCE_DICE_SUMS, 0, [prob, tgt, sums], [npix]),
           P.Op(P.OP_BCE_DICE_FINALIZE, 0, [sums, out], [npix]),
           P.Op(P.OP_HEAD_BWD, dt, [prob, tgt, sums, step, x.ref, wt, dx.ref, dw, db1, db], [npix, x.ld, c, dx.ld, 1, npix])]
    compare(ops, img, dt, state=dict(loss_scale=256.0), tol=4e-3)


@pytest.mark.parametrize("cin,cout", [(64, 32), (80, 48)])
def test_pack_weights_table_matches_per_call_packing(cin, cout):
    """OP_PACK_WEIGHTS (one launch for all layers) + packed-weight pointers give bit-identical results to the
    per-call packing for every mode: conv fwd / dgrad, transposed-conv fwd / dgrad"""
    from gpu_harness import run_ops_gpu
    n, h, w = 1, 16, 16                       # (80, 48): not multiples of 32 -> ragged pack tiles
    img = Img(51)
    x = img.view(n, h, w, cin, dt, fill="uniform")
    y = img.view(n, h, w, cout, dt, fill=None)
    dy = img.view(n, h, w, cout, dt, scale=0.5)
    dx = img.view(n, h, w, cin, dt, fill=None)
    up = img.view(n, 2 * h, 2 * w, cout, dt, fill=None)
    dup = img.view(n, 2 * h, 2 * w, cout, dt, scale=0.5)
    dxt = img.view(n, h, w, cin, dt, fill=None)
    wt = img.farr(img.par, 9 * cin * cout, scale=0.05)
    b = img.farr(img.par, cout, scale=0.1)
    wtt = img.farr(img.par, 4 * cout * cin, scale=0.1)
    entries = [(wt.off // 4, 0, 9, cout, cin), (wt.off // 4, 1, 9, cin, cout), (wtt.off // 4, 2, 1, 4 * cout, cin),
               (wtt.off // 4, 3, 4, cin, cout)]
    tab = np.zeros((len(entries), 8), np.int64)
    refs, start, dst = [], 0, 0
    for r, (src, mode, taps, j, k) in enumerate(entries):
        tab[r] = (src, dst, start, mode, taps, j, k, 0)
        refs.append(P.Ref("wpack", dst * 2))
        start += taps * ((j + 31) // 32) * ((k + 31) // 32)
        dst += (taps * j * k + 127) // 128 * 128
    def ops(packed):
        pk = refs if packed else [None] * 4
        lst = [P.Op(P.OP_CONV3X3_FWD, dt, [x.ref, wt, b, y.ref, None, pk[0]], [x.ld, cin, 1, y.ld, cout, n, h, w]),
               P.Op(P.OP_CONV3X3_DGRAD, dt, [dy.ref, wt, dx.ref, None, None, pk[1]], [dy.ld, cout, dx.ld, cin, 0, 0, 0, n, h, w]),
               P.Op(P.OP_CONVT_FWD, dt, [x.ref, wtt, b, up.ref, None, pk[2]], [x.ld, cin, up.ld, cout, n, h, w, 0]),
               P.Op(P.OP_CONVT_DGRAD, dt, [dup.ref, wtt, dxt.ref, None, None, pk[3]], [dup.ld, cout, dxt.ld, cin, 0, 0, 0, n, h, w])]
        if packed:
            lst.insert(0, P.Op(P.OP_PACK_WEIGHTS, 0, [P.Ref("wtab", 0), P.Ref("params", 0), P.Ref("wpack", 0)], [len(entries), start]))
        return lst
    mem = img.mem()
    mem["wtab"] = tab.reshape(-1).view(np.uint8).copy()
    mem["wpack"] = np.zeros(dst * 2 + 256, np.uint8)
    st = E.Emulator({"act": 0, "f32": 0, "zero": 0, "params": 0, "state": 0, "step": 0}).state
    compare(ops(False), img, dt, tol=4e-3)          # per-call packing against the emulator first
    out0, _ = run_ops_gpu(ops(False), mem, st)
    out1, _ = run_ops_gpu(ops(True), mem, st)
    a0 = np.frombuffer(out0["act"], np.float16, (len(out0["act"]) - 256) // 2).astype(np.float32)
    a1 = np.frombuffer(out1["act"], np.float16, (len(out1["act"]) - 256) // 2).astype(np.float32)
    for name, v, npx in (("conv fwd", y, n * h * w), ("conv dgrad", dx, n * h * w), ("convT fwd", up, 4 * n * h * w),
                         ("convT dgrad", dxt, n * h * w)):
        lo, cnt = v.ref.off // 2, npx * v.ld
        d = np.abs(a0[lo:lo + cnt] - a1[lo:lo + cnt])
        assert d.max() == 0, "%s: packed-weight path differs, max |d| %.3e at %d (of %d), values %.4f vs %.4f" % (
            name, d.max(), int(d.argmax()), cnt, a0[lo + d.argmax()], a1[lo + d.argmax()])
        assert np.abs(a1[lo:lo + cnt]).max() > 0


@pytest.mark.parametrize("c,cout", [(64, 32), (512, 256), (96, 32)])
def test_bn_bwd_sums_from_wgrad(c, cout):
    img = Img(61)
    w = img.farr(img.par, 9 * c * cout, scale=0.1)
    dw = img.farr(img.gr, 9 * c * cout, scale=2.0)
    cs = img.farr(img.f32, c, scale=5.0)
    gamma, beta = img.farr(img.par, c, fill="pos"), img.farr(img.par, c, scale=0.3)
    sums = img.farr(img.zero, 2 * c, scale=1.0, dtype=np.float64)
    ops = [P.Op(P.OP_BN_BWD_SUMS_WGRAD, 0, [w, dw, cs, gamma, beta, sums], [c, cout, 9])]
    compare(ops, img, P.F32, tol=1e-5)


# ---- 1-bit ReLU masks (planner option relu_bits, B2U_ACT_RELU_BITS)
@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 32, 32, 32, 32), (1, 32, 40, 32, 64), (1, 16, 16, 64, 64), (2, 24, 40, 128, 128),
                                            (1, 16, 16, 256, 512), (1, 56, 56, 16, 16), (1, 16, 16, 32, 80)])
@pytest.mark.parametrize("dwmerge", [0, 1, "rowstrip"])
def test_tc_conv3x3_relu_bits_roundtrip(n, h, w, cin, cout, dwmerge):
    """forward writes the packed mask of y > 0 (epilogue variant kF_BITS_OUT), the data gradient of the next conv reads
    it (kF_BITS_IN) with and without column sums.  A mask bit flips with the sign of a near-zero pre-activation, so the
    inputs are small integers and the weights multiples of 1/8: every partial sum is exact in fp32 in any order and
    the GPU must reproduce the emulator's bits, outputs and gradients EXACTLY (tolerance 0 up to fp16 storage)."""
    lib = importlib.import_module(PKG + "._lib").lib()
    # 1: thin layers through conv_tc3w.cu, "rowstrip": the Cout = 16 / 32 ones through conv_tc3r.cu (same bits, same sums)
    old = lib.b2u_set_option(b"tc_dwmerge", 1 if dwmerge == 1 else 0)
    old_r = lib.b2u_set_option(b"tc_rowstrip", 1 if dwmerge == "rowstrip" else 0)
    try:
        _relu_bits_roundtrip(n, h, w, cin, cout)
    finally:
        lib.b2u_set_option(b"tc_dwmerge", old)
        lib.b2u_set_option(b"tc_rowstrip", old_r)


def _relu_bits_roundtrip(n, h, w, cin, cout):
    img = Img(81)
    x = img.view(n, h, w, cin, dt, fill="int")
    y = img.view(n, h, w, cout, dt, ld=cout + 16, c0=8, fill=None)
    wt = img.farr(img.par, 9 * cin * cout, fill="int", scale=0.125)
    b = img.farr(img.par, cout, fill="int", scale=0.125)
    bits = img.act.alloc(n * h * w * cout // 8)
    stats = img.zero.alloc(2 * cout * 8)
    dy = img.view(n, h, w, cin, dt, fill="int")                # gradient of a following conv's output (cin channels again)
    wt2 = img.farr(img.par, 9 * cout * cin, fill="int", scale=0.125)
    dx = img.view(n, h, w, cout, dt, ld=2 * cout, c0=cout, fill=None)
    db = img.farr(img.gr, cout, fill="int", scale=1.0)
    for st, cs in ((None, None), (stats, db)):
        ops = [P.Op(P.OP_CONV3X3_FWD, dt, [x.ref, wt, b, y.ref, st, None, bits], [x.ld, cin, 1, y.ld, cout, n, h, w]),
               P.Op(P.OP_CONV3X3_DGRAD, dt, [dy.ref, wt2, dx.ref, bits, cs],
                    [dy.ld, cin, dx.ld, cout, cout, P.ACT_RELU_BITS, 0, n, h, w])]
        compare(ops, img, dt, tol=1e-6)


# ---- streamed weights shared by CTA pairs (TMA multicast inside clusters of two, b2u_set_option("tc_mcast", 1)) -------------
@pytest.mark.parametrize("n,h,w,cin,cout", [(1, 32, 32, 256, 512), (1, 14, 14, 256, 512), (1, 16, 16, 512, 256), (2, 24, 40, 128, 128),
                                            (3, 16, 8, 256, 256), (1, 16, 8, 256, 256), (4, 128, 128, 128, 128), (5, 40, 24, 192, 128)])
@pytest.mark.parametrize("cs", [2, 8])
def test_tc_conv3x3_streamed_weights_multicast_pairs(n, h, w, cin, cout, cs):
    """wide layers stream their filter bank through shared memory; with tc_mcast the kernel runs as clusters of two CTAs that
    work on the same N tile for two pixel tiles and load every weight tile once for the pair.  Same results as the
    single-CTA schedule: forward with BN statistics, data gradient with mask + column sums; odd pixel-tile counts (a dummy
    tile in the last pair), a single pixel tile (falls back), several N tiles, ring slots of one and of three tiles"""
    lib = importlib.import_module(PKG + "._lib").lib()
    old = lib.b2u_set_option(b"tc_mcast", cs)
    try:
        img = Img(141)
        x = img.view(n, h, w, cin, dt, fill="uniform")
        y = img.view(n, h, w, cout, dt, ld=cout + 16, c0=8, fill=None)
        wt = img.farr(img.par, 9 * cin * cout, scale=(2.0 / (9 * cin)) ** 0.5)
        b = img.farr(img.par, cout, scale=0.1)
        stats = img.zero.alloc(2 * cout * 8)
        dy = img.view(n, h, w, cout, dt, scale=0.5)
        dx = img.view(n, h, w, cin, dt, fill=None)
        db = img.farr(img.gr, cin, scale=0.01)
        ops = [P.Op(P.OP_CONV3X3_FWD, dt, [x.ref, wt, b, y.ref, stats], [x.ld, cin, 1, y.ld, cout, n, h, w]),
               P.Op(P.OP_CONV3X3_DGRAD, dt, [dy.ref, wt, dx.ref, x.ref, db], [dy.ld, cout, dx.ld, cin, x.ld, 1, 0, n, h, w])]
        compare(ops, img, dt, tol=4e-3)
    finally:
        lib.b2u_set_option(b"tc_mcast", old)


# ---- dw-merged thin-layer kernel (conv_tc3w.cu, b2u_set_option("tc_dwmerge", 1))
@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 32, 32, 32, 32), (1, 24, 40, 64, 32), (1, 16, 30, 64, 64), (1, 56, 56, 16, 16),
                                            (1, 8, 14, 32, 48), (2, 64, 64, 32, 64), (4, 128, 128, 32, 32)])
def test_tc_conv3x3_dwmerge_fwd_and_dgrad(n, h, w, cin, cout):
    lib = importlib.import_module(PKG + "._lib").lib()
    old = lib.b2u_set_option(b"tc_dwmerge", 1)
    try:
        img = Img(91)
        x = img.view(n, h, w, cin, dt, ld=2 * cin, c0=cin, fill="uniform")
        y = img.view(n, h, w, cout, dt, ld=cout + 16, c0=8, fill=None)
        wt = img.farr(img.par, 9 * cin * cout, scale=(2.0 / (9 * cin)) ** 0.5)
        b = img.farr(img.par, cout, scale=0.1)
        stats = img.zero.alloc(2 * cout * 8)
        for act in (1, 2):
            ops = [P.Op(P.OP_CONV3X3_FWD, dt, [x.ref, wt, b, y.ref, stats], [x.ld, cin, act, y.ld, cout, n, h, w])]
            compare(ops, img, dt, tol=3e-3)
        img = Img(92)
        dy = img.view(n, h, w, cout, dt, ld=cout + 8, scale=0.5)
        dx = img.view(n, h, w, cin, dt, ld=2 * cin, c0=0, scale=0.3)
        mask = img.view(n, h, w, cin, dt)
        wt = img.farr(img.par, 9 * cin * cout, scale=(2.0 / (9 * cin)) ** 0.5)
        db = img.farr(img.gr, cin, scale=0.01)
        for acc, mact, cs in ((0, 1, db), (1, 2, None), (0, 0, db)):
            if cin > 64:
                break
            ops = [P.Op(P.OP_CONV3X3_DGRAD, dt, [dy.ref, wt, dx.ref, mask.ref if mact else None, cs],
                        [dy.ld, cout, dx.ld, cin, mask.ld, mact, acc, n, h, w])]
            compare(ops, img, dt, tol=4e-3)
    finally:
        lib.b2u_set_option(b"tc_dwmerge", old)


# ---- row-strip kernel for Cout = 16 / 32 (conv_tc3r.cu, b2u_set_option("tc_rowstrip", 1)) ----------------------------
@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 20, 300, 32, 32), (1, 64, 128, 64, 32), (1, 9, 140, 16, 16), (2, 40, 96, 32, 16),
                                            (1, 33, 257, 128, 32), (3, 5, 32, 48, 32), (1, 1, 130, 32, 32), (8, 64, 64, 32, 32)])
def test_tc_conv3x3_rowstrip_fwd_and_dgrad(n, h, w, cin, cout):
    """row strips of 128 pixels (ragged last strip, images narrower than a strip, one-row images, CTA ranges that cross
    strip and image boundaries), every epilogue variant: statistics, 1-bit / fp16 masks, accumulate, column sums"""
    lib = importlib.import_module(PKG + "._lib").lib()
    old = lib.b2u_set_option(b"tc_rowstrip", 1)
    try:
        img = Img(93)
        x = img.view(n, h, w, cin, dt, ld=2 * cin, c0=cin, fill="uniform")
        y = img.view(n, h, w, cout, dt, ld=cout + 16, c0=8, fill=None)
        wt = img.farr(img.par, 9 * cin * cout, scale=(2.0 / (9 * cin)) ** 0.5)
        b = img.farr(img.par, cout, scale=0.1)
        stats = img.zero.alloc(2 * cout * 8)
        for act, st in ((1, stats), (2, None)):
            ops = [P.Op(P.OP_CONV3X3_FWD, dt, [x.ref, wt, b, y.ref, st], [x.ld, cin, act, y.ld, cout, n, h, w])]
            compare(ops, img, dt, tol=3e-3)
        # data gradient INTO a thin tensor: dy has `cin` channels here, dx `cout` (= 16 / 32)
        img = Img(94)
        dy = img.view(n, h, w, cin, dt, ld=cin + 8, scale=0.5)
        dx = img.view(n, h, w, cout, dt, ld=2 * cout, c0=0, scale=0.3)
        mask = img.view(n, h, w, cout, dt)
        wt = img.farr(img.par, 9 * cout * cin, scale=(2.0 / (9 * cin)) ** 0.5)
        db = img.farr(img.gr, cout, scale=0.01)
        for acc, mact, cs in ((0, 1, db), (1, 2, None), (0, 0, db), (0, 0, None)):
            ops = [P.Op(P.OP_CONV3X3_DGRAD, dt, [dy.ref, wt, dx.ref, mask.ref if mact else None, cs],
                        [dy.ld, cin, dx.ld, cout, mask.ld, mact, acc, n, h, w])]
            compare(ops, img, dt, tol=4e-3)
    finally:
        lib.b2u_set_option(b"tc_rowstrip", old)
