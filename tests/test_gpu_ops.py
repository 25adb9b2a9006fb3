"""Per-kernel parity: every op kind, run on the GPU through the C-ABI, against tests/emulator.py on the
same memory image -- both storage types, strided (concat-slice) views, accumulate / mask variants,
ragged sizes (H, W not multiples of the tile; channel counts that are not powers of two)."""
import numpy as np
import pytest

import emulator as E
from gpu_harness import run_ops_gpu
from helpers import P

pytestmark = pytest.mark.gpu

NPDT = E.NPDT
TOL = {P.F32: 2e-5, P.F16: 2e-3}


class Img:
    """builds a memory image + views for one test"""

    def __init__(self, seed=0):
        self.act, self.f32, self.zero, self.par, self.gr = (P.Arena(n) for n in ("act", "f32", "zero", "params", "grads"))
        self.init = []          # (ref, ndarray)
        self.rng = np.random.default_rng(seed)

    def view(self, n, h, w, c, dt, ld=None, c0=0, fill="normal", scale=1.0):
        ld = ld or c
        ref = self.act.alloc(n * h * w * ld * P.ELEM[dt])
        v = P.View(ref + c0 * P.ELEM[dt], ld, c, h, w, dt)
        if fill is not None:
            if fill == "int":            # small integers: sums of products stay exact in fp32 whatever the order
                full = self.rng.integers(-3, 4, (n * h * w, ld)).astype(np.float64) * scale
            else:
                full = self.rng.standard_normal((n * h * w, ld)) * scale if fill == "normal" else self.rng.random((n * h * w, ld))
            self.init.append((ref, full.astype(NPDT[dt])))
        return v

    def farr(self, arena, count, fill="normal", scale=1.0, dtype=np.float32):
        ref = arena.alloc(count * np.dtype(dtype).itemsize)
        if fill == "normal":
            a = (self.rng.standard_normal(count) * scale).astype(dtype)
        elif fill == "uniform":
            a = self.rng.random(count).astype(dtype)
        elif fill == "int":              # multiples of `scale` (a power of two): exact in fp16 and in fp32 sums
            a = (self.rng.integers(-4, 5, count) * scale).astype(dtype)
        elif fill == "pos":
            a = (0.5 + self.rng.random(count)).astype(dtype)
        else:
            a = np.zeros(count, dtype)
        self.init.append((ref, a))
        return ref

    def mem(self):
        sizes = {"act": self.act.size, "f32": self.f32.size, "zero": self.zero.size, "params": self.par.size,
                 "grads": self.gr.size}
        m = {k: np.zeros(int(v) + 256, np.uint8) for k, v in sizes.items()}
        for ref, a in self.init:
            raw = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
            m[ref.arena][ref.off:ref.off + raw.size] = raw
        return m


def compare(ops, img, dt, state=None, tol=None, graph=False, arenas=("act", "f32", "zero", "params", "grads")):
    mem = img.mem()
    em = E.Emulator({"act": 0, "f32": 0, "zero": 0, "params": 0, "state": 0, "step": 0})
    em.mem = {k: v.copy() for k, v in mem.items()}
    if state:
        em.state.update(state)
    st0 = dict(em.state)
    em.run(ops)
    out, st = run_ops_gpu(ops, mem, st0, graph=graph)
    tol = tol or TOL[dt]
    for a in arenas:
        for kind, nb in ((NPDT[dt], P.ELEM[dt]), (np.float32, 4)) if a == "act" else ((np.float32, 4),) if a != "zero" else ((np.float64, 8),):
            n = (len(mem[a]) - 256) // nb
            if n == 0:
                continue
            g = np.frombuffer(out[a], dtype=kind, count=n).astype(np.float64)
            w = np.frombuffer(em.mem[a], dtype=kind, count=n).astype(np.float64)
            ok = np.isfinite(w)
            if a == "act" and kind != NPDT[dt]:
                continue
            scale = max(np.abs(w[ok]).max() if ok.any() else 0, 1e-6)
            err = np.abs(g[ok] - w[ok]).max() / scale
            assert err < tol, "arena %s: rel err %.3e (scale %.3e)" % (a, err, scale)
    return out, st


CONV_SHAPES = [  # n, h, w, cin, cout
    (2, 16, 16, 8, 32), (1, 20, 12, 32, 64), (2, 8, 8, 64, 16), (1, 33, 17, 16, 96), (2, 16, 16, 1, 32), (1, 16, 16, 3, 16),
    (2, 21, 37, 3, 32), (3, 40, 70, 3, 16),          # Cin = 3 (Task-2 slices replicated to RGB): ragged tiles, several tiles
]


@pytest.mark.parametrize("dt", [P.F32, P.F16])
@pytest.mark.parametrize("n,h,w,cin,cout", CONV_SHAPES)
def test_conv3x3_fwd(dt, n, h, w, cin, cout):
    img = Img(1)
    x = img.view(n, h, w, cin, dt, ld=cin + (8 if cin % 8 == 0 else 0), c0=0)
    y = img.view(n, h, w, cout, dt, ld=cout + 16, c0=8, fill=None)
    wt = img.farr(img.par, 9 * cin * cout, scale=(2.0 / (9 * cin)) ** 0.5)
    b = img.farr(img.par, cout, scale=0.1)
    stats = img.zero.alloc(2 * cout * 8)
    for act in (1, 2):
        ops = [P.Op(P.OP_CONV3X3_FWD, dt, [x.ref, wt, b, y.ref, stats], [x.ld, cin, act, y.ld, cout, n, h, w])]
        compare(ops, img, dt, tol=3e-3 if dt == P.F16 else 2e-5)


@pytest.mark.parametrize("dt", [P.F32, P.F16])
@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 16, 16, 1, 32), (1, 20, 12, 1, 16), (1, 16, 16, 8, 32), (1, 12, 20, 32, 48)])
def test_conv3x3_relu_bits_generic_paths(dt, n, h, w, cin, cout):
    """packed 1-bit ReLU masks outside the tensor-core epilogues: written by the Cin = 1 kernel in the same pass
    (conv2d_1 of every network) or by one extra pass (exact path), applied to a data gradient by one extra pass.
    Integer data: exact sums, so the bits must equal the emulator's bit for bit."""
    img = Img(83)
    x = img.view(n, h, w, cin, dt, ld=cin + (8 if cin % 8 == 0 else 0), fill="int")
    y = img.view(n, h, w, cout, dt, ld=cout + 16, c0=8, fill=None)
    wt = img.farr(img.par, 9 * cin * cout, fill="int", scale=0.125)
    b = img.farr(img.par, cout, fill="int", scale=0.125)
    bits = img.act.alloc(n * h * w * cout // 8)
    dy = img.view(n, h, w, 16, dt, fill="int")
    wt2 = img.farr(img.par, 9 * cout * 16, fill="int", scale=0.125)
    dx = img.view(n, h, w, cout, dt, fill=None)
    ops = [P.Op(P.OP_CONV3X3_FWD, dt, [x.ref, wt, b, y.ref, None, None, bits], [x.ld, cin, 1, y.ld, cout, n, h, w]),
           P.Op(P.OP_CONV3X3_DGRAD, dt, [dy.ref, wt2, dx.ref, bits, None],
                [dy.ld, 16, dx.ld, cout, cout, P.ACT_RELU_BITS, 0, n, h, w])]
    compare(ops, img, dt, tol=1e-6)


@pytest.mark.parametrize("dt", [P.F32, P.F16])
@pytest.mark.parametrize("n,h,w,cin,cout", CONV_SHAPES[:4])
def test_conv3x3_dgrad_and_wgrad(dt, n, h, w, cin, cout):
    img = Img(2)
    x = img.view(n, h, w, cin, dt, fill="uniform")
    dy = img.view(n, h, w, cout, dt, ld=cout + 8, scale=0.5)
    dx = img.view(n, h, w, cin, dt, ld=cin + 8, c0=8, scale=0.3)
    mask = img.view(n, h, w, cin, dt)
    wt = img.farr(img.par, 9 * cin * cout, scale=(2.0 / (9 * cin)) ** 0.5)
    dw = img.farr(img.gr, 9 * cin * cout, scale=0.01)
    db = img.farr(img.gr, cout, scale=0.01)
    for acc, mact in ((0, 1), (1, 2), (0, 0)):
        ops = [P.Op(P.OP_CONV3X3_DGRAD, dt, [dy.ref, wt, dx.ref, mask.ref if mact else None],
                    [dy.ld, cout, dx.ld, cin, mask.ld, mact, acc, n, h, w]),
               P.Op(P.OP_CONV3X3_WGRAD, dt, [x.ref, dy.ref, dw, db], [x.ld, cin, dy.ld, cout, n, h, w])]
        compare(ops, img, dt, tol=4e-3 if dt == P.F16 else 5e-5)


@pytest.mark.parametrize("dt", [P.F32, P.F16])
@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 8, 8, 64, 32), (1, 5, 7, 32, 16), (1, 4, 4, 512, 256)])
def test_convt2x2(dt, n, h, w, cin, cout):
    img = Img(3)
    x = img.view(n, h, w, cin, dt, fill="uniform")
    y = img.view(n, 2 * h, 2 * w, cout, dt, ld=2 * cout, c0=0, fill=None)
    dy = img.view(n, 2 * h, 2 * w, cout, dt, ld=2 * cout, c0=cout, scale=0.5)
    dx = img.view(n, h, w, cin, dt, scale=0.2)
    wt = img.farr(img.par, 4 * cout * cin, scale=(1.0 / cin) ** 0.5)
    b = img.farr(img.par, cout, scale=0.1)
    dw = img.farr(img.gr, 4 * cout * cin, fill="zero")
    db = img.farr(img.gr, cout, fill="zero")
    ops = [P.Op(P.OP_CONVT_FWD, dt, [x.ref, wt, b, y.ref, None], [x.ld, cin, y.ld, cout, n, h, w, 0]),
           P.Op(P.OP_CONVT_DGRAD, dt, [dy.ref, wt, dx.ref, x.ref], [dy.ld, cout, dx.ld, cin, x.ld, 1, 1, n, h, w]),
           P.Op(P.OP_CONVT_WGRAD, dt, [x.ref, dy.ref, dw, db], [x.ld, cin, dy.ld, cout, n, h, w])]
    compare(ops, img, dt, tol=4e-3 if dt == P.F16 else 5e-5)


@pytest.mark.parametrize("dt", [P.F32, P.F16])
@pytest.mark.parametrize("c,npix_shape", [(32, (2, 16, 16)), (96, (1, 10, 6)), (512, (1, 4, 4)), (16, (3, 8, 8))])
def test_batchnorm_all_passes(dt, c, npix_shape):
    n, h, w = npix_shape
    npix = n * h * w
    img = Img(4)
    x = img.view(n, h, w, c, dt, ld=c + 8, fill="uniform")
    y = img.view(n, h, w, c, dt, ld=2 * c, c0=c, fill=None)
    dy = img.view(n, h, w, c, dt, scale=0.5)
    dx = img.view(n, h, w, c, dt, fill=None)
    gamma, beta = img.farr(img.par, c, fill="pos"), img.farr(img.par, c, scale=0.1)
    mm, mv = img.farr(img.par, c, scale=0.1), img.farr(img.par, c, fill="pos")
    scale, shift, mean, inv = (img.f32.alloc(c * 4) for _ in range(4))
    dg, dbt = img.farr(img.gr, c, scale=0.01), img.farr(img.gr, c, scale=0.01)
    s1, s2 = img.zero.alloc(2 * c * 8), img.zero.alloc(2 * c * 8)
    for training in (1, 0):
        ops = [P.Op(P.OP_BN_STATS, dt, [x.ref, s1], [x.ld, c, npix]),
               P.Op(P.OP_BN_FINALIZE, 0, [s1, gamma, beta, mm, mv, scale, shift, mean, inv], [npix, training, c], [0.99, 1e-3]),
               P.Op(P.OP_BN_APPLY, dt, [x.ref, y.ref, scale, shift, None], [x.ld, y.ld, c, npix, 0])]
        if training:
            ops += [P.Op(P.OP_BN_BWD_REDUCE, dt, [dy.ref, x.ref, mean, inv, s2], [dy.ld, x.ld, c, npix]),
                    P.Op(P.OP_BN_BWD_APPLY, dt, [dy.ref, x.ref, dx.ref, gamma, mean, inv, s2, dg, dbt, x.ref],
                         [dy.ld, x.ld, dx.ld, c, npix, x.ld, 1, npix])]
        compare(ops, img, dt, tol=3e-3 if dt == P.F16 else 3e-5)


@pytest.mark.parametrize("dt", [P.F32, P.F16])
@pytest.mark.parametrize("ca,cb,npix_shape", [(32, 32, (2, 16, 16)), (64, 32, (1, 10, 6)), (8, 24, (3, 8, 8)), (256, 256, (1, 4, 4))])
def test_batchnorm_over_a_split_concatenate(dt, ca, cb, npix_shape):
    """BN apply / backward apply whose input (and input gradient) is a two-input concatenate kept as two dense tensors
    (plan.py split_concat): channels [0, ca) from / to the first tensor, the rest from / to the second"""
    n, h, w = npix_shape
    npix, c = n * h * w, ca + cb
    img = Img(41)
    xa = img.view(n, h, w, ca, dt, fill="uniform")
    xb = img.view(n, h, w, cb, dt, ld=cb + 8, fill="uniform")
    y = img.view(n, h, w, c, dt, fill=None)
    dy = img.view(n, h, w, c, dt, scale=0.5)
    da, db_ = img.view(n, h, w, ca, dt, ld=ca + 16, c0=8, fill=None), img.view(n, h, w, cb, dt, fill=None)
    gamma, beta = img.farr(img.par, c, fill="pos"), img.farr(img.par, c, scale=0.1)
    mm, mv = img.farr(img.par, c, scale=0.1), img.farr(img.par, c, fill="pos")
    scale, shift, mean, inv = (img.f32.alloc(c * 4) for _ in range(4))
    dg, dbt = img.farr(img.gr, c, scale=0.01), img.farr(img.gr, c, scale=0.01)
    s1, s2 = img.zero.alloc(2 * c * 8), img.zero.alloc(2 * c * 8)
    ostats = img.zero.alloc(2 * c * 8)
    ops = [P.Op(P.OP_BN_STATS, dt, [xa.ref, s1], [xa.ld, ca, npix, c]),            # sums of the two halves, squares at offset c
           P.Op(P.OP_BN_STATS, dt, [xb.ref, s1 + ca * 8], [xb.ld, cb, npix, c]),
           P.Op(P.OP_BN_FINALIZE, 0, [s1, gamma, beta, mm, mv, scale, shift, mean, inv], [npix, 1, c], [0.99, 1e-3]),
           P.Op(P.OP_BN_APPLY, dt, [xa.ref, y.ref, scale, shift, ostats, xb.ref], [xa.ld, y.ld, c, npix, c, ca, xb.ld]),
           P.Op(P.OP_BN_BWD_APPLY, dt, [dy.ref, xa.ref, da.ref, gamma, mean, inv, s2, dg, dbt, None, None, xb.ref, db_.ref],
                [dy.ld, xa.ld, da.ld, c, npix, 0, 0, npix, ca, xb.ld, db_.ld])]
    # backward sums: any fixed numbers do (the kernels only consume them)
    img.init.append((s2, np.linspace(-3.0, 3.0, 2 * c).astype(np.float64)))
    compare(ops, img, dt, tol=3e-3 if dt == P.F16 else 3e-5)


@pytest.mark.parametrize("dt", [P.F32, P.F16])
@pytest.mark.parametrize("c,npix_shape,rate", [(32, (2, 16, 16), 0.5), (64, (1, 10, 6), 0.2), (128, (3, 8, 8), 0.5), (8, (1, 7, 5), 0.1)])
def test_batchnorm_over_an_unmaterialised_dropout(dt, c, npix_shape, rate):
    """Conv2D -> Dropout -> BatchNormalization without the Dropout pass (plan.py fuse_dropout_bn): the statistics op draws the
    keep mask (same Philox stream / element index as dropout_kernel) and stores it as packed bits, apply / backward reduction /
    backward apply read the bits -- and the explicit Dropout op on the same stream gives the same masked values"""
    n, h, w = npix_shape
    npix = n * h * w
    img = Img(51)
    x = img.view(n, h, w, c, dt, ld=c + 8, fill="uniform")
    xd = img.view(n, h, w, c, dt, fill=None)                  # explicit dropout output (reference path inside this test)
    y, y2 = img.view(n, h, w, c, dt, fill=None), img.view(n, h, w, c, dt, fill=None)
    dy = img.view(n, h, w, c, dt, scale=0.5)
    dx = img.view(n, h, w, c, dt, ld=c + 16, c0=8, fill=None)
    gamma, beta = img.farr(img.par, c, fill="pos"), img.farr(img.par, c, scale=0.1)
    mm, mv = img.farr(img.par, c, scale=0.1), img.farr(img.par, c, fill="pos")
    scale, shift, mean, inv = (img.f32.alloc(c * 4) for _ in range(4))
    dg, dbt = img.farr(img.gr, c, scale=0.01), img.farr(img.gr, c, scale=0.01)
    bias_g = img.farr(img.gr, c, scale=0.01)
    s1, s2 = img.zero.alloc(2 * c * 8), img.zero.alloc(2 * c * 8)
    step = P.Ref("step", 0)
    op_id = 5
    bits = img.act.alloc(npix * c // 8)                       # packed keep mask: written by the statistics op, read by the others
    ops = [P.Op(P.OP_BN_STATS, dt, [x.ref, s1, step, bits], [x.ld, c, npix, 0, op_id], [rate]),
           P.Op(P.OP_BN_FINALIZE, 0, [s1, gamma, beta, mm, mv, scale, shift, mean, inv], [npix, 1, c], [0.99, 1e-3]),
           P.Op(P.OP_BN_APPLY, dt, [x.ref, y.ref, scale, shift, None, None, bits], [x.ld, y.ld, c, npix, 0], [rate]),
           P.Op(P.OP_BN_BWD_REDUCE, dt, [dy.ref, x.ref, mean, inv, s2, bits], [dy.ld, x.ld, c, npix], [rate]),
           P.Op(P.OP_BN_BWD_APPLY, dt, [dy.ref, x.ref, dx.ref, gamma, mean, inv, s2, dg, dbt, x.ref, bias_g, None, None, bits],
                [dy.ld, x.ld, dx.ld, c, npix, x.ld, 2, npix], [rate]),
           # the explicit pair on the same stream: dropout, then BN apply of its output
           P.Op(P.OP_DROPOUT_FWD, dt, [x.ref, xd.ref, step], [x.ld, xd.ld, c, npix, op_id], [rate]),
           P.Op(P.OP_BN_APPLY, dt, [xd.ref, y2.ref, scale, shift, None], [xd.ld, y2.ld, c, npix, 0])]
    out, _ = compare(ops, img, dt, state=dict(seed=987654321, step=11), tol=3e-3 if dt == P.F16 else 3e-5)
    # GPU: fused and explicit outputs agree (the explicit path rounds the dropout output to the storage type first)
    es = P.ELEM[dt]
    a = np.frombuffer(out["act"], dtype=NPDT[dt], count=npix * c, offset=y.ref.off).astype(np.float32)
    b = np.frombuffer(out["act"], dtype=NPDT[dt], count=npix * c, offset=y2.ref.off).astype(np.float32)
    assert np.abs(a - b).max() <= (4e-3 if dt == P.F16 else 1e-6) * max(1.0, np.abs(b).max())


@pytest.mark.parametrize("dt", [P.F32, P.F16])
@pytest.mark.parametrize("p", [0.0, 0.25])
def test_maxpool_dropout_fwd_bwd(dt, p):
    n, h, w, c = 2, 12, 20, 32
    img = Img(5)
    x = img.view(n, h, w, c, dt, ld=2 * c, c0=c)
    # force ties (post-ReLU zeros are the common case): quantise the input
    ref0, full = img.init[-1]
    img.init[-1] = (ref0, np.round(full.astype(np.float32) * 2).astype(NPDT[dt]) / 2)
    y = img.view(n, h // 2, w // 2, c, dt, fill=None)
    dy = img.view(n, h // 2, w // 2, c, dt)
    dx = img.view(n, h, w, c, dt, ld=2 * c, c0=c, scale=0.1)
    step = P.Ref("step", 0)
    for acc in (0, 1):
        ops = [P.Op(P.OP_MAXPOOL_FWD, dt, [x.ref, y.ref, step if p else None], [x.ld, y.ld, c, n, h, w, 3], [p]),
               P.Op(P.OP_MAXPOOL_BWD, dt, [x.ref, dy.ref, dx.ref, step if p else None],
                    [x.ld, dy.ld, dx.ld, c, n, h, w, 3, acc], [p])]
        compare(ops, img, dt, state=dict(seed=1234567890123, step=17))


@pytest.mark.parametrize("dt", [P.F32, P.F16])
def test_dropout_copy_slice(dt):
    n, h, w, c = 2, 6, 10, 96
    img = Img(6)
    x = img.view(n, h, w, c, dt)
    y = img.view(n, h, w, c, dt, ld=c + 32, c0=16, fill=None)
    dy = img.view(n, h, w, c, dt)
    dx = img.view(n, h, w, c, dt, fill=None)
    d2 = img.view(n, h, w, c, dt, ld=2 * c, c0=c)
    step = P.Ref("step", 0)
    ops = [P.Op(P.OP_DROPOUT_FWD, dt, [x.ref, y.ref, step], [x.ld, y.ld, c, n * h * w, 2], [0.4]),
           P.Op(P.OP_DROPOUT_BWD, dt, [dy.ref, dx.ref, step, x.ref], [dy.ld, dx.ld, c, n * h * w, 2, x.ld, 2], [0.4]),
           P.Op(P.OP_COPY_SLICE, dt, [x.ref, d2.ref], [x.ld, d2.ld, c, n * h * w, 1]),
           P.Op(P.OP_COPY_SLICE, dt, [dy.ref, y.ref], [dy.ld, y.ld, c, n * h * w, 0])]
    compare(ops, img, dt, state=dict(seed=99, step=2 ** 33 + 5))


@pytest.mark.parametrize("dt", [P.F32, P.F16])
def test_head_loss_and_backward(dt):
    n, h, w, c = 2, 16, 24, 32
    npix = n * h * w
    img = Img(7)
    x = img.view(n, h, w, c, dt, fill="uniform")
    dx = img.view(n, h, w, c, dt, fill=None)
    wt, b = img.farr(img.par, c, scale=0.5), img.farr(img.par, 1, scale=0.1)
    prob = img.f32.alloc(npix * 4)
    tgt = img.farr(img.f32, npix, fill="uniform")
    out = img.f32.alloc(16)
    sums = img.zero.alloc(32)
    dw, db = img.farr(img.gr, c, fill="zero"), img.farr(img.gr, 1, fill="zero")
    step = P.Ref("step", 0)
    ops = [P.Op(P.OP_HEAD_FWD, dt, [x.ref, wt, b, prob], [x.ld, c, npix]),
           P.Op(P.OP_BCE_DICE_SUMS, 0, [prob, tgt, sums], [npix]),
           P.Op(P.OP_BCE_DICE_FINALIZE, 0, [sums, out], [npix]),
           P.Op(P.OP_HEAD_BWD, dt, [prob, tgt, sums, step, x.ref, wt, dx.ref, dw, db], [npix, x.ld, c, dx.ld, 1, npix])]
    compare(ops, img, dt, state=dict(loss_scale=1024.0), tol=3e-3 if dt == P.F16 else 5e-5)


def test_adam_and_state_advance():
    n = 100003
    img = Img(8)
    p = img.farr(img.par, n)
    g = img.farr(img.gr, n, scale=3.0)
    m = img.farr(img.f32, n, scale=0.1)
    v = img.farr(img.f32, n, fill="uniform")
    step = P.Ref("step", 0)
    ops = [P.Op(P.OP_ADAM, 0, [p, g, m, v, step], [n]), P.Op(P.OP_STATE_ADVANCE, 0, [step])]
    st = dict(lr=5e-4, beta1_pow=0.9 ** 3, beta2_pow=0.999 ** 3, loss_scale=8.0, grad_div=2.0, step=2)
    out, st2 = compare(ops, img, P.F32, state=st, tol=1e-5)
    assert st2.step == 3 and abs(st2.beta1_pow - 0.9 ** 4) < 1e-6 and st2.overflow == 0


def test_adam_skips_the_whole_step_on_a_non_finite_gradient():
    """one non-finite gradient value (fp16 overflow under the static loss scale): NOTHING is updated -- parameters and
    both moments keep their values, the step counter and the beta powers do not advance -- the loss scale is halved and
    the sticky overflow flag tells the host; the next (finite) step then runs normally with the smaller scale"""
    n = 4099
    img = Img(9)
    p = img.farr(img.par, n)
    g = img.farr(img.gr, n)
    m, v = img.farr(img.f32, n, scale=0.1), img.farr(img.f32, n, fill="uniform")
    step = P.Ref("step", 0)
    ops = [P.Op(P.OP_ADAM, 0, [p, g, m, v, step], [n]), P.Op(P.OP_STATE_ADVANCE, 0, [step])]
    st = dict(lr=5e-4, beta1_pow=0.9 ** 3, beta2_pow=0.999 ** 3, loss_scale=1024.0, grad_div=1.0, step=2)
    for bad in (np.inf, np.nan):
        img.init[1][1][n - 2] = bad                                   # (init[1] is the gradient buffer `g`)
        mem = img.mem()
        out, st2 = compare(ops, img, P.F32, state=st, tol=1e-6, arenas=("params", "f32"))
        assert st2.overflow == 1 and st2.skip_step == 0 and st2.step == 2 and st2.loss_scale == 512.0
        assert abs(st2.beta1_pow - 0.9 ** 3) < 1e-7
        for arena in ("params", "f32"):
            assert np.array_equal(out[arena], mem[arena])            # bit-identical: no partial update
    img.init[1][1][n - 2] = 0.25
    out, st3 = compare(ops, img, P.F32, state=dict(st, loss_scale=512.0), tol=1e-5, arenas=("params", "f32"))
    assert st3.overflow == 0 and st3.step == 3 and st3.loss_scale == 512.0


@pytest.mark.parametrize("n", [6, 70])                      # split-K Dense(32): a partial and several ragged sample groups
@pytest.mark.parametrize("dt", [P.F32, P.F16])
def test_dense_and_bce(dt, n):
    k = 28 * 28 * 64
    img = Img(10)
    x = img.view(n, 1, 1, k, dt, fill="uniform")
    dxv = img.view(n, 1, 1, k, dt, fill=None)
    w1, b1 = img.farr(img.par, k * 32, scale=0.01), img.farr(img.par, 32, scale=0.1)
    w2, b2 = img.farr(img.par, 32, scale=0.3), img.farr(img.par, 1, scale=0.1)
    y1, y2 = img.f32.alloc(n * 32 * 4), img.f32.alloc(n * 4)
    d1, d2 = img.f32.alloc(n * 32 * 4), img.f32.alloc(n * 4)
    tgt, sw, out = img.farr(img.f32, n, fill="uniform"), img.farr(img.f32, n, fill="pos"), img.f32.alloc(16)
    gw1, gb1 = img.farr(img.gr, k * 32, fill="zero"), img.farr(img.gr, 32, fill="zero")
    gw2, gb2 = img.farr(img.gr, 32, fill="zero"), img.farr(img.gr, 1, fill="zero")
    step = P.Ref("step", 0)
    ops = [P.Op(P.OP_DENSE_FWD, dt, [x.ref, w1, b1, y1], [k, 1, 32, n]),
           P.Op(P.OP_DENSE_FWD, P.F32, [y1, w2, b2, y2], [32, 3, 1, n]),
           P.Op(P.OP_BCE_FWD, 0, [y2, tgt, sw, out], [n]),
           P.Op(P.OP_BCE_SIGMOID_BWD, P.F32, [y2, tgt, sw, step, d2], [n]),
           P.Op(P.OP_DENSE_BWD, P.F32, [y1, w2, y2, d2, d1, y1, gw2, gb2], [32, 0, 1, 1, n]),
           P.Op(P.OP_DENSE_BWD, dt, [x.ref, w1, y1, d1, dxv.ref, None, gw1, gb1], [k, 0, 0, 32, n])]
    compare(ops, img, dt, tol=3e-3 if dt == P.F16 else 5e-5)


def test_threshold_counts_against_oracle():
    import ctypes as C
    import torch
    from gpu_harness import LIB
    from oracle import keras_ref as K
    rng = np.random.default_rng(3)
    n = 3 * 224 * 224 + 17                       # ragged length
    p = rng.random(n).astype(np.float32)
    t = np.clip(rng.random(n) * 1.5 - 0.25, 0, 1).astype(np.float32)
    thr = np.array([0.05, 0.3, 0.5, 0.547, 0.9, 0.999], np.float32)
    pd, td, thd = (torch.from_numpy(a).cuda() for a in (p, t, thr))
    tp = torch.zeros(len(thr), dtype=torch.float64, device="cuda")
    spr = torch.zeros_like(tp)
    sgt = torch.zeros(1, dtype=torch.float64, device="cuda")
    l = LIB.lib()
    LIB.check(l.b2u_threshold_counts(pd.data_ptr(), td.data_ptr(), n, thd.data_ptr(), len(thr), tp.data_ptr(),
                                     spr.data_ptr(), sgt.data_ptr(), None))
    torch.cuda.synchronize()
    for k, th in enumerate(thr):
        m = K.sm_threshold_metrics(t, p, float(th))
        assert spr[k].item() == m["sum_pr"]                         # integer counts: exact
        assert tp[k].item() == pytest.approx(m["tp"], rel=1e-6)
    assert sgt.item() == pytest.approx(float(t.astype(np.float64).sum()), rel=1e-6)


def test_gather_batch():
    import torch
    from gpu_harness import LIB
    src = torch.rand(10, 7 * 9, device="cuda")
    idx = torch.tensor([3, 3, 9, 0], dtype=torch.int32, device="cuda")
    for dt, tdt in ((P.F32, torch.float32), (P.F16, torch.float16)):
        dst = torch.zeros(4, 63, dtype=tdt, device="cuda")
        LIB.check(LIB.lib().b2u_gather_batch(dt, src.data_ptr(), idx.data_ptr(), dst.data_ptr(), 63, 4, None))
        torch.cuda.synchronize()
        assert torch.equal(dst, src[idx.long()].to(tdt))
    with pytest.raises(LIB.B2UError):
        LIB.check(LIB.lib().b2u_gather_batch(0, src.data_ptr(), None, src.data_ptr(), 63, 0, None))


def test_graph_capture_equals_eager():
    dt = P.F32
    img = Img(11)
    n, h, w, cin, cout = 1, 16, 16, 16, 32
    x = img.view(n, h, w, cin, dt)
    y = img.view(n, h, w, cout, dt, fill=None)
    wt, b = img.farr(img.par, 9 * cin * cout, scale=0.1), img.farr(img.par, cout, scale=0.1)
    ops = [P.Op(P.OP_MEMSET, 0, [P.Ref("zero", 0)], [8]),
           P.Op(P.OP_CONV3X3_FWD, dt, [x.ref, wt, b, y.ref, None], [x.ld, cin, 1, y.ld, cout, n, h, w])]
    img.zero.alloc(8)
    compare(ops, img, dt, graph=True)


@pytest.mark.parametrize("dt", [P.F32, P.F16])
@pytest.mark.parametrize("acc", [0, 1])
def test_maxpool_bwd_with_fused_bn_backward_statistics(dt, acc):
    """max-pool backward that completes a BatchNorm output's gradient also emits sum(dx), sum(dx*xhat)"""
    n, h, w, c = 2, 16, 24, 64
    img = Img(12)
    x = img.view(n, h, w, c, dt, ld=2 * c, c0=c)
    dy = img.view(n, h // 2, w // 2, c, dt)
    dx = img.view(n, h, w, c, dt, ld=2 * c, c0=c, scale=0.1)
    gamma, beta = img.farr(img.par, c, fill="pos"), img.farr(img.par, c, scale=0.2)
    sums = img.zero.alloc(2 * c * 8)
    step = P.Ref("step", 0)
    ops = [P.Op(P.OP_MAXPOOL_BWD, dt, [x.ref, dy.ref, dx.ref, step, sums, gamma, beta],
                [x.ld, dy.ld, dx.ld, c, n, h, w, 1, acc], [0.25])]
    compare(ops, img, dt, state=dict(seed=5, step=3), tol=3e-3 if dt == P.F16 else 3e-5)


@pytest.mark.parametrize("dt", [P.F32, P.F16])
@pytest.mark.parametrize("n,h,w,c,p_drop", [(2, 16, 24, 32, 0.25), (1, 10, 6, 64, 0.0), (3, 8, 8, 256, 0.25), (1, 12, 20, 96, 0.2)])
def test_bn_apply_pool_fused(dt, n, h, w, c, p_drop):
    """BN apply + 2x2 max-pool (+ dropout) in one pass: skip tensor into a concat slice, pooled tensor, statistics"""
    img = Img(71)
    x = img.view(n, h, w, c, dt, scale=2.0)
    y = img.view(n, h, w, c, dt, ld=2 * c, c0=c, fill=None)
    yp = img.view(n, h // 2, w // 2, c, dt, ld=c + 8, fill=None)
    scale, shift = img.farr(img.f32, c, fill="pos"), img.farr(img.f32, c, scale=0.5)
    stats = img.zero.alloc((2 * c + c) * 8)
    step = P.Ref("step", 0)
    ops = [P.Op(P.OP_BN_APPLY_POOL, dt, [x.ref, y.ref, scale, shift, stats, yp.ref, step if p_drop > 0 else None],
                [x.ld, y.ld, c, n * h * w, 2 * c, n, h, w, yp.ld, 3], [p_drop])]
    compare(ops, img, dt, state=dict(seed=11, step=4), tol=2e-3 if dt == P.F16 else 2e-5)
