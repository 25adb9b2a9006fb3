"""TEST INFRASTRUCTURE -- a numpy/torch interpreter for plan.py op lists.

Each op kind is emulated with exactly the semantics documented for the CUDA kernel of the same name in
include/b200unet.h (views with ld strides, accumulate flags, activation masks, in-place concat slices).
It lets the CPU suite (`-m "not gpu"`) validate the *schedule* the planner emits -- fusions, gradient
routing, zero-copy concat layout -- against the oracle's autograd, and on the GPU box it is the per-op
reference the kernels are compared with.  It is never used by the product path.
"""
import numpy as np
import torch
import torch.nn.functional as F

import importlib

PKG = "one-stop-for-covid-19-infection-and-lung-segmentation-plus-classification_b200"
P = importlib.import_module(PKG + ".plan")
from oracle import philox

NPDT = {P.F32: np.float32, P.F16: np.float16}


class Emulator:
    def __init__(self, sizes, extra=None):
        self.mem = {k: np.zeros(int(v) + 64, np.uint8) for k, v in sizes.items()}
        for k in ("grads", "adam_m", "adam_v"):
            self.mem[k] = np.zeros(int(sizes["params"]) + 64, np.uint8)
        if extra:
            self.mem.update(extra)
        self.state = dict(seed=0, step=0, lr=5e-4, beta1=0.9, beta2=0.999, eps=1e-7, beta1_pow=0.9,
                          beta2_pow=0.999, loss_scale=1.0, grad_div=1.0, overflow=0)
        # True: conv / transposed-conv kernels whose channel counts the tcgen05 path takes (multiples of 16) are
        # rounded to fp16 first, as the GPU's packed operand copies are -- the emulator then has the SAME rounding
        # points as the fp16 engine (fp16 operands, fp32 accumulation, fp16 stores)
        self.fp16_weights = False
        # numpy Generator or None: multiplies every conv / transposed-conv accumulator by (1 + 2^-22 * N(0,1)) before it
        # is rounded to the storage type -- the size of fp32 summation-ORDER noise.  Two jittered runs differ by what
        # ANY correct fp16 implementation may differ by (rounding decisions, ReLU / max-pool flips, amplified through
        # the BatchNorm chain): the measured noise floor the GPU-vs-emulator comparison is judged against.
        self.jitter = None

    def _jit(self, a):
        if self.jitter is None:
            return a
        return (a * (1.0 + 2.384185791015625e-07 * self.jitter.standard_normal(a.shape).astype(np.float32))).astype(np.float32)

    def _kernel(self, ref, count, cin, cout, dt):
        w = self.f32(ref, count)
        if self.fp16_weights and dt == P.F16 and cin % 16 == 0 and cout % 16 == 0:
            w = w.astype(np.float16).astype(np.float32)
        return w

    # ---- raw access ------------------------------------------------------------------------
    def arr(self, ref, count, dtype):
        dtype = np.dtype(dtype)
        return np.frombuffer(self.mem[ref.arena], dtype=dtype, count=int(count), offset=ref.off)

    def view(self, ref, ld, c, npix, dt):
        """(npix, c) strided numpy view."""
        dtype = np.dtype(NPDT[dt])
        need = (npix - 1) * ld + c
        base = self.arr(ref, need, dtype)
        return np.lib.stride_tricks.as_strided(base, shape=(int(npix), int(c)), strides=(ld * dtype.itemsize, dtype.itemsize))

    def f32(self, ref, n):
        return self.arr(ref, n, np.float32)

    def f64(self, ref, n):
        return self.arr(ref, n, np.float64)

    # ---- helpers ---------------------------------------------------------------------------
    @staticmethod
    def _dact(y, act):
        y = y.astype(np.float32)
        if act == 1:
            return (y > 0).astype(np.float32)
        if act == 2:
            return np.where(y > 0, 1.0, y + 1.0).astype(np.float32)
        return np.ones_like(y)

    @staticmethod
    def _act(x, act):
        if act == 1:
            return np.maximum(x, 0)
        if act == 2:
            return np.where(x > 0, x, np.expm1(np.minimum(x, 0)))
        return x

    def _drop_factor(self, o, shape, bits_ref, op_id=None):
        """unmaterialised Dropout in front of a BatchNormalization op (f[0] = rate): keep / (1 - p) per element, or None.
        The statistics op (op_id given) draws the keep mask from the Philox stream and stores it as packed bits (bit index
        = dense element index), every other op reads those bits."""
        p = o.f[0] if len(o.f) > 0 else 0.0
        if not p:
            return None
        n = int(np.prod(shape))
        if op_id is not None:
            keep = self._keep(n, p, op_id).reshape(-1)
            self.arr(bits_ref, n // 8, np.uint8)[:] = np.packbits(keep.astype(np.uint8), bitorder="little")
        else:
            keep = np.unpackbits(self.arr(bits_ref, n // 8, np.uint8), bitorder="little")[:n].astype(np.float32)
        return keep.reshape(shape).astype(np.float32) * np.float32(1.0 / (1.0 - np.float32(p)))

    def _keep(self, nelem, p, op_id):
        return philox.dropout_keep_mask(nelem, p, self.state["seed"], self.state["step"], op_id)

    # ---- the interpreter -------------------------------------------------------------------
    def run(self, ops):
        for o in ops:
            if o.dt > 0xff:                  # executor flags (side stream / join) do not change what an op computes
                o = P.Op(o.kind, o.dt & 0xff, o.p, o.i, o.f, o.tag)
            getattr(self, "op_" + P.OP_NAMES[o.kind][3:].lower())(o)

    def op_pack_weights(self, o):
        pass        # fp16 operand copies are an implementation detail of the tensor-core kernels

    def op_memset(self, o):
        self.arr(o.p[0], o.i[0], np.uint8)[:] = 0

    def op_conv3x3_fwd(self, o):
        ldx, cin, act, ldy, cout, n, h, w = o.i[:8]
        x = self.view(o.p[0], ldx, cin, n * h * w, o.dt).astype(np.float32).reshape(n, h, w, cin)
        k_src = o.i[9] if len(o.i) > 9 else 0
        if k_src:                                        # zero-padded input: the kernel has k_src real input channels
            wt = np.zeros((3, 3, cin, cout), np.float32)
            wt[:, :, :k_src] = self._kernel(o.p[1], 9 * k_src * cout, cin, cout, o.dt).reshape(3, 3, k_src, cout)
        else:
            wt = self._kernel(o.p[1], 9 * cin * cout, cin, cout, o.dt).reshape(3, 3, cin, cout)
        b = self.f32(o.p[2], cout)
        y = F.conv2d(torch.from_numpy(x).permute(0, 3, 1, 2), torch.from_numpy(wt.copy()).permute(3, 2, 0, 1),
                     torch.from_numpy(b.copy()), padding=1).permute(0, 2, 3, 1).numpy()
        y = self._act(self._jit(y), act).reshape(-1, cout)
        if len(o.p) > 8 and o.p[7] is not None:          # inference plans: the following BatchNormalization as an affine
            y = y * self.f32(o.p[7], cout)[None, :] + self.f32(o.p[8], cout)[None, :]
        yv = self.view(o.p[3], ldy, cout, n * h * w, o.dt)
        yv[:] = y.astype(yv.dtype)
        if o.p[4] is not None:
            s = self.f64(o.p[4], 2 * cout)
            ys = yv.astype(np.float64)
            s[:cout] += ys.sum(0)
            s[cout:] += (ys * ys).sum(0)
        if len(o.p) > 6 and o.p[6] is not None:          # packed 1-bit ReLU mask of the values as stored
            packed = np.packbits((yv > 0).reshape(-1), bitorder="little")
            self.arr(o.p[6], packed.size, np.uint8)[:] = packed

    def op_conv3x3_dgrad(self, o):
        lddy, cout, lddx, cin, ldm, mact, acc, n, h, w = o.i[:10]
        dy = self.view(o.p[0], lddy, cout, n * h * w, o.dt).astype(np.float32).reshape(n, h, w, cout)
        wt = self._kernel(o.p[1], 9 * cin * cout, cin, cout, o.dt).reshape(3, 3, cin, cout)
        wo = torch.from_numpy(wt.copy()).permute(3, 2, 0, 1)           # (cout, cin, 3, 3)
        dx = F.conv_transpose2d(torch.from_numpy(dy).permute(0, 3, 1, 2), wo, padding=1).permute(0, 2, 3, 1).numpy()
        dx = self._jit(dx.reshape(-1, cin))
        if o.p[3] is not None and mact == 4:             # B2U_ACT_RELU_BITS: bit pix*cin + c
            nbits = n * h * w * cin
            bits = np.unpackbits(self.arr(o.p[3], nbits // 8, np.uint8), bitorder="little")[:nbits]
            dx = dx * bits.reshape(-1, cin).astype(np.float32)
        elif o.p[3] is not None:
            dx = dx * self._dact(self.view(o.p[3], ldm, cin, n * h * w, o.dt), mact)
        dv = self.view(o.p[2], lddx, cin, n * h * w, o.dt)
        if acc:
            dx = dx + dv.astype(np.float32)
        dv[:] = dx.astype(dv.dtype)
        self._colsum(o, 4, dv, cin)

    def _colsum(self, o, k, dv, c):
        """optional trailing pointer p[k]: per-channel sums of the values just written (a bias gradient)"""
        if len(o.p) > k and o.p[k] is not None:
            self.f32(o.p[k], c)[:] += dv.astype(np.float32).sum(0)

    def op_conv3x3_wgrad(self, o):
        ldx, cin, lddy, cout, n, h, w = o.i[:7]
        x = self.view(o.p[0], ldx, cin, n * h * w, o.dt).astype(np.float32).reshape(n, h, w, cin)
        dy = self.view(o.p[1], lddy, cout, n * h * w, o.dt).astype(np.float32).reshape(n, h, w, cout)
        xt = torch.from_numpy(x).permute(0, 3, 1, 2)
        wz = torch.zeros(cout, cin, 3, 3, requires_grad=True)
        y = F.conv2d(xt, wz, padding=1)
        (gw,) = torch.autograd.grad(y, wz, torch.from_numpy(dy).permute(0, 3, 1, 2))
        self.f32(o.p[2], 9 * cin * cout)[:] += gw.permute(2, 3, 1, 0).reshape(-1).numpy()
        if o.p[3] is not None:
            self.f32(o.p[3], cout)[:] += dy.reshape(-1, cout).sum(0)

    def op_convt_fwd(self, o):
        ldx, cin, ldy, cout, n, h, w = o.i[:7]
        x = self.view(o.p[0], ldx, cin, n * h * w, o.dt).astype(np.float32).reshape(n, h, w, cin)
        wt = self._kernel(o.p[1], 4 * cout * cin, cin, cout, o.dt).reshape(2, 2, cout, cin)
        b = self.f32(o.p[2], cout)
        y = F.conv_transpose2d(torch.from_numpy(x).permute(0, 3, 1, 2), torch.from_numpy(wt.copy()).permute(3, 2, 0, 1),
                               torch.from_numpy(b.copy()), stride=2).permute(0, 2, 3, 1).numpy()
        yv = self.view(o.p[3], ldy, cout, n * 4 * h * w, o.dt)
        yv[:] = self._jit(y.reshape(-1, cout)).astype(yv.dtype)
        if len(o.p) > 4 and o.p[4] is not None:          # statistics for a following BN over the concat buffer
            sq = o.i[7]
            s = self.f64(o.p[4], sq + cout)
            ys = yv.astype(np.float64)
            s[:cout] += ys.sum(0)
            s[sq:sq + cout] += (ys * ys).sum(0)

    def op_convt_dgrad(self, o):
        lddy, cout, lddx, cin, ldm, mact, acc, n, h, w = o.i[:10]
        dy = self.view(o.p[0], lddy, cout, n * 4 * h * w, o.dt).astype(np.float32).reshape(n, 2 * h, 2 * w, cout)
        wt = self._kernel(o.p[1], 4 * cout * cin, cin, cout, o.dt).reshape(2, 2, cout, cin)
        dx = F.conv2d(torch.from_numpy(dy).permute(0, 3, 1, 2), torch.from_numpy(wt.copy()).permute(3, 2, 0, 1),
                      stride=2).permute(0, 2, 3, 1).numpy().reshape(-1, cin)
        dx = self._jit(dx)
        if o.p[3] is not None:
            dx = dx * self._dact(self.view(o.p[3], ldm, cin, n * h * w, o.dt), mact)
        dv = self.view(o.p[2], lddx, cin, n * h * w, o.dt)
        if acc:
            dx = dx + dv.astype(np.float32)
        dv[:] = dx.astype(dv.dtype)
        self._colsum(o, 4, dv, cin)

    def op_convt_wgrad(self, o):
        ldx, cin, lddy, cout, n, h, w = o.i[:7]
        x = self.view(o.p[0], ldx, cin, n * h * w, o.dt).astype(np.float32).reshape(n, h, w, cin)
        dy = self.view(o.p[1], lddy, cout, n * 4 * h * w, o.dt).astype(np.float32).reshape(n, 2 * h, 2 * w, cout)
        # dw[a,b,co,ci] = sum x[n,i,j,ci] dy[n,2i+a,2j+b,co]
        d5 = dy.reshape(n, h, 2, w, 2, cout)
        gw = np.einsum("nijc,niajbo->aboc", x, d5)
        self.f32(o.p[2], 4 * cout * cin)[:] += gw.reshape(-1)
        if o.p[3] is not None:
            self.f32(o.p[3], cout)[:] += dy.reshape(-1, cout).sum(0)

    def op_bn_stats(self, o):
        ldx, c, npix = o.i[:3]
        sq = o.i[3] if len(o.i) > 3 and o.i[3] else c        # offset of the squares (one half of a split concatenate)
        x = self.view(o.p[0], ldx, c, npix, o.dt).astype(np.float32)
        f = self._drop_factor(o, x.shape, o.p[3] if len(o.p) > 3 else None, o.i[4] if len(o.i) > 4 else 0)
        if f is not None:
            x = x * f
        x = x.astype(np.float64)
        s = self.f64(o.p[1], sq + c)
        s[:c] += x.sum(0)
        s[sq:sq + c] += (x * x).sum(0)

    def op_bn_finalize(self, o):
        count, training, c = o.i[:3]
        mom, eps = o.f[:2]
        g, b = self.f32(o.p[1], c), self.f32(o.p[2], c)
        mm, mv = self.f32(o.p[3], c), self.f32(o.p[4], c)
        if training:
            s = self.f64(o.p[0], 2 * c)
            mean = s[:c] / count
            var = np.maximum(s[c:] / count - mean * mean, 0)
            unb = var * (count / (count - (1.0 + eps)))
            mm[:] = np.float32(mom) * mm + np.float32(1 - mom) * mean.astype(np.float32)
            mv[:] = np.float32(mom) * mv + np.float32(1 - mom) * unb.astype(np.float32)
            mean, var = mean.astype(np.float32), var.astype(np.float32)
        else:
            mean, var = mm.copy(), mv.copy()
        inv = (1.0 / np.sqrt(var.astype(np.float64) + eps)).astype(np.float32)
        sc = g * inv
        self.f32(o.p[5], c)[:] = sc
        self.f32(o.p[6], c)[:] = b - mean * sc
        self.f32(o.p[7], c)[:] = mean
        self.f32(o.p[8], c)[:] = inv

    def op_bn_apply(self, o):
        ldx, ldy, c, npix = o.i[:4]
        if len(o.p) > 5 and o.p[5] is not None and o.kind == P.OP_BN_APPLY:      # split concatenate: second source tensor
            split, ldx2 = o.i[5], o.i[6]
            x = np.concatenate([self.view(o.p[0], ldx, split, npix, o.dt).astype(np.float32),
                                self.view(o.p[5], ldx2, c - split, npix, o.dt).astype(np.float32)], axis=1)
        else:
            x = self.view(o.p[0], ldx, c, npix, o.dt).astype(np.float32)
        if o.kind == P.OP_BN_APPLY:
            f = self._drop_factor(o, x.shape, o.p[6] if len(o.p) > 6 else None)
            if f is not None:
                x = x * f
        y = self.view(o.p[1], ldy, c, npix, o.dt)
        y[:] = (x * self.f32(o.p[2], c) + self.f32(o.p[3], c)).astype(y.dtype)
        if len(o.p) > 4 and o.p[4] is not None:
            sq = o.i[4]
            s = self.f64(o.p[4], sq + c)
            ys = y.astype(np.float64)
            s[:c] += ys.sum(0)
            s[sq:sq + c] += (ys * ys).sum(0)

    def op_bn_apply_pool(self, o):
        ldx, ldy, c, npix, sq, n, h, w, ldp, op_id = o.i[:10]
        self.op_bn_apply(P.Op(P.OP_BN_APPLY, o.dt, o.p[:5], [ldx, ldy, c, npix, sq]))
        self.op_maxpool_fwd(P.Op(P.OP_MAXPOOL_FWD, o.dt, [o.p[1], o.p[5], o.p[6]], [ldy, ldp, c, n, h, w, op_id], [o.f[0]]))

    def op_bn_bwd_sums_wgrad(self, o):
        c, cout, taps = o.i[:3]
        w = self.f32(o.p[0], taps * c * cout).reshape(taps, c, cout).astype(np.float64)
        dw = self.f32(o.p[1], taps * c * cout).reshape(taps, c, cout).astype(np.float64)
        cs = self.f32(o.p[2], c).astype(np.float64)
        g, b = self.f32(o.p[3], c).astype(np.float64), self.f32(o.p[4], c).astype(np.float64)
        s = self.f64(o.p[5], 2 * c)
        s[:c] += cs
        s[c:] += np.where(np.abs(g) > 1e-12, ((w * dw).sum((0, 2)) - b * cs) / np.where(np.abs(g) > 1e-12, g, 1.0), 0.0)

    def op_bn_bwd_reduce(self, o):
        lddy, ldx, c, npix = o.i[:4]
        dy = self.view(o.p[0], lddy, c, npix, o.dt).astype(np.float32)
        x = self.view(o.p[1], ldx, c, npix, o.dt).astype(np.float32)
        f = self._drop_factor(o, x.shape, o.p[5] if len(o.p) > 5 else None)
        if f is not None:
            x = x * f
        xh = (x - self.f32(o.p[2], c)) * self.f32(o.p[3], c)
        sq = o.i[4] if len(o.i) > 4 and o.i[4] else c        # offset of the second sums (one half of a split concatenate)
        s = self.f64(o.p[4], sq + c)
        s[:c] += dy.astype(np.float64).sum(0)
        s[sq:sq + c] += (dy * xh).astype(np.float64).sum(0)

    def op_bn_bwd_apply(self, o):
        lddy, ldx, lddx, c, npix, ldm, mact, count = o.i[:8]
        dy = self.view(o.p[0], lddy, c, npix, o.dt).astype(np.float32)
        two = len(o.p) > 12 and o.p[11] is not None                             # split concatenate: second (x, dx) pair
        if two:
            split, ldx2, lddx2 = o.i[8], o.i[9], o.i[10]
            x = np.concatenate([self.view(o.p[1], ldx, split, npix, o.dt).astype(np.float32),
                                self.view(o.p[11], ldx2, c - split, npix, o.dt).astype(np.float32)], axis=1)
        else:
            x = self.view(o.p[1], ldx, c, npix, o.dt).astype(np.float32)
        g, mean, inv = self.f32(o.p[3], c), self.f32(o.p[4], c), self.f32(o.p[5], c)
        s = self.f64(o.p[6], 2 * c)
        f = self._drop_factor(o, x.shape, o.p[13] if len(o.p) > 13 else None)
        if f is not None:
            x = x * f
        xh = (x - mean) * inv
        dx = g * inv * (dy - (s[:c] / count).astype(np.float32) - xh * (s[c:] / count).astype(np.float32))
        if f is not None:
            dx = dx * f                                    # the dropout backward
        if o.p[9] is not None:
            dx = dx * self._dact(self.view(o.p[9], ldm, c, npix, o.dt), mact)
        if two:
            da, db_ = self.view(o.p[2], lddx, split, npix, o.dt), self.view(o.p[12], lddx2, c - split, npix, o.dt)
            da[:] = dx[:, :split].astype(da.dtype)
            db_[:] = dx[:, split:].astype(db_.dtype)
            dv = np.concatenate([da, db_], axis=1)
        else:
            dv = self.view(o.p[2], lddx, c, npix, o.dt)
            dv[:] = dx.astype(dv.dtype)
        self._colsum(o, 10, dv, c)
        if o.p[7] is not None:
            self.f32(o.p[7], c)[:] += s[c:].astype(np.float32)
            self.f32(o.p[8], c)[:] += s[:c].astype(np.float32)

    def _pool_windows(self, ref, ld, c, n, h, w, dt):
        x = self.view(ref, ld, c, n * h * w, dt).astype(np.float32).reshape(n, h // 2, 2, w // 2, 2, c)
        return x.transpose(0, 1, 3, 2, 4, 5).reshape(n, h // 2, w // 2, 4, c)   # window order (0,0),(0,1),(1,0),(1,1)

    def op_maxpool_fwd(self, o):
        ldx, ldy, c, n, h, w, op_id = o.i[:7]
        p = o.f[0]
        win = self._pool_windows(o.p[0], ldx, c, n, h, w, o.dt)
        y = win.max(axis=3).reshape(-1, c)
        if p > 0:
            keep = self._keep(y.size, p, op_id).reshape(y.shape)
            y = y * keep * np.float32(1.0 / (1.0 - np.float32(p)))
        yv = self.view(o.p[1], ldy, c, n * (h // 2) * (w // 2), o.dt)
        yv[:] = y.astype(yv.dtype)

    def op_maxpool_bwd(self, o):
        ldx, lddy, lddx, c, n, h, w, op_id, acc = o.i[:9]
        p = o.f[0]
        win = self._pool_windows(o.p[0], ldx, c, n, h, w, o.dt)
        gy = self.view(o.p[1], lddy, c, n * (h // 2) * (w // 2), o.dt).astype(np.float32)
        if p > 0:
            keep = self._keep(gy.size, p, op_id).reshape(gy.shape)
            gy = gy * keep * np.float32(1.0 / (1.0 - np.float32(p)))
        sel = win.argmax(axis=3)                                        # first maximum
        g = np.zeros_like(win)
        np.put_along_axis(g, sel[:, :, :, None, :], gy.reshape(n, h // 2, w // 2, 1, c), axis=3)
        g = g.reshape(n, h // 2, w // 2, 2, 2, c).transpose(0, 1, 3, 2, 4, 5).reshape(-1, c)
        dv = self.view(o.p[2], lddx, c, n * h * w, o.dt)
        if acc:
            g = g + dv.astype(np.float32)
        dv[:] = g.astype(dv.dtype)
        if len(o.p) > 4 and o.p[4] is not None:          # fused BN-backward statistics (x = gamma*xhat + beta)
            gm, bt = self.f32(o.p[5], c), self.f32(o.p[6], c)
            xs = self.view(o.p[0], ldx, c, n * h * w, o.dt).astype(np.float32)
            rg = np.where(np.abs(gm) > 1e-12, 1.0 / np.where(gm == 0, 1, gm), 0.0).astype(np.float32)
            xh = (xs - bt) * rg
            s = self.f64(o.p[4], 2 * c)
            s[:c] += g.astype(np.float64).sum(0)
            s[c:] += (g * xh).astype(np.float64).sum(0)

    def op_dropout_fwd(self, o):
        ldx, ldy, c, npix, op_id = o.i[:5]
        p = o.f[0]
        x = self.view(o.p[0], ldx, c, npix, o.dt).astype(np.float32)
        keep = self._keep(x.size, p, op_id).reshape(x.shape)
        yv = self.view(o.p[1], ldy, c, npix, o.dt)
        yv[:] = (x * keep * np.float32(1.0 / (1.0 - np.float32(p)))).astype(yv.dtype)

    def op_dropout_bwd(self, o):
        lddy, lddx, c, npix, op_id, ldm, mact = o.i[:7]
        p = o.f[0]
        gy = self.view(o.p[0], lddy, c, npix, o.dt).astype(np.float32)
        keep = self._keep(gy.size, p, op_id).reshape(gy.shape)
        g = gy * keep * np.float32(1.0 / (1.0 - np.float32(p)))
        if o.p[3] is not None:
            g = g * self._dact(self.view(o.p[3], ldm, c, npix, o.dt), mact)
        dv = self.view(o.p[1], lddx, c, npix, o.dt)
        dv[:] = g.astype(dv.dtype)

    def op_copy_slice(self, o):
        lds, ldd, c, npix, acc = o.i[:5]
        s = self.view(o.p[0], lds, c, npix, o.dt)
        d = self.view(o.p[1], ldd, c, npix, o.dt)
        if acc:
            d[:] = (d.astype(np.float32) + s.astype(np.float32)).astype(d.dtype)
        else:
            d[:] = s

    def op_head_fwd(self, o):
        ldx, cin, npix = o.i[:3]
        x = self.view(o.p[0], ldx, cin, npix, o.dt).astype(np.float32)
        z = x @ self.f32(o.p[1], cin) + self.f32(o.p[2], 1)[0]
        self.f32(o.p[3], npix)[:] = 1.0 / (1.0 + np.exp(-z.astype(np.float64)))

    @staticmethod
    def _bce(t, p):
        ph = np.clip(p.astype(np.float64), 1e-7, 1 - 1e-7)
        return -(t * np.log(ph) + (1 - t) * np.log1p(-ph))

    def op_bce_dice_sums(self, o):
        cnt = o.i[0]
        p, t = self.f32(o.p[0], cnt).astype(np.float64), self.f32(o.p[1], cnt).astype(np.float64)
        s = self.f64(o.p[2], 4)
        s[0] += (t * p).sum(); s[1] += t.sum(); s[2] += p.sum(); s[3] += self._bce(t, p).sum()

    def op_bce_dice_finalize(self, o):
        s = self.f64(o.p[0], 4)
        dice = (2 * s[0] + 1) / (s[1] + s[2] + 1)
        out = self.f32(o.p[1], 2)
        out[0] = 0.5 * s[3] / o.i[0] + 0.5 * (1 - dice)
        out[1] = dice

    def op_head_bwd(self, o):
        count, ldx, cin, lddx, xact, npix = o.i[:6]
        p, t = self.f32(o.p[0], npix).astype(np.float64), self.f32(o.p[1], npix).astype(np.float64)
        s = self.f64(o.p[2], 4)
        I, S = s[0], s[1] + s[2]
        inside = (p >= 1e-7) & (p <= 1 - 1e-7)
        with np.errstate(divide="ignore", invalid="ignore"):
            g = np.where(inside, 0.5 / count * (-t / p + (1 - t) / (1 - p)), 0.0)
        g = g - 0.5 * (2 * t * (S + 1) - (2 * I + 1)) / (S + 1) ** 2
        dl = (g * p * (1 - p) * self.state["loss_scale"]).astype(np.float32)
        x = self.view(o.p[4], ldx, cin, npix, o.dt).astype(np.float32)
        w = self.f32(o.p[5], cin)
        dx = dl[:, None] * w[None, :] * self._dact(x, xact)
        dv = self.view(o.p[6], lddx, cin, npix, o.dt)
        dv[:] = dx.astype(dv.dtype)
        self._colsum(o, 9, dv, cin)
        self.f32(o.p[7], cin)[:] += (dl[:, None] * x).sum(0)
        self.f32(o.p[8], 1)[:] += dl.sum()

    def op_dense_fwd(self, o):
        k, act, m, n = o.i[:4]
        x = self.view(o.p[0], k, k, n, o.dt).astype(np.float32)
        z = x @ self.f32(o.p[1], k * m).reshape(k, m) + self.f32(o.p[2], m)
        y = 1.0 / (1.0 + np.exp(-z)) if act == 3 else self._act(z, act)
        self.f32(o.p[3], n * m)[:] = y.reshape(-1)

    def op_dense_bwd(self, o):
        k, act, mact, m, n = o.i[:5]
        x = self.view(o.p[0], k, k, n, o.dt).astype(np.float32)
        w = self.f32(o.p[1], k * m).reshape(k, m)
        y = self.f32(o.p[2], n * m).reshape(n, m)
        dpre = self.f32(o.p[3], n * m).reshape(n, m) * self._dact(y, act)
        if o.p[4] is not None:
            dx = dpre @ w.T
            if o.p[5] is not None:
                dx = dx * self._dact(self.view(o.p[5], k, k, n, o.dt), mact)
            dv = self.view(o.p[4], k, k, n, o.dt)
            dv[:] = dx.astype(dv.dtype)
        self.f32(o.p[6], k * m)[:] += (x.T @ dpre).reshape(-1)
        self.f32(o.p[7], m)[:] += dpre.sum(0)

    def op_bce_fwd(self, o):
        n = o.i[0]
        p, t = self.f32(o.p[0], n), self.f32(o.p[1], n)
        sw = self.f32(o.p[2], n) if o.p[2] is not None else np.ones(n, np.float32)
        self.f32(o.p[3], 1)[0] = (sw * self._bce(t.astype(np.float64), p)).mean()

    def op_bce_sigmoid_bwd(self, o):
        n = o.i[0]
        p, t = self.f32(o.p[0], n).astype(np.float64), self.f32(o.p[1], n).astype(np.float64)
        sw = self.f32(o.p[2], n) if o.p[2] is not None else np.ones(n, np.float32)
        inside = (p >= 1e-7) & (p <= 1 - 1e-7)
        with np.errstate(divide="ignore", invalid="ignore"):
            g = np.where(inside, (-t / p + (1 - t) / (1 - p)) * p * (1 - p), 0.0)
        out = self.arr(o.p[4], n, NPDT[o.dt])
        out[:] = (g * sw / n * self.state["loss_scale"]).astype(out.dtype)

    def op_adam(self, o):
        n = o.i[0]
        st = self.state
        p, g, m, v = (self.f32(o.p[k], n) for k in range(4))
        if not np.isfinite(g).all():              # whole step skipped (b2u_adam's check pass), see op_state_advance
            st["skip_step"] = 1
            return
        gi = g / np.float32(st["loss_scale"] * st["grad_div"])
        lr_t = np.float32(st["lr"] * np.sqrt(1 - st["beta2_pow"]) / (1 - st["beta1_pow"]))
        m[:] = np.float32(st["beta1"]) * m + np.float32(1 - st["beta1"]) * gi
        v[:] = np.float32(st["beta2"]) * v + np.float32(1 - st["beta2"]) * gi * gi
        p[:] = p - lr_t * m / (np.sqrt(v) + np.float32(st["eps"]))

    def op_state_advance(self, o):
        if self.state.get("skip_step"):
            self.state.update(skip_step=0, overflow=1, loss_scale=max(self.state["loss_scale"] * 0.5, 1.0))
            return
        self.state["step"] += 1
        self.state["beta1_pow"] *= self.state["beta1"]
        self.state["beta2_pow"] *= self.state["beta2"]

    def _allreduce(self, o, dtype):
        # single-process emulation: identity; under torch.distributed (gloo): the real exchange
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            a = self.arr(o.p[0], o.i[0], dtype)
            t = torch.from_numpy(a.copy())
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            a[:] = t.numpy()

    def op_allreduce_f32(self, o):
        self._allreduce(o, np.float32)

    def op_allreduce_f64(self, o):
        self._allreduce(o, np.float64)
