"""GPU probe (not a test): what does it cost that a tensor lives in one half of a 128-byte concat pixel?
The level-1 ops of the U-Net step that read / write the skip tensor or the transposed-conv output (32 of 64 channels at
512 x 512, batch 8), timed with the tensor strided (ld = 64, as placed today) and dense (ld = 32).
  python tools/half_line_probe.py"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_harness import LIB, P  # noqa: E402

lib = LIB.lib()
n, h, w, c = 8, 512, 512, 32
npix = n * h * w
dt = P.F16
ws = torch.empty(int(lib.b2u_ws_bytes()), dtype=torch.uint8, device="cuda")
stream = torch.cuda.Stream()


class R:
    def __init__(self, t, off=0):
        self.a = t.data_ptr() + off


def timed(op, reps=20):
    arr = LIB.make_ops([op], lambda r: r.a)
    run = lambda: LIB.check(lib.b2u_run_ops(arr, 1, C.c_void_p(ws.data_ptr()), ws.numel(), None, C.c_void_p(stream.cuda_stream)), "run")
    for _ in range(3):
        run()
    stream.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        run()
    e1.record(stream)
    stream.synchronize()
    return e0.elapsed_time(e1) / reps


flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for ld in (64, 32):
    y = (torch.rand(npix, ld, device="cuda") - 0.3).half()            # skip tensor (BN output) in its buffer
    g = torch.randn(npix, ld, device="cuda").half() * 0.1              # its gradient buffer
    gp = torch.randn(npix // 4, 64, device="cuda").half() * 0.1        # pooled gradient (dense 32 of... own tensor)
    x = torch.randn(npix, c, device="cuda").half()                     # BN input (dense)
    dx = torch.empty(npix, c, device="cuda", dtype=torch.float16)
    gamma = torch.ones(c, device="cuda"); beta = torch.zeros(c, device="cuda")
    mean = torch.zeros(c, device="cuda"); invstd = torch.ones(c, device="cuda")
    sums = torch.zeros(2 * c, device="cuda", dtype=torch.float64)
    dg = torch.zeros(c, device="cuda"); db = torch.zeros(c, device="cuda")
    off = (ld - c) * 2                                                  # upper half when strided
    res = {}
    res["maxpool_bwd L1"] = timed(P.Op(P.OP_MAXPOOL_BWD, dt, [R(y, off), R(gp), R(g, off), None, R(sums), R(gamma), R(beta)],
                                       [ld, 64, ld, c, n, h, w, 0, 1], [0.0]))
    res["bn_bwd_apply L1 (dy in the buffer)"] = timed(P.Op(P.OP_BN_BWD_APPLY, dt, [R(g, off), R(x), R(dx), R(gamma), R(mean), R(invstd),
                                                                                  R(sums), R(dg), R(db), None, None],
                                                           [ld, c, c, c, npix, 0, 0, npix]))
    # transposed conv 64 -> 32 at 256^2 -> 512^2: forward writes, data / weight gradient read the buffer half
    xs = torch.randn(npix // 4, 64, device="cuda").half()
    wt = torch.randn(4 * 32 * 64, device="cuda") * 0.05
    bias = torch.zeros(32, device="cuda")
    gxs = torch.empty(npix // 4, 64, device="cuda", dtype=torch.float16)
    dw = torch.zeros(4 * 32 * 64, device="cuda"); dbias = torch.zeros(32, device="cuda")
    res["convt_fwd L1"] = timed(P.Op(P.OP_CONVT_FWD, dt, [R(xs), R(wt), R(bias), R(y), None, None], [64, 64, ld, c, n, h // 2, w // 2, 0]))
    res["convt_dgrad L1"] = timed(P.Op(P.OP_CONVT_DGRAD, dt, [R(g), R(wt), R(gxs), None, None, None],
                                       [ld, c, 64, 64, 0, 0, 0, n, h // 2, w // 2]))
    res["convt_wgrad L1"] = timed(P.Op(P.OP_CONVT_WGRAD, dt, [R(xs), R(g), R(dw), R(dbias)], [64, 64, ld, c, n, h // 2, w // 2]))
    res["bn_apply_pool L1 (writes the skip)"] = timed(P.Op(P.OP_BN_APPLY_POOL, dt, [R(x), R(y, off), R(gamma), R(beta), None, R(gp), None],
                                                           [c, ld, c, npix, 0, n, h, w, 64, 0], [0.0]))
    print("ld = %d (%s):" % (ld, "strided half of a 64-channel buffer" if ld == 64 else "dense"))
    for k, v in res.items():
        print("   %-40s %.4f ms" % (k, v))
