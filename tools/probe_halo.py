"""GPU probe (not a test): which halo-tile variants of the tcgen05 3x3 conv (csrc/conv_tc3.cu) reproduce the
emulator, and how fast each is on the thin full-resolution layers.  Usage: python tools/probe_halo.py"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import emulator as E  # noqa: E402
from gpu_harness import LIB, run_ops_gpu  # noqa: E402
from helpers import P  # noqa: E402
from test_gpu_ops import Img  # noqa: E402

dt = P.F16
lib = LIB.lib()
SHAPES = [(1, 16, 8, 64, 64), (2, 32, 32, 32, 32), (1, 40, 24, 64, 32), (1, 32, 32, 256, 512), (1, 20, 12, 128, 64),
          (1, 56, 56, 16, 16), (1, 16, 16, 96, 32)]


def check(mode, n, h, w, cin, cout, dgrad):
    lib.b2u_set_option(b"tc_halo", mode)
    img = Img(41)
    x = img.view(n, h, w, cin, dt, ld=2 * cin, c0=cin, fill="uniform")
    y = img.view(n, h, w, cout, dt, ld=cout + 16, c0=8, fill=None)
    wt = img.farr(img.par, 9 * cin * cout, scale=(2.0 / (9 * cin)) ** 0.5)
    b = img.farr(img.par, cout, scale=0.1)
    stats = img.zero.alloc(2 * cout * 8)
    if dgrad:
        ops = [P.Op(P.OP_CONV3X3_DGRAD, dt, [x.ref, wt, y.ref, None], [x.ld, cin, y.ld, cout, 0, 0, 0, n, h, w])]
        # dgrad reads a tensor with `cin` channels as dy and writes `cout` channels: weights are (3,3,cout,cin)
    else:
        ops = [P.Op(P.OP_CONV3X3_FWD, dt, [x.ref, wt, b, y.ref, stats], [x.ld, cin, 1, y.ld, cout, n, h, w])]
    mem = img.mem()
    em = E.Emulator({"act": 0, "f32": 0, "zero": 0, "params": 0, "state": 0, "step": 0})
    em.mem = {k: v.copy() for k, v in mem.items()}
    em.run(ops)
    out, _ = run_ops_gpu(ops, mem, dict(em.state))
    nel = (len(mem["act"]) - 256) // 2
    g = np.frombuffer(out["act"], np.float16, nel).astype(np.float64)
    r = np.frombuffer(em.mem["act"], np.float16, nel).astype(np.float64)
    return float(np.abs(g - r).max() / max(np.abs(r).max(), 1e-6))


def bench(mode, n, h, w, cin, cout, reps=20):
    lib.b2u_set_option(b"tc_halo", mode)
    x = torch.rand(n, h, w, cin, device="cuda").half()
    y = torch.empty(n, h, w, cout, device="cuda", dtype=torch.float16)
    wt = torch.randn(3, 3, cin, cout, device="cuda") * 0.05
    b = torch.zeros(cout, device="cuda")
    ws = torch.empty(int(lib.b2u_ws_bytes()), dtype=torch.uint8, device="cuda")
    s = torch.cuda.Stream()
    sp = C.c_void_p(s.cuda_stream)

    def run():
        LIB.check(lib.b2u_conv3x3_fwd(1, x.data_ptr(), cin, cin, wt.data_ptr(), b.data_ptr(), 1, y.data_ptr(), cout, cout,
                                      None, n, h, w, ws.data_ptr(), ws.numel(), sp))
    for _ in range(3):
        run()
    s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(reps):
        run()
    e1.record(s)
    s.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2 * 9 * cin * cout * n * h * w
    by = n * h * w * (cin + cout) * 2
    return ms, fl / ms / 1e9, by / ms / 1e6


if __name__ == "__main__":
    ok = {}
    for mode in (0, 1, 2, 3):
        errs = []
        for shp in SHAPES:
            for dg in (0, 1):
                try:
                    errs.append(check(mode, *shp, dg))
                except Exception as ex:  # noqa: BLE001
                    errs.append(float("nan"))
                    print("mode", mode, shp, "dgrad" if dg else "fwd", "ERROR", str(ex)[:200])
        ok[mode] = all(e == e and e < 4e-3 for e in errs)
        print("mode %d: %s  max rel errs %s" % (mode, "OK" if ok[mode] else "MISMATCH", ["%.1e" % e for e in errs]))
    for shp in [(8, 512, 512, 32, 32), (8, 512, 512, 64, 32), (8, 256, 256, 64, 64), (8, 256, 256, 128, 64),
                (8, 128, 128, 128, 128), (8, 64, 64, 256, 256), (8, 32, 32, 512, 512)]:
        for mode in (0, 1, 2, 3):
            if ok[mode]:
                ms, tf, gbs = bench(mode, *shp)
                print("shape %s mode %d: %.4f ms  %.1f TFLOP/s  %.0f GB/s (algorithmic)" % (shp, mode, ms, tf / 1e3, gbs / 1e3))
    lib.b2u_set_option(b"tc_halo", 1)
