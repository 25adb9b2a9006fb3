"""run ONE transposed-conv op of the U-Net step a few times (CUDA-event time); under ncu this is the process to profile.
usage: one_convt.py KIND N H W CIN COUT [stats]     KIND = fwd | dgrad | wgrad   (H, W = the SMALL image)"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_harness import LIB, P  # noqa: E402

kind = sys.argv[1]
n, h, w, cin, cout = [int(v) for v in sys.argv[2:7]]
flags = sys.argv[7:]
lib = LIB.lib()
for f in flags:
    if f.startswith("opt:"):
        lib.b2u_set_option(f[4:].split("=")[0].encode(), int(f.split("=")[1]))
npix = n * h * w
x = (torch.rand(npix, cin, device="cuda") - 0.3).half()
y = torch.empty(4 * npix, cout, device="cuda", dtype=torch.float16)
gy = torch.randn(4 * npix, cout, device="cuda").half() * 0.1
gx = torch.empty(npix, cin, device="cuda", dtype=torch.float16)
wt = torch.randn(4 * cin * cout, device="cuda") * 0.05
bias = torch.zeros(cout, device="cuda")
stats = torch.zeros(4 * cout, device="cuda", dtype=torch.float64)
colsum = torch.zeros(cin, device="cuda")
dw = torch.zeros(4 * cin * cout, device="cuda")
db = torch.zeros(cout, device="cuda")
ws = torch.empty(int(lib.b2u_ws_bytes()), dtype=torch.uint8, device="cuda")


class R:
    def __init__(self, t):
        self.a = t.data_ptr()


if kind == "fwd":
    op = P.Op(P.OP_CONVT_FWD, P.F16, [R(x), R(wt), R(bias), R(y), R(stats) if "stats" in flags else None, None],
              [cin, cin, cout, cout, n, h, w, 2 * cout])
    by = npix * cin * 2 + 4 * npix * cout * 2
elif kind == "dgrad":
    op = P.Op(P.OP_CONVT_DGRAD, P.F16, [R(gy), R(wt), R(gx), R(x) if "mask" in flags else None, R(colsum) if "colsum" in flags else None, None],
              [cout, cout, cin, cin, cin, 1 if "mask" in flags else 0, 0, n, h, w])
    by = npix * cin * 2 * (2 if "mask" in flags else 1) + 4 * npix * cout * 2
else:
    op = P.Op(P.OP_CONVT_WGRAD, P.F16, [R(x), R(gy), R(dw), R(db)], [cin, cin, cout, cout, n, h, w])
    by = npix * cin * 2 + 4 * npix * cout * 2
arr = LIB.make_ops([op], lambda r: r.a)
stream = torch.cuda.Stream()
run = lambda: LIB.check(lib.b2u_run_ops(arr, 1, C.c_void_p(ws.data_ptr()), ws.numel(), None, C.c_void_p(stream.cuda_stream)), "run_ops")
for _ in range(3):
    run()
stream.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(10):
    run()
e1.record(stream)
stream.synchronize()
ms = e0.elapsed_time(e1) / 10
print("%s: %.4f ms per call (packs weights per call), %.0f GB/s algorithmic" % (" ".join(sys.argv[1:]), ms, by / ms / 1e6))
