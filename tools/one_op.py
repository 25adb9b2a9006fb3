"""run ONE conv op of the U-Net step (forward / data gradient, with its real epilogue features) a few times, time it with
CUDA events and dump the in-kernel timeline of CTA 0 (b2u_set_option("tc_debug")).
usage: one_op.py KIND N H W K J [stats|colsum|bits|dwmerge ...]      KIND = fwd | dgrad
  fwd   : x (K ch) -> y (J ch), ReLU; 'stats' adds BN statistics, 'bits' writes the 1-bit ReLU mask
  dgrad : dy (K ch) -> dx (J ch);     'colsum' adds column sums, 'bits' applies a 1-bit mask"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_harness import LIB, P  # noqa: E402

kind = sys.argv[1]
n, h, w, K, J = [int(v) for v in sys.argv[2:7]]
flags = sys.argv[7:]
lib = LIB.lib()
for f in flags:                                   # opt:name=value -> b2u_set_option(name, value)
    if f.startswith("opt:"):
        lib.b2u_set_option(f[4:].split("=")[0].encode(), int(f.split("=")[1]))
lib.b2u_set_option(b"tc_dwmerge", 1 if "dwmerge" in flags else 0)
lib.b2u_set_option(b"tc_rowstrip", 1 if "rowstrip" in flags else 0)
npix = n * h * w
x = (torch.rand(npix, K, device="cuda") - 0.3).half()
y = torch.empty(npix, J, device="cuda", dtype=torch.float16)
wt = torch.randn(9 * K * J, device="cuda") * 0.05
bias = torch.zeros(J, device="cuda")
stats = torch.zeros(2 * J, device="cuda", dtype=torch.float64)
colsum = torch.zeros(J, device="cuda")
bits = torch.randint(0, 255, (npix * J // 8,), device="cuda", dtype=torch.uint8)
ws = torch.empty(int(lib.b2u_ws_bytes()), dtype=torch.uint8, device="cuda")
wp = torch.empty(9 * K * J, device="cuda", dtype=torch.float16)
ptr = lambda t: t.data_ptr()


class R:                       # absolute "refs": the resolver below is the identity
    def __init__(self, a):
        self.a = a


if kind == "fwd":
    op = P.Op(P.OP_CONV3X3_FWD, P.F16, [R(ptr(x)), R(ptr(wt)), R(ptr(bias)), R(ptr(y)), R(ptr(stats)) if "stats" in flags else None,
                                        None, R(ptr(bits)) if "bits" in flags else None], [K, K, 1, J, J, n, h, w])
else:
    op = P.Op(P.OP_CONV3X3_DGRAD, P.F16, [R(ptr(x)), R(ptr(wt)), R(ptr(y)), R(ptr(bits)) if "bits" in flags else None,
                                          R(ptr(colsum)) if "colsum" in flags else None, None],
              [K, K, J, J, J, P.ACT_RELU_BITS if "bits" in flags else 0, 0, n, h, w])
arr = LIB.make_ops([op], lambda r: r.a)
stream = torch.cuda.Stream()


def run():
    LIB.check(lib.b2u_run_ops(arr, 1, C.c_void_p(ws.data_ptr()), ws.numel(), None, C.c_void_p(stream.cuda_stream)), "run_ops")


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
by = npix * (K + J) * 2 + (npix * J // 8 if "bits" in flags else 0)
print("%s: %.4f ms per call (packs weights per call), %.0f GB/s algorithmic, %.1f clk/px/SM at 1.9 GHz"
      % (" ".join(sys.argv[1:]), ms, by / ms / 1e6, ms * 1e-3 * 1.9e9 * 148 / npix))
if "notimeline" in flags:
    sys.exit(0)
if "rowstrip" in flags:
    lib.b2u_set_option(b"tc_debug", 1)
    run()
    stream.synchronize()
    buf = (C.c_longlong * 512)()
    lib.b2u_debug_read.argtypes = [C.POINTER(C.c_longlong), C.c_int]
    assert lib.b2u_debug_read(buf, 512) == 0
    a = np.array(buf[:], dtype=np.int64).reshape(64, 8)
    t0 = a[0, 0]
    print("row: tma_issue | mma_start mma_acc_empty mma_a_full mma_commit | epi_start(even rows) epi_landed epi_prefetched")
    for k in range(8, 28):
        print(k, [int(v - t0) if v else None for v in a[k]])
    lib.b2u_set_option(b"tc_debug", 0)
    sys.exit(0)
lib.b2u_set_option(b"tc_debug", 1)
run()
stream.synchronize()
buf = (C.c_longlong * 512)()
lib.b2u_debug_read.argtypes = [C.POINTER(C.c_longlong), C.c_int]
assert lib.b2u_debug_read(buf, 512) == 0
a = np.array(buf[:], dtype=np.int64).reshape(64, 8)
t0 = a[0, 0] if a[0, 0] else a[0, 2]
d = lambda k, i, j: int(a[k, i] - a[k, j]) if a[k, i] and a[k, j] else None
first = int(os.environ.get("TL_FIRST", "6"))        # steady state: TL_FIRST=36 (the first tiles run on an idle memory system)
rows = range(first, first + 16)
if os.environ.get("TL_DUMP"):
    print("iter: prod_wait prod_issued | mma_tempty mma_afull mma_commit | epi_tfull epi_done  (cycles from start)")
    for k in rows:
        print(k, [int(v - t0) if v else None for v in a[k, :7]])
per = np.median([a[k + 1, 2] - a[k, 2] for k in rows if a[k + 1, 2] and a[k, 2]])
mma = np.median([d(k, 4, 3) for k in rows if d(k, 4, 3)])
epi = np.median([d(k, 6, 5) for k in rows if d(k, 6, 5)])
lat = np.median([d(k, 5, 4) for k in rows if d(k, 5, 4)])
print("  per tile (CTA 0, median of tiles 6..21): period %d clk, MMA issue %d, commit->epilogue start %d, epilogue %d" % (per, mma, lat, epi))
lib.b2u_set_option(b"tc_debug", 0)
