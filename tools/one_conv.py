"""run one tcgen05 conv forward a few times (ncu target) and optionally dump the in-kernel timeline of CTA 0.
usage: one_conv.py N H W CIN COUT HALO_MODE [timeline] [dwmerge]"""
import ctypes as C
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_harness import LIB  # noqa: E402
n, h, w, cin, cout, mode = [int(v) for v in sys.argv[1:7]]
flags = sys.argv[7:]
lib = LIB.lib()
lib.b2u_set_option(b"tc_halo", mode)
lib.b2u_set_option(b"tc_dwmerge", 1 if "dwmerge" in flags else 0)
x = torch.rand(n, h, w, cin, device="cuda").half()
y = torch.empty(n, h, w, cout, device="cuda", dtype=torch.float16)
wt = torch.randn(3, 3, cin, cout, device="cuda") * 0.05
b = torch.zeros(cout, device="cuda")
ws = torch.empty(int(lib.b2u_ws_bytes()), dtype=torch.uint8, device="cuda")


def run():
    LIB.check(lib.b2u_conv3x3_fwd(1, x.data_ptr(), cin, cin, wt.data_ptr(), b.data_ptr(), 1, y.data_ptr(), cout, cout,
                                  None, n, h, w, ws.data_ptr(), ws.numel(), None))


for _ in range(4):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
print("%s: %.4f ms per call (incl. the weight-pack launch)" % (" ".join(sys.argv[1:]), e0.elapsed_time(e1) / 10))
if "timeline" in flags:
    import numpy as np
    lib.b2u_set_option(b"tc_debug", 1)
    run()
    torch.cuda.synchronize()
    buf = (C.c_longlong * 512)()
    lib.b2u_debug_read.argtypes = [C.POINTER(C.c_longlong), C.c_int]
    assert lib.b2u_debug_read(buf, 512) == 0
    a = np.array(buf[:], dtype=np.int64).reshape(64, 8)
    t0 = a[0, 0]
    print("iter: prod_wait prod_issued | mma_tempty mma_afull mma_commit | epi_tfull epi_done epi_first_ld  (cycles from start)")
    for k in range(24):
        print(k, [int(v - t0) if v else None for v in a[k, :8]])
    lib.b2u_set_option(b"tc_debug", 0)
