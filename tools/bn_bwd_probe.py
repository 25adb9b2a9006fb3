"""GPU probe (not a test): BN_BWD_APPLY variants at the 512^2 x 32-channel level (batch 8): plain, + activation mask read from
the BN input itself, + column sums, + stored dropout bits."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gpu_harness import LIB, P  # noqa: E402

lib = LIB.lib()
for f in sys.argv[1:]:                            # opt:name=value -> b2u_set_option
    if f.startswith("opt:"):
        lib.b2u_set_option(f[4:].split("=")[0].encode(), int(f.split("=")[1]))
n, h, w, c = 8, 512, 512, int(os.environ.get("PROBE_C", "32"))
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


dy = torch.randn(npix, c, device="cuda").half() * 0.1
x = (torch.rand(npix, c, device="cuda") - 0.3).half()
x2 = x.clone()
dx = torch.empty(npix, c, device="cuda", dtype=torch.float16)
gamma = torch.ones(c, device="cuda"); mean = torch.zeros(c, device="cuda"); invstd = torch.ones(c, device="cuda")
sums = torch.zeros(2 * c, device="cuda", dtype=torch.float64)
dg = torch.zeros(c, device="cuda"); db = torch.zeros(c, device="cuda"); cs = torch.zeros(c, device="cuda")
bits = torch.randint(0, 255, (npix * c // 8,), device="cuda", dtype=torch.uint8)
base_p = [R(dy), R(x), R(dx), R(gamma), R(mean), R(invstd), R(sums), R(dg), R(db)]
base_i = [c, c, c, c, npix]
for name, p, i, f in (
        ("plain", base_p + [None, None], base_i + [0, 0, npix], []),
        ("mask = the BN input itself", base_p + [R(x), None], base_i + [c, 1, npix], []),
        ("mask = another tensor", base_p + [R(x2), None], base_i + [c, 1, npix], []),
        ("mask (input) + column sums", base_p + [R(x), R(cs)], base_i + [c, 1, npix], []),
        ("column sums only", base_p + [None, R(cs)], base_i + [0, 0, npix], []),
        ("mask (input) + column sums + dropout bits", base_p + [R(x), R(cs), None, None, R(bits)], base_i + [c, 2, npix], [0.5]),
        ("dropout bits only", base_p + [None, None, None, None, R(bits)], base_i + [0, 0, npix], [0.5])):
    print("%-45s %.4f ms" % (name, timed(P.Op(P.OP_BN_BWD_APPLY, dt, p, i, f))))
