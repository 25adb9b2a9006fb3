"""Runs a plan.Op list through libb200unet.so on the GPU over caller-provided arenas (the same
byte-for-byte memory image the CPU emulator interprets), so kernels are compared op by op."""
import ctypes as C
import importlib

import numpy as np
import torch

from conftest import PKG

LIB = importlib.import_module(PKG + "._lib")
P = importlib.import_module(PKG + ".plan")


def run_ops_gpu(ops, mem, state, graph=False):
    """mem: dict arena -> np.uint8 array (copied); state: dict like Emulator.state. Returns new mem dict."""
    l = LIB.lib()
    dev = {k: torch.from_numpy(v.copy()).cuda() for k, v in mem.items()}
    st = LIB.StepState(seed=state["seed"], step=state["step"], lr=state["lr"], beta1=state["beta1"], beta2=state["beta2"],
                       eps=state["eps"], beta1_pow=state["beta1_pow"], beta2_pow=state["beta2_pow"],
                       loss_scale=state["loss_scale"], grad_div=state["grad_div"], overflow=0, skip_step=0)
    dev["step"] = torch.from_numpy(np.frombuffer(bytes(st), dtype=np.uint8).copy()).cuda()
    ws = torch.empty(int(l.b2u_ws_bytes()), dtype=torch.uint8, device="cuda")
    arr = LIB.make_ops(ops, lambda r: dev[r.arena].data_ptr() + r.off)
    stream = torch.cuda.Stream()
    torch.cuda.synchronize()
    if graph:
        h = C.c_void_p()
        LIB.check(l.b2u_graph_create(arr, len(ops), C.c_void_p(ws.data_ptr()), ws.numel(), None,
                                     C.c_void_p(stream.cuda_stream), C.byref(h)), "graph_create")
        LIB.check(l.b2u_graph_launch(h, C.c_void_p(stream.cuda_stream)), "graph_launch")
        stream.synchronize()
        l.b2u_graph_destroy(h)
    else:
        LIB.check(l.b2u_run_ops(arr, len(ops), C.c_void_p(ws.data_ptr()), ws.numel(), None,
                                C.c_void_p(stream.cuda_stream)), "run_ops")
        stream.synchronize()
    torch.cuda.synchronize()
    out = {k: v.cpu().numpy() for k, v in dev.items()}
    raw = out.pop("step").tobytes()
    st2 = LIB.StepState()
    C.memmove(C.addressof(st2), raw, C.sizeof(LIB.StepState))
    return out, st2
