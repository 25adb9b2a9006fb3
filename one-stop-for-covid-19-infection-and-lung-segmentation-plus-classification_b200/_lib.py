"""ctypes binding of libb200unet.so (the C-ABI declared in include/b200unet.h).

The product path has no CPU fallback: if the shared library is missing or a call fails, this module
raises.  Build it with `python __graft_entry__.py` (or `<package>/build.sh`).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200unet.so")


class B2UError(RuntimeError):
    pass


class StepState(C.Structure):
    """mirror of b2u_step_state (include/b200unet.h)"""
    _fields_ = [("seed", C.c_uint64), ("step", C.c_uint64), ("lr", C.c_float), ("beta1", C.c_float),
                ("beta2", C.c_float), ("eps", C.c_float), ("beta1_pow", C.c_float), ("beta2_pow", C.c_float),
                ("loss_scale", C.c_float), ("grad_div", C.c_float), ("overflow", C.c_uint32), ("skip_step", C.c_uint32)]


class Op(C.Structure):
    """mirror of b2u_op"""
    _fields_ = [("kind", C.c_int32), ("dt", C.c_int32), ("p", C.c_void_p * 14), ("i", C.c_int64 * 12),
                ("f", C.c_float * 4)]


_lib = None

# every symbol include/b200unet.h declares (tests/test_abi.py checks the .so exports all of them)
SYMBOLS = [
    "b2u_version", "b2u_last_error", "b2u_ws_bytes", "b2u_tensor_path_available", "b2u_set_option", "b2u_state_advance",
    "b2u_conv3x3_fwd", "b2u_conv3x3_dgrad", "b2u_conv3x3_wgrad", "b2u_convt2x2_fwd", "b2u_convt2x2_dgrad",
    "b2u_convt2x2_wgrad", "b2u_bn_stats", "b2u_bn_finalize", "b2u_bn_apply", "b2u_bn_bwd_reduce",
    "b2u_bn_bwd_apply", "b2u_maxpool_fwd", "b2u_maxpool_bwd", "b2u_dropout_fwd", "b2u_dropout_bwd",
    "b2u_copy_slice", "b2u_head_fwd", "b2u_bce_dice_sums", "b2u_bce_dice_finalize", "b2u_head_bwd",
    "b2u_dense_fwd", "b2u_dense_bwd", "b2u_bce_fwd", "b2u_bce_sigmoid_bwd", "b2u_adam", "b2u_gather_batch", "b2u_gather_batch_pad",
    "b2u_threshold_counts", "b2u_clahe_u8", "b2u_crop_resize", "b2u_resize_u8", "b2u_resize_area_f64", "b2u_run_ops", "b2u_run_ops_timed", "b2u_graph_create",
    "b2u_graph_launch", "b2u_graph_destroy", "b2u_launch_count", "b2u_comm_unique_id", "b2u_comm_create",
    "b2u_comm_destroy", "b2u_allreduce", "b2u_pack_weights", "b2u_bn_bwd_sums_from_wgrad", "b2u_bn_apply_pool",
]


def lib():
    """Load (once) and return the shared library; raises B2UError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B2UError("libb200unet.so not found at %s -- build it first (python -c 'import __graft_entry__ as g; "
                       "g.build()'); there is no CPU fallback" % LIB_PATH)
    l = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32, i64, f32, sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t
    l.b2u_version.restype = C.c_int
    l.b2u_last_error.restype = C.c_char_p
    l.b2u_ws_bytes.restype = sz
    l.b2u_launch_count.restype = i64
    l.b2u_tensor_path_available.restype = C.c_int
    l.b2u_set_option.argtypes = [C.c_char_p, C.c_int]
    l.b2u_run_ops.argtypes = [C.POINTER(Op), i32, vp, sz, vp, vp]
    l.b2u_run_ops_timed.argtypes = [C.POINTER(Op), i32, vp, sz, vp, vp, C.POINTER(C.c_float)]
    l.b2u_graph_create.argtypes = [C.POINTER(Op), i32, vp, sz, vp, vp, C.POINTER(vp)]
    l.b2u_graph_launch.argtypes = [vp, vp]
    l.b2u_graph_destroy.argtypes = [vp]
    l.b2u_gather_batch.argtypes = [i32, vp, vp, vp, i64, i32, vp]
    l.b2u_gather_batch_pad.argtypes = [i32, vp, vp, vp, i64, i32, i32, i32, vp]
    l.b2u_threshold_counts.argtypes = [vp, vp, i64, vp, i32, vp, vp, vp, vp]
    l.b2u_comm_unique_id.argtypes = [vp]
    l.b2u_comm_create.argtypes = [vp, i32, i32, C.POINTER(vp)]
    l.b2u_comm_destroy.argtypes = [vp]
    l.b2u_allreduce.argtypes = [vp, vp, i64, i32, vp]
    l.b2u_clahe_u8.argtypes = [vp, vp, i32, i32, i32, f32, i32, vp, sz, vp]
    l.b2u_crop_resize.argtypes = [vp, i32, i32, i32, vp, i32, i32, i32, vp, vp, vp]
    l.b2u_resize_u8.argtypes = [vp, i32, i32, i32, vp, i32, i32, i32, vp]
    l.b2u_resize_area_f64.argtypes = [vp, i32, i32, i32, vp, i32, i32, i32, vp]
    # per-op entry points (used directly by the per-kernel parity tests)
    l.b2u_conv3x3_fwd.argtypes = [i32, vp, i32, i32, vp, vp, i32, vp, i32, i32, vp, i32, i32, i32, vp, sz, vp]
    l.b2u_conv3x3_dgrad.argtypes = [i32, vp, i32, i32, vp, vp, i32, i32, vp, i32, i32, i32, i32, i32, i32, vp, sz, vp]
    l.b2u_conv3x3_wgrad.argtypes = [i32, vp, i32, i32, vp, i32, i32, vp, vp, i32, i32, i32, vp, sz, vp]
    l.b2u_convt2x2_fwd.argtypes = [i32, vp, i32, i32, vp, vp, vp, i32, i32, vp, i32, i32, i32, i32, vp, sz, vp]
    l.b2u_convt2x2_dgrad.argtypes = [i32, vp, i32, i32, vp, vp, i32, i32, vp, i32, i32, i32, i32, i32, i32, vp, sz, vp]
    l.b2u_convt2x2_wgrad.argtypes = [i32, vp, i32, i32, vp, i32, i32, vp, vp, i32, i32, i32, vp, sz, vp]
    l.b2u_adam.argtypes = [vp, vp, vp, vp, i64, vp, vp]
    l.b2u_pack_weights.argtypes = [vp, i32, vp, vp, i64, vp]
    l.b2u_state_advance.argtypes = [vp, vp]
    _lib = l
    return l


def check(rc, what=""):
    if rc != 0:
        msg = lib().b2u_last_error()
        raise B2UError("%s failed (rc=%d): %s" % (what or "libb200unet call", rc, msg.decode() if msg else "?"))


def make_ops(op_list, resolve):
    """plan.Op list -> ctypes array of b2u_op; `resolve(Ref) -> int` gives absolute device addresses."""
    arr = (Op * len(op_list))()
    for k, o in enumerate(op_list):
        r = arr[k]
        r.kind, r.dt = o.kind, o.dt
        for j, p in enumerate(o.p):
            r.p[j] = resolve(p) if p is not None else None
        for j, v in enumerate(o.i):
            r.i[j] = v
        for j, v in enumerate(o.f):
            r.f[j] = v
    return arr
