"""Device-side runtime of the engine: owns the HBM buffers (torch tensors as storage only), binds
plan.py op lists to them and launches them through libb200unet.so -- eagerly or as a CUDA graph.

This is the part of `model.fit / evaluate / predict`
(/root/reference/Scripts/task1_preprocessing_plus_unet_with_comments.py:1059, 1101, 1137) that ran
inside Keras/TensorFlow for the reference.  There is no CPU fallback: everything here needs the CUDA
library and a B200.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from . import plan as P


# --------------------------------------------------------------------------------------------
# initialisers (Keras VarianceScaling semantics: he_normal = truncated normal, glorot_uniform)
# --------------------------------------------------------------------------------------------
def init_weights(graph, seed=42):
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape, init, _tr in graph.weight_specs():
        kind = init[0]
        if kind == "zeros":
            w = np.zeros(shape, np.float32)
        elif kind == "ones":
            w = np.ones(shape, np.float32)
        elif kind == "he_normal":
            std = math.sqrt(2.0 / init[1]) / 0.87962566103423978
            w = rng.standard_normal(shape)
            bad = np.abs(w) > 2.0
            while bad.any():
                w[bad] = rng.standard_normal(int(bad.sum()))
                bad = np.abs(w) > 2.0
            w = (w * std).astype(np.float32)
        elif kind == "glorot_uniform":
            lim = math.sqrt(6.0 / (init[1] + init[2]))
            w = rng.uniform(-lim, lim, size=shape).astype(np.float32)
        else:
            raise ValueError("unknown initializer %r" % (kind,))
        out[name] = w
    return out


class Comm:
    """NCCL communicator for the data-parallel gradient all-reduce (one process per GPU).  The unique id
    is created on rank 0 and broadcast through torch.distributed (any backend)."""

    def __init__(self, rank, world):
        import torch.distributed as dist
        self.rank, self.world = rank, world
        l = _lib.lib()
        idbuf = (C.c_char * 128)()
        if rank == 0:
            _lib.check(l.b2u_comm_unique_id(idbuf), "comm_unique_id")
        t = torch.frombuffer(bytearray(bytes(idbuf)), dtype=torch.uint8).clone()
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.broadcast(t, src=0)
        raw = bytes(t.cpu().numpy().tobytes())
        self.handle = C.c_void_p()
        _lib.check(l.b2u_comm_create(C.c_char_p(raw), rank, world, C.byref(self.handle)), "comm_create")

    def close(self):
        if self.handle:
            _lib.lib().b2u_comm_destroy(self.handle)
            self.handle = None


class _Bound:
    """A plan bound to device addresses, with its ctypes op arrays and (optionally) a captured graph."""

    def __init__(self, plan):
        self.plan = plan
        self.ops = {}          # name -> (ctypes array, n)
        self.graphs = {}       # name -> graph handle
        self.wtab = None       # device copy of plan.pack_table()


class Engine:
    def __init__(self, graph, precision="float16", device=None, seed=42, loss="bce_dice", comm=None,
                 sync_stats=False, use_graph=True, dropout_seed=7, loss_scale=None, plan_options=None):
        if not torch.cuda.is_available():
            raise _lib.B2UError("the b200unet engine needs a CUDA device (no CPU fallback)")
        self.lib = _lib.lib()
        self.graph = graph
        self.dt = P.F16 if precision in ("float16", "fp16", "half", "mixed") else P.F32
        if precision not in ("float16", "fp16", "half", "mixed", "float32", "fp32"):
            raise ValueError("precision must be float16 or float32")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        torch.cuda.set_device(self.device)
        self.stream = torch.cuda.Stream(self.device)
        self.loss = loss
        self.comm = comm
        self.world = comm.world if comm is not None else 1
        self.rank = comm.rank if comm is not None else 0
        self.sync_stats = bool(sync_stats) and self.world > 1
        self.use_graph = use_graph
        self.layout = P.ParamLayout(graph)
        self.user_loss_scale = loss_scale
        self.plan_options = dict(plan_options or {})       # extra plan.Plan keyword arguments (fusion switches)
        npar = self.layout.n_params
        dev = self.device
        with torch.cuda.stream(self.stream):
            self.params = torch.zeros(npar, dtype=torch.float32, device=dev)
            self.state = torch.zeros(max(self.layout.n_state, 1), dtype=torch.float32, device=dev)
            self.grads = torch.zeros(npar, dtype=torch.float32, device=dev)
            self.adam_m = torch.zeros(npar, dtype=torch.float32, device=dev)
            self.adam_v = torch.zeros(npar, dtype=torch.float32, device=dev)
            self.step_dev = torch.zeros(P.STEP_STATE_BYTES, dtype=torch.uint8, device=dev)
            self.ws = torch.empty(int(self.lib.b2u_ws_bytes()), dtype=torch.uint8, device=dev)
        # data parallel: every rank draws its own dropout masks (same counters, rank-specific key)
        self.host_state = _lib.StepState(seed=dropout_seed + 7919 * self.rank, step=0, lr=5e-4, beta1=0.9, beta2=0.999, eps=1e-7,
                                         beta1_pow=0.9, beta2_pow=0.999, loss_scale=1.0,
                                         grad_div=1.0 if self.sync_stats else float(self.world), overflow=0, skip_step=0)
        self._arenas = {}        # name -> torch uint8 tensor (grown on demand, shared between plans)
        self._bound = {}         # (n, training, dropout, loss, tap) -> _Bound
        self._prep_owner = None  # the inference plan whose weight-only ops (plan.prep_ops) are current in the arenas
        self._push_state()
        self.set_weights(init_weights(graph, seed))
        if self.comm is not None:
            # the gradient-bucket all-reduces run on the executor's side stream: create it now (not inside a capture)
            was = self.lib.b2u_set_option(b"comm_overlap", 1)
            if was < 0:
                raise _lib.B2UError("cannot create the communication stream: " + self.lib.b2u_last_error().decode())
            if was == 0:                 # switched off by the caller (A/B runs): keep it off, the stream exists now
                self.lib.b2u_set_option(b"comm_overlap", 0)
            self.broadcast_weights()
            # NCCL connects its transports lazily at the first collective; that must not happen inside a CUDA
            # graph capture, so run one eager all-reduce of the real gradient buffer (zeros) now
            _lib.check(self.lib.b2u_allreduce(self.comm.handle, C.c_void_p(self.grads.data_ptr()), npar, 0,
                                              C.c_void_p(self.stream.cuda_stream)), "allreduce warm-up")
            self.stream.synchronize()

    # ---------------------------------------------------------------------------------------
    # step state
    # ---------------------------------------------------------------------------------------
    def _push_state(self):
        """host -> device copy of the whole step state (blocking; only on (re)configuration)."""
        raw = np.frombuffer(bytes(self.host_state), dtype=np.uint8).copy()
        self.stream.synchronize()
        with torch.cuda.stream(self.stream):
            self.step_dev[:raw.size].copy_(torch.from_numpy(raw))
        self.stream.synchronize()

    def _pull_state(self):
        self.stream.synchronize()
        raw = self.step_dev.cpu().numpy().tobytes()
        C.memmove(C.addressof(self.host_state), raw, C.sizeof(_lib.StepState))
        return self.host_state

    def _set_fields(self, **kw):
        st = self._pull_state()
        changed = False
        for k, v in kw.items():
            if getattr(st, k) != v:
                setattr(st, k, v)
                changed = True
        if changed:
            self._push_state()

    def reset_optimizer(self, lr=5e-4, beta1=0.9, beta2=0.999, eps=1e-7):
        """model.compile(optimizer=Adam(lr)) -- fresh optimizer state, weights kept (CV4:1062-1083)."""
        self.stream.synchronize()
        with torch.cuda.stream(self.stream):
            self.adam_m.zero_()
            self.adam_v.zero_()
        st = self._pull_state()
        st.lr, st.beta1, st.beta2, st.eps = lr, beta1, beta2, eps
        st.beta1_pow, st.beta2_pow, st.overflow = beta1, beta2, 0
        self._push_state()

    @property
    def lr(self):
        return float(self._pull_state().lr)

    @lr.setter
    def lr(self, v):
        self._set_fields(lr=float(np.float32(v)))

    def guard_small_gamma(self, threshold=1e-2):
        """The default BatchNorm-backward fusions recover sum(dy * xhat) from the stored BN OUTPUT as (y - beta) / gamma
        (plan.py: statistics from the max-pool backward / from <W, dW>): exact for |gamma| of order one, but the
        rounding of y is amplified by 1 / gamma and a channel with gamma == 0 loses its d(gamma) altogether.
        Keras initialises gamma = 1 and training rarely drives it to zero; when any |gamma| falls below `threshold` the
        engine switches -- once, with a warning -- to the unfused reductions over dy and the BN input, which do not
        divide by gamma.  Called by Model.fit after every epoch and by set_weights.  Returns True if it switched."""
        if not self.plan_options.get("fuse_bn_bwd", True) and not self.plan_options.get("fuse_bn_bwd_wgrad", True):
            return False
        self.stream.synchronize()
        p = self.params.cpu().numpy()
        lo = min((float(np.abs(p[off:off + int(np.prod(shape))]).min()) for name, (arena, off, shape) in self.layout.offsets.items()
                  if name.endswith("/gamma")), default=1.0)
        if lo >= threshold:
            return False
        import warnings
        warnings.warn("b200unet: min |BatchNorm gamma| = %.3g < %g: switching the BatchNorm backward statistics to the "
                      "unfused reductions (no division by gamma)" % (lo, threshold))
        self.plan_options.update(fuse_bn_bwd=False, fuse_bn_bwd_wgrad=False)
        for b in self._bound.values():
            for gph in b.graphs.values():
                self.lib.b2u_graph_destroy(gph)
        self._bound.clear()
        return True

    def overflowed(self):
        """True once a step was skipped because its gradient was not finite (the device then halved the loss scale)"""
        return bool(self._pull_state().overflow)

    @property
    def loss_scale(self):
        """the loss scale on the device: the static choice of train_batch, halved by every skipped step"""
        return float(self._pull_state().loss_scale)

    # ---------------------------------------------------------------------------------------
    # weights
    # ---------------------------------------------------------------------------------------
    def set_weights(self, weights):
        fp, fs = self.layout.pack(weights)
        self.stream.synchronize()
        with torch.cuda.stream(self.stream):
            self.params.copy_(torch.from_numpy(fp))
            self.state[:fs.size].copy_(torch.from_numpy(fs))
        self.stream.synchronize()
        self._prep_owner = None
        self.guard_small_gamma()

    def get_weights(self):
        self.stream.synchronize()
        return self.layout.unpack(self.params.cpu().numpy(), self.state.cpu().numpy())

    def get_grads(self):
        self.stream.synchronize()
        return self.layout.unpack(self.grads.cpu().numpy(), None)

    def broadcast_weights(self):
        """make every rank start from rank 0's weights"""
        import torch.distributed as dist
        for t in (self.params, self.state):
            if dist.get_backend() == "nccl":
                dist.broadcast(t, src=0)
            else:
                h = t.cpu()
                dist.broadcast(h, src=0)
                t.copy_(h)
        torch.cuda.synchronize(self.device)
        self._prep_owner = None

    def average_moving_statistics(self):
        """data parallel without sync_stats: mean over the ranks of the BatchNorm moving means / variances (the `state`
        buffer); every rank then validates and checkpoints with the same statistics"""
        if self.world == 1:
            return
        import torch.distributed as dist
        self.stream.synchronize()
        if dist.get_backend() == "nccl":
            dist.all_reduce(self.state)
            self.state /= self.world
        else:
            h = self.state.cpu()
            dist.all_reduce(h)
            self.state.copy_(h / self.world)
        torch.cuda.synchronize(self.device)
        self._prep_owner = None

    # ---------------------------------------------------------------------------------------
    # plans
    # ---------------------------------------------------------------------------------------
    def _arena(self, name, nbytes):
        t = self._arenas.get(name)
        if t is None or t.numel() < nbytes:
            self.stream.synchronize()
            for b in self._bound.values():        # addresses change: drop bound op arrays / graphs
                for g in b.graphs.values():
                    self.lib.b2u_graph_destroy(g)
                b.ops.clear()
                b.graphs.clear()
            with torch.cuda.stream(self.stream):
                t = torch.zeros(int(nbytes) + 256, dtype=torch.uint8, device=self.device)
            self._arenas[name] = t
            self._prep_owner = None
        return t

    def _get_bound(self, n, training, dropout=True, loss=None, tap=False):
        """tap=True: a plan whose every layer output is materialised as the reference's layer would produce it
        (`layer_output`); the default inference plan folds BatchNormalization into the producing conv."""
        key = (int(n), bool(training), bool(dropout), loss or self.loss, bool(tap))
        b = self._bound.get(key)
        if b is None:
            opts = dict(self.plan_options)
            if tap:
                opts["fuse_bn_infer"] = False
            pl = P.Plan(self.graph, n, dt=self.dt, training=training, dropout=dropout, loss=loss or self.loss,
                        world=self.world, sync_stats=self.sync_stats, layout=self.layout, rank=self.rank,
                        **opts)
            b = _Bound(pl)
            self._bound[key] = b
        sizes = b.plan.arena_sizes()
        for name in ("act", "f32", "zero", "wpack"):
            self._arena(name, sizes[name])
        if b.wtab is None:              # the plan's weight-packing table (OP_PACK_WEIGHTS), uploaded once
            with torch.cuda.stream(self.stream):
                b.wtab = torch.from_numpy(b.plan.pack_table().reshape(-1).copy()).to(self.device)
            self.stream.synchronize()
        return b

    def _resolve(self, ref):
        a = ref.arena
        if a in self._arenas:
            base = self._arenas[a].data_ptr()
        elif a == "params":
            base = self.params.data_ptr()
        elif a == "state":
            base = self.state.data_ptr()
        elif a == "grads":
            base = self.grads.data_ptr()
        elif a == "adam_m":
            base = self.adam_m.data_ptr()
        elif a == "adam_v":
            base = self.adam_v.data_ptr()
        elif a == "step":
            base = self.step_dev.data_ptr()
        else:
            raise KeyError(a)
        return base + ref.off

    def _ops(self, b, name):
        if name not in b.ops:
            pl = b.plan
            lst = {"train": pl.train_ops, "forward": lambda: pl.forward_ops(prep=False),
                   "forward_loss": lambda: pl.forward_ops(with_loss=True, prep=False), "prep": pl.prep_ops}[name]()
            resolve = lambda ref: (b.wtab.data_ptr() + ref.off) if ref.arena == "wtab" else self._resolve(ref)
            b.ops[name] = (_lib.make_ops(lst, resolve), len(lst))
        return b.ops[name]

    def _run(self, b, name):
        comm = self.comm.handle if self.comm is not None else None
        s = C.c_void_p(self.stream.cuda_stream)
        # inference plans keep their weight-only ops (operand packing, BN scale / shift) out of the per-batch list: they
        # run here, eagerly, when the weights have changed or another plan has used the shared arenas since
        if name != "train" and b.plan.hoist_prep and self._prep_owner is not b:
            parr, pn = self._ops(b, "prep")
            if pn:
                _lib.check(self.lib.b2u_run_ops(parr, pn, C.c_void_p(self.ws.data_ptr()), self.ws.numel(), comm, s),
                           "run_ops(prep)")
        self._prep_owner = b if name != "train" else None
        arr, n = self._ops(b, name)
        if self.use_graph:
            g = b.graphs.get(name)
            if g is None:
                h = C.c_void_p()
                _lib.check(self.lib.b2u_graph_create(arr, n, C.c_void_p(self.ws.data_ptr()), self.ws.numel(), comm, s,
                                                     C.byref(h)), "graph_create(%s)" % name)
                b.graphs[name] = g = h
            _lib.check(self.lib.b2u_graph_launch(g, s), "graph_launch(%s)" % name)
        else:
            _lib.check(self.lib.b2u_run_ops(arr, n, C.c_void_p(self.ws.data_ptr()), self.ws.numel(), comm, s),
                       "run_ops(%s)" % name)

    def _gather(self, dt, src, idx, dst_ptr, per_sample, nb):
        _lib.check(self.lib.b2u_gather_batch(dt, C.c_void_p(src.data_ptr()),
                                             C.c_void_p(idx.data_ptr()) if idx is not None else None,
                                             C.c_void_p(dst_ptr), per_sample, nb, C.c_void_p(self.stream.cuda_stream)),
                   "gather_batch")

    def _load_inputs(self, b, x_src, idx, n, t_src=None, sw_src=None):
        pl = b.plan
        xv = pl.x_view
        assert x_src.dtype == torch.float32 and x_src.is_contiguous()
        if pl.x_pad:                     # channels zero-padded for the tensor-core first conv (plan.py)
            _lib.check(self.lib.b2u_gather_batch_pad(self.dt, C.c_void_p(x_src.data_ptr()),
                                                     C.c_void_p(idx.data_ptr()) if idx is not None else None,
                                                     C.c_void_p(self._resolve(xv.ref)), xv.h * xv.w, pl.x_cin, pl.x_pad, n,
                                                     C.c_void_p(self.stream.cuda_stream)), "gather_batch_pad")
        else:
            self._gather(self.dt, x_src, idx, self._resolve(xv.ref), xv.h * xv.w * xv.c, n)
        if t_src is not None:
            per_t = int(math.prod(pl.prob_shape[1:]))
            assert t_src.dtype == torch.float32 and t_src.is_contiguous()
            self._gather(P.F32, t_src, idx, self._resolve(pl.target), per_t, n)
            if pl.loss == "bce":
                if sw_src is None:
                    raise ValueError("the weighted-BCE plan needs per-sample weights")
                self._gather(P.F32, sw_src, idx, self._resolve(pl.sample_w), 1, n)

    def _loss_scale_for(self, pl):
        if self.user_loss_scale is not None:
            return float(self.user_loss_scale)
        if self.dt == P.F32:
            return 1.0
        nel = int(math.prod(pl.prob_shape)) * (self.world if self.sync_stats else 1)
        s = 2.0 ** math.floor(math.log2(max(nel / 64.0, 1.0)))
        return float(min(max(s, 1.0), 32768.0))

    # ---------------------------------------------------------------------------------------
    # public device API (all asynchronous on self.stream unless they return host data)
    # ---------------------------------------------------------------------------------------
    def train_batch(self, x_src, t_src, idx, n, dropout=True, sw_src=None):
        """one optimisation step on samples x_src[idx] (device-resident fp32, NHWC)."""
        b = self._get_bound(n, True, dropout)
        ls = self._loss_scale_for(b.plan)
        if getattr(self, "_cur_ls", None) != ls:
            self._set_fields(loss_scale=ls)
            self._cur_ls = ls
        self._load_inputs(b, x_src, idx, n, t_src, sw_src)
        self._run(b, "train")
        return b

    def forward_batch(self, x_src, idx, n, t_src=None, sw_src=None, training=False, tap=False):
        """inference-mode forward (BN moving statistics, no dropout); with t_src also the loss."""
        b = self._get_bound(n, training, False, tap=tap)
        self._load_inputs(b, x_src, idx, n, t_src, sw_src)
        self._run(b, "forward_loss" if t_src is not None else "forward")
        return b

    def probs(self, b):
        """device view of the last forward's sigmoid outputs, shape plan.prob_shape (fp32)."""
        pl = b.plan
        nel = int(math.prod(pl.prob_shape))
        a = self._arenas["f32"]
        return a[pl.prob.off:pl.prob.off + nel * 4].view(torch.float32).view(*pl.prob_shape)

    def targets(self, b):
        pl = b.plan
        nel = int(math.prod(pl.prob_shape))
        a = self._arenas["f32"]
        return a[pl.target.off:pl.target.off + nel * 4].view(torch.float32).view(*pl.prob_shape)

    def loss_dev(self, b):
        """device view: [loss, metric] of the last step (fp32)."""
        a = self._arenas["f32"]
        off = b.plan.loss_out.off
        return a[off:off + 8].view(torch.float32)

    def layer_output(self, b, name):
        """host copy of an intermediate activation (Model(inputs, get_layer(name).output), T1H:1386)."""
        v = b.plan.layer_out[name]
        if name in b.plan.folded_into_next:
            raise ValueError("layer %r is fused with the BatchNormalization behind it in this plan: run forward_batch(..., "
                             "tap=True) to read its own output" % name)
        self.stream.synchronize()

        def read(v):
            a = self._arenas[v.ref.arena]
            es = P.ELEM[v.dt]
            npix = b.plan.n * v.h * v.w
            tdt = torch.float16 if v.dt == P.F16 else torch.float32
            raw = a[v.ref.off:v.ref.off + ((npix - 1) * v.ld + v.c) * es].view(tdt)
            out = torch.as_strided(raw, (npix, v.c), (v.ld, 1)).float().cpu().numpy()
            return out.reshape(b.plan.n, v.h, v.w, v.c)

        if isinstance(v, P.SplitView):          # a concatenate kept as two dense tensors
            return np.concatenate([read(pv) for pv in v.parts], axis=-1)
        return read(v)

    def threshold_counts(self, b, thresholds_dev, tp, spr, sgt):
        pl = b.plan
        nel = int(math.prod(pl.prob_shape))
        _lib.check(self.lib.b2u_threshold_counts(C.c_void_p(self._resolve(pl.prob)), C.c_void_p(self._resolve(pl.target)),
                                                 nel, C.c_void_p(thresholds_dev.data_ptr()), thresholds_dev.numel(),
                                                 C.c_void_p(tp.data_ptr()), C.c_void_p(spr.data_ptr()),
                                                 C.c_void_p(sgt.data_ptr()), C.c_void_p(self.stream.cuda_stream)),
                   "threshold_counts")

    def profile_train_ops(self, n, dropout=True):
        """One eager training step with a CUDA event between ops (b2u_run_ops_timed): returns
        [(plan.Op, milliseconds)].  Inputs must already be in place (call train_batch once before)."""
        b = self._get_bound(n, True, dropout)
        arr, cnt = self._ops(b, "train")
        self._prep_owner = None
        ms = (C.c_float * cnt)()
        comm = self.comm.handle if self.comm is not None else None
        _lib.check(self.lib.b2u_run_ops_timed(arr, cnt, C.c_void_p(self.ws.data_ptr()), self.ws.numel(), comm,
                                              C.c_void_p(self.stream.cuda_stream), ms), "run_ops_timed")
        return list(zip(b.plan.train_ops(), [float(v) for v in ms]))

    def close(self):
        self.stream.synchronize()
        for b in self._bound.values():
            for g in b.graphs.values():
                self.lib.b2u_graph_destroy(g)
            b.graphs.clear()
