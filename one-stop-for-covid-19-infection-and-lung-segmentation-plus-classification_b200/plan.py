"""Lowers a layers.Graph to the flat device op lists libb200unet.so executes (b2u_run_ops).

This is the host-side "compiler" of the engine: it decides the HBM layout (one activation arena with
zero-copy concat buffers, flat fp32 parameter / gradient / Adam buffers in Keras weight order), the
fusions (BN statistics in the conv epilogue, max-pool + dropout in one pass, activation-derivative
masks folded into whichever backward kernel produces a gradient) and the backward schedule.  It is pure
Python with no CUDA dependency, so the schedule is testable on a CPU box (tests/emulator.py interprets
the same op list with numpy).

Replaces what `model.compile(...)` + the Keras/TF graph builder did for the reference
(/root/reference/Scripts/task1_preprocessing_plus_unet_with_comments.py:1053-1061).
"""
from collections import OrderedDict
import math

from . import layers as L

# ---- enums mirrored from include/b200unet.h ------------------------------------------------------
F32, F16 = 0, 1
ACT = {None: 0, "linear": 0, "relu": 1, "elu": 2, "sigmoid": 3}
ACT_RELU_BITS = 4       # mask_act of a backward op whose mask is a packed 1-bit ReLU mask (B2U_ACT_RELU_BITS)
(OP_CONV3X3_FWD, OP_CONV3X3_DGRAD, OP_CONV3X3_WGRAD, OP_CONVT_FWD, OP_CONVT_DGRAD, OP_CONVT_WGRAD,
 OP_BN_STATS, OP_BN_FINALIZE, OP_BN_APPLY, OP_BN_BWD_REDUCE, OP_BN_BWD_APPLY, OP_MAXPOOL_FWD, OP_MAXPOOL_BWD,
 OP_DROPOUT_FWD, OP_DROPOUT_BWD, OP_COPY_SLICE, OP_HEAD_FWD, OP_BCE_DICE_SUMS, OP_BCE_DICE_FINALIZE, OP_HEAD_BWD,
 OP_DENSE_FWD, OP_DENSE_BWD, OP_BCE_FWD, OP_BCE_SIGMOID_BWD, OP_ADAM, OP_MEMSET, OP_ALLREDUCE_F32,
 OP_ALLREDUCE_F64, OP_STATE_ADVANCE, OP_GATHER_BATCH, OP_PACK_WEIGHTS, OP_BN_BWD_SUMS_WGRAD, OP_BN_APPLY_POOL) = range(1, 34)
OP_NAMES = {v: k for k, v in list(globals().items()) if k.startswith("OP_")}

OPF_SIDE, OPF_JOIN, OPF_COMM = 0x100, 0x200, 0x400    # executor flags OR-ed into Op.dt (include/b200unet.h B2U_OPF_*)
ELEM = {F32: 4, F16: 2}
STEP_STATE_BYTES = 64


class Ref:
    """A location inside a named arena (resolved to a device pointer at bind time)."""
    __slots__ = ("arena", "off")

    def __init__(self, arena, off):
        self.arena, self.off = arena, int(off)

    def __add__(self, nbytes):
        return Ref(self.arena, self.off + int(nbytes))

    def __repr__(self):
        return "%s+%d" % (self.arena, self.off)


class View:
    """NHWC activation view: base ref, ld (elements between pixels), c channels, dt storage type.
    Flat (n,k) tensors use h = w = 1, c = ld = k."""
    __slots__ = ("ref", "ld", "c", "h", "w", "dt")

    def __init__(self, ref, ld, c, h, w, dt):
        self.ref, self.ld, self.c, self.h, self.w, self.dt = ref, int(ld), int(c), int(h), int(w), dt

    def slice(self, c0, c):
        return View(self.ref + c0 * ELEM[self.dt], self.ld, c, self.h, self.w, self.dt)


class SplitView:
    """A two-input concatenate kept as two dense tensors (`parts`): what reads it -- the BatchNormalization behind it --
    takes both sources, so no kernel ever touches one half of an interleaved 2c-channel pixel."""
    __slots__ = ("parts",)

    def __init__(self, a, b):
        self.parts = (a, b)

    c = property(lambda self: self.parts[0].c + self.parts[1].c)
    h = property(lambda self: self.parts[0].h)
    w = property(lambda self: self.parts[0].w)
    dt = property(lambda self: self.parts[0].dt)


class Op:
    __slots__ = ("kind", "dt", "p", "i", "f", "tag")

    def __init__(self, kind, dt, p=(), i=(), f=(), tag=""):
        self.kind, self.dt, self.p, self.i, self.f, self.tag = kind, dt, list(p), [int(v) for v in i], \
            [float(v) for v in f], tag
        assert len(self.p) <= 14 and len(self.i) <= 12 and len(self.f) <= 4

    def __repr__(self):
        return "%s[%s] p=%s i=%s f=%s" % (OP_NAMES[self.kind], self.tag, self.p, self.i, self.f)


class Arena:
    def __init__(self, name):
        self.name, self.size = name, 0

    def alloc(self, nbytes, align=256):
        off = (self.size + align - 1) // align * align
        self.size = off + int(nbytes)
        return Ref(self.name, off)


class ParamLayout:
    """Flat fp32 buffers in Keras weight order: `params` (trainable: kernels, biases, gamma, beta) and
    `state` (BatchNormalization moving statistics).  grads / adam_m / adam_v mirror `params`."""

    def __init__(self, graph):
        self.specs = graph.weight_specs()
        self.offsets = OrderedDict()     # name -> (arena, element offset, shape)
        np_, ns = 0, 0
        for name, shape, init, trainable in self.specs:
            n = int(math.prod(shape))
            if trainable:
                self.offsets[name] = ("params", np_, shape)
                np_ += (n + 3) // 4 * 4          # keep every tensor 16-byte aligned
            else:
                self.offsets[name] = ("state", ns, shape)
                ns += (n + 3) // 4 * 4
        self.n_params, self.n_state = np_, ns

    def ref(self, name, arena=None):
        a, off, _ = self.offsets[name]
        return Ref(arena or a, off * 4)

    def pack(self, weights):
        """name -> ndarray dict (Keras layouts)  ->  (flat params fp32, flat state fp32)."""
        import numpy as np
        flat = {"params": np.zeros(self.n_params, np.float32), "state": np.zeros(max(self.n_state, 1), np.float32)}
        for name, (arena, off, shape) in self.offsets.items():
            a = np.asarray(weights[name], np.float32)
            if tuple(a.shape) != tuple(shape):
                raise ValueError("weight %s: expected shape %s, got %s" % (name, shape, a.shape))
            flat[arena][off:off + a.size] = a.reshape(-1)
        return flat["params"], flat["state"]

    def unpack(self, params, state, arena_for_trainable="params"):
        """inverse of pack(); `params` may also be a gradient / Adam buffer with the same layout."""
        import numpy as np
        out = OrderedDict()
        for name, (arena, off, shape) in self.offsets.items():
            n = int(math.prod(shape))
            src = params if arena == "params" else state
            if src is None:
                continue
            out[name] = np.array(src[off:off + n], np.float32).reshape(shape)
        return out


class Plan:
    """Op lists + arena sizes for one (graph, batch size, storage type, mode) combination."""

    def __init__(self, graph, n, dt=F32, training=True, dropout=True, loss="bce_dice", world=1,
                 sync_stats=False, layout=None, rank=0, fuse_bn_bwd=True, fuse_bn_stats=True, fuse_bias_grad=True,
                 prepack=True, fuse_bn_bwd_wgrad=True, fuse_bn_pool=True, relu_bits=None, grad_bucket_bytes=6 << 20,
                 fuse_bn_infer=True, hoist_prep=True, split_concat=True, fuse_dropout_bn=True):
        self.graph, self.n, self.dt, self.training = graph, int(n), dt, training
        self.dropout = dropout and training
        self.loss, self.world, self.sync_stats = loss, int(world), bool(sync_stats) and world > 1
        self.rank = int(rank)
        self.fuse_bn_bwd = bool(fuse_bn_bwd)
        self.fuse_bn_stats = bool(fuse_bn_stats)
        self.fuse_bias_grad = bool(fuse_bias_grad)
        self.fuse_bn_bwd_wgrad = bool(fuse_bn_bwd_wgrad)
        self.fuse_bn_pool = bool(fuse_bn_pool)
        # inference plans (fp16 storage): Conv2D(relu) -> BatchNormalization (T2:749-750, 757-758, 766-767) runs as ONE kernel,
        # the BN of the moving statistics being a per-channel affine after the activation in the conv epilogue; and the
        # ops that depend on the weights only (operand packing, scale / shift of every BN) leave the per-batch op list:
        # `prep_ops()` runs when the weights have changed (engine.py), `forward_ops(prep=False)` per batch.
        # Classifier 224 x 224 x 3, batch 64 on B200: 0.446 -> see profiles/NOTES_r2.md.
        # concatenate([up-sampled, skip]) -> BatchNormalization -> Conv2D (every U-Net decoder level, T1H:887-889): the two
        # inputs stay dense tensors of their own and the BN kernels read / write both.  In one interleaved buffer each input
        # is one 64-byte half of a 128-byte pixel, and every kernel that touches only one input (max-pool backward, the
        # transposed conv's three passes, the encoder BN backward) fetches twice its bytes from DRAM (measured on B200 at
        # 512^2 x 32 channels, tools/half_line_probe.py: max-pool backward 0.125 -> 0.101 ms, transposed-conv weight
        # gradient 0.103 -> 0.072 ms, data gradient 0.069 -> 0.054 ms, BN backward 0.083 -> 0.066 ms; in the step:
        # max-pool backward 0.125 -> 0.092, transposed-conv weight gradient 0.064 -> 0.049, step -0.05 ms).
        # Conv2D -> Dropout -> BatchNormalization (the U-Net++ decoder blocks, UPP:874-876 ...): the Dropout output is never
        # written; the BN statistics pass runs Philox once and stores the keep mask as packed bits, the apply / backward reduce /
        # backward apply kernels read one byte per 8 elements and the conv output instead -- one full read + write pass less in
        # each direction per block (regenerating the mask in all four passes cost what the removed passes had cost).
        self.fuse_dropout_bn = bool(fuse_dropout_bn)
        self._lazy_drop = {}             # id(dropout output tensor) -> (rate, dropout op index)
        self.split_concat = int(split_concat)        # 0: never, 1: inputs narrower than a 128-byte line, 2: every eligible concat
        self.fuse_bn_infer = bool(fuse_bn_infer) and not training and dt == F16
        self.hoist_prep = bool(hoist_prep) and not training
        self.prep = []                   # inference plans: weight-only ops (BN finalize from the moving statistics)
        # data parallel: the flat gradient buffer is exchanged in buckets of about this many bytes, each all-reduced as soon
        # as the last backward op that writes into it has been issued (0 = one all-reduce after the whole backward)
        self.grad_bucket_bytes = int(grad_bucket_bytes)
        # 1-bit ReLU masks: a ReLU conv whose only consumer is another 3x3 conv also writes (y > 0) as packed bits, and
        # that consumer's data gradient reads 1 bit instead of 16 per element (and one word per thread and tile, issued
        # before the accumulator wait).  Default: on with fp16 storage, where the tensor-core epilogues write / read the
        # bits in place (measured on B200, U-Net 512^2 batch 8: data gradients 1.30 -> 1.07 ms per step); the exact fp32
        # path would need an extra pass per tensor, so it keeps the activation tensor as its mask.
        self.relu_bits = (dt == F16 if relu_bits is None else bool(relu_bits)) and self.training
        self._bits = {}                  # id(tensor) -> Ref of its packed ReLU mask
        self._bias_done = set()          # id(conv layer) whose bias gradient is produced by another backward op
        # fp16 operand copies of the conv kernels: ONE pack launch per step for the whole model (OP_PACK_WEIGHTS)
        # instead of one small launch in front of every conv call
        self.prepack = bool(prepack) and dt == F16
        self.wpack = Arena("wpack")
        self.pack_entries = []           # [src element offset, dst element offset, mode, taps, J, K]
        self.layout = layout or ParamLayout(graph)
        self.act, self.f32, self.zero = Arena("act"), Arena("f32"), Arena("zero")
        self.fwd, self.bwd, self.opt = [], [], []
        self.views, self.gviews = {}, {}         # id(SymTensor) -> View
        self.layer_out = OrderedDict()           # layer name -> View (model.get_layer(name).output)
        self.n_dropout_ops = 0
        self.folded_into_next = set()            # conv layers whose stored output is already the following BN's
        self._lower()

    # ---------------------------------------------------------------------------------------
    # helpers
    # ---------------------------------------------------------------------------------------
    def _npix(self, v):
        return self.n * v.h * v.w

    def _packed(self, layer, mode, taps, j, k, k_src=0):
        """Ref of the packed fp16 copy of `layer`'s kernel for `mode` (include/b200unet.h b2u_pack_weights), or None
        when the tensor-core kernels cannot take the shape anyway.  k_src > 0: the kernel has only k_src input
        channels, the packed copy is zero-padded to k (first conv of an inference plan on a padded input)."""
        if not self.prepack or j % 16 or k % 16:
            return None
        ref = self.wpack.alloc(taps * j * k * 2)
        arena, off, _ = self.layout.offsets["%s/kernel" % layer.name]
        self.pack_entries.append([off, ref.off // 2, mode, taps, j, k, k_src])
        return ref

    def pack_table(self):
        """int64 (n, 8) table for OP_PACK_WEIGHTS: src, dst, first 32x32 work tile, mode, taps, J, K, source K (0 = K)"""
        import numpy as np
        tab = np.zeros((max(len(self.pack_entries), 1), 8), np.int64)
        start = 0
        for r, (src, dst, mode, taps, j, k, k_src) in enumerate(self.pack_entries):
            tab[r] = (src, dst, start, mode, taps, j, k, k_src)
            start += self._pack_tiles(taps, j, k)
        return tab

    @staticmethod
    def _pack_tiles(taps, j, k):
        return taps * ((j + 31) // 32) * ((k + 31) // 32)

    def _alloc_view(self, shape, dt, arena=None):
        if len(shape) == 3:
            h, w, c = shape
        else:
            h, w, c = 1, 1, shape[0]
        arena = arena or self.act
        ref = arena.alloc(self.n * h * w * c * ELEM[dt])
        return View(ref, c, c, h, w, dt)

    def _w(self, layer, short, arena=None):
        return self.layout.ref("%s/%s" % (layer.name, short), arena)

    def _bn_apply_op(self, xv, yv, aux, c, tag, lazy=None):
        if lazy is not None:                 # unmaterialised dropout on the input: p[6] = its keep bits, f[0] = rate
            return Op(OP_BN_APPLY, xv.dt, [xv.ref, yv.ref, aux["scale"], aux["shift"], None, None, lazy[2]],
                      [xv.ld, yv.ld, c, self._npix(xv), 0], [lazy[0]], tag=tag)
        if isinstance(xv, SplitView):
            a, b = xv.parts
            return Op(OP_BN_APPLY, xv.dt, [a.ref, yv.ref, aux["scale"], aux["shift"], None, b.ref],
                      [a.ld, yv.ld, c, self._npix(xv), 0, a.c, b.ld], tag=tag)
        return Op(OP_BN_APPLY, xv.dt, [xv.ref, yv.ref, aux["scale"], aux["shift"], None], [xv.ld, yv.ld, c, self._npix(xv), 0],
                  tag=tag)

    def _splittable(self, l):
        """concatenate of two tensors that only a BatchNormalization reads, that BN read by one 3x3 conv (its backward
        statistics then come from that conv's weight gradient): see `split_concat` in __init__"""
        if not self.split_concat or len(l.inputs) != 2:
            return False
        t = l.output
        if len(t.consumers) != 1 or t.consumers[0].kind != "batch_normalization" or t is self.graph.output:
            return False
        bn = t.consumers[0]
        cons = bn.output.consumers
        if len(cons) != 1 or cons[0].kind != "conv2d" or tuple(cons[0].kernel_size) != (3, 3) or bn.output is self.graph.output:
            return False
        # only where an input is narrower than a 128-byte line (fp16: fewer than 64 channels -- the 512^2 level of the U-Net);
        # wider inputs already fill whole lines in the interleaved buffer and the two-source BN kernels gain nothing
        if all(ti.channels * ELEM[self.dt] >= 128 for ti in l.inputs) and self.split_concat != 2:
            return False
        for ti in l.inputs:
            if ti.producer.kind not in ("conv2d_transpose", "batch_normalization") or ti.channels % 8:
                return False
            if sum(1 for cn in ti.consumers if cn.kind == "concatenate") != 1:
                return False
        return True

    def _bn_finalize(self, l, aux, c, count):
        """inference-mode BatchNormalization: scale / shift from the moving statistics -- a weight-only op"""
        for nm in ("scale", "shift", "mean", "invstd"):
            aux[nm] = self.f32.alloc(c * 4)
        aux["count"] = count
        op = Op(OP_BN_FINALIZE, 0,
                [None, self._w(l, "gamma"), self._w(l, "beta"), self._w(l, "moving_mean"), self._w(l, "moving_variance"),
                 aux["scale"], aux["shift"], aux["mean"], aux["invstd"]], [count, 0, c], [l.momentum, l.epsilon], tag=l.name)
        (self.prep if self.hoist_prep else self.fwd).append(op)

    @staticmethod
    def _act_of(t):
        l = t.producer
        a = getattr(l, "activation", None)
        return a if a in ("relu", "elu") else None

    def _mask_for(self, t):
        """(mask view, act code) if the gradient written for tensor t must be multiplied by its
        producer's activation derivative, else (None, 0)."""
        # look through identity dropouts (inference / dropout disabled / folded into the max-pool)
        while t.producer.kind == "dropout" and self.views[id(t)] is self.views[id(t.producer.inputs[0])]:
            t = t.producer.inputs[0]
        a = self._act_of(t)
        if a is None:
            return None, 0
        if len(t.consumers) != 1:
            raise NotImplementedError("activated tensor %r with %d consumers" % (t, len(t.consumers)))
        return self.views[id(t)], ACT[a]

    def _bias_sink(self, t):
        """The op that writes the FINAL gradient of tensor t (activation derivative applied, single writer) can also
        emit its per-channel sums: that is the bias gradient of the conv that produced t, so the weight-gradient op
        need not read the whole gradient again.  Returns the Ref of that bias gradient (and remembers the layer),
        or None when t is not such a tensor."""
        if not self.fuse_bias_grad:
            return None
        while t.producer.kind == "dropout" and self.views[id(t)] is self.views[id(t.producer.inputs[0])]:
            t = t.producer.inputs[0]
        l = t.producer
        if l.kind != "conv2d" or tuple(l.kernel_size) != (3, 3) or len(t.consumers) != 1 or t.channels % 8:
            return None
        self._bias_done.add(id(l))
        return self._w(l, "bias", "grads")

    # ---------------------------------------------------------------------------------------
    # lowering
    # ---------------------------------------------------------------------------------------
    def _lower(self):
        g, n, dt = self.graph, self.n, self.dt
        layers = g.layers
        # ---- placement: a tensor consumed by concatenate lives inside its LAST concat's buffer ----
        home = {}        # id(tensor) -> (concat layer, channel offset)
        split = set()    # id(concat layer) kept as two dense tensors (SplitView)
        for l in layers:
            if l.kind == "concatenate" and self._splittable(l):
                split.add(id(l))
        self.split_concats = sorted(l.name for l in layers if id(l) in split)
        for l in layers:
            if l.kind == "concatenate" and id(l) not in split:
                off = 0
                for t in l.inputs:
                    if t.producer.kind in ("concatenate", "flatten", "input"):
                        raise NotImplementedError("concatenate of %s outputs" % t.producer.kind)
                    home[id(t)] = (l, off)       # later concats overwrite earlier ones: last wins
                    off += t.channels
        concat_buf, concat_gbuf = {}, {}
        for l in layers:
            if l.kind == "concatenate" and id(l) not in split:
                concat_buf[id(l)] = self._alloc_view(l.output.shape, dt)
                if self.training:
                    concat_gbuf[id(l)] = self._alloc_view(l.output.shape, dt)

        def place(t, tdt):
            if id(t) in home:
                cl, off = home[id(t)]
                self.views[id(t)] = concat_buf[id(cl)].slice(off, t.channels)
                if self.training:
                    self.gviews[id(t)] = concat_gbuf[id(cl)].slice(off, t.channels)
            else:
                self.views[id(t)] = self._alloc_view(t.shape, tdt)
                if self.training:
                    self.gviews[id(t)] = self._alloc_view(t.shape, tdt)

        self.step_ref = Ref("step", 0)
        self.x_view = None
        drop_index = {}
        k = 0
        for l in layers:
            if l.kind == "dropout":
                drop_index[id(l)] = k
                k += 1
        self.n_dropout_ops = k
        prod_op = {}             # id(tensor) -> forward Op that can emit BN statistics of what it writes
        fused_drop = set()       # dropout layers folded into the preceding max-pool
        bn_aux = {}              # id(bn layer) -> dict of small buffers
        written = set()          # id(tensor) whose gradient view has been written (backward)

        # ======================================= forward ==========================================
        for l in layers:
            t = l.output
            if l.kind == "input":
                # inference plans with fp16 storage: an input of 2..15 channels feeding ONE 3x3 conv is stored zero-padded
                # to 16 channels, so that conv runs on the tensor cores (K = 16, the kernel's packed copy padded alike)
                # instead of the CUDA-core kernel (classifier on 224 x 224 x 3, batch 64: 0.18 ms -> 0.07 ms on B200)
                cin = t.shape[-1]
                self.x_cin, self.x_pad = cin, 0
                if (not self.training and dt == F16 and self.prepack and 1 < cin < 16 and len(t.consumers) == 1 and
                        t.consumers[0].kind == "conv2d" and tuple(t.consumers[0].kernel_size) == (3, 3) and
                        t.consumers[0].filters % 16 == 0):
                    self.x_pad = 16
                    self.views[id(t)] = self._alloc_view(tuple(t.shape[:-1]) + (16,), dt)
                else:
                    self.views[id(t)] = self._alloc_view(t.shape, dt)
                self.x_view = self.views[id(t)]
                self.layer_out[l.name] = self.x_view
                continue
            x = l.inputs[0]
            xv = self.views[id(x)]
            if l.kind == "conv2d" and l.kernel_size == (3, 3):
                place(t, dt)
                yv = self.views[id(t)]
                stats = None
                cons = t.consumers
                if self.training and len(cons) == 1 and cons[0].kind == "batch_normalization":
                    stats = self.zero.alloc(2 * t.channels * 8)
                    bn_aux[id(cons[0])] = {"sums": stats, "fused": True}
                bits = None
                if (self.relu_bits and l.activation == "relu" and len(cons) == 1 and cons[0].kind == "conv2d" and
                        tuple(cons[0].kernel_size) == (3, 3) and yv.c % 16 == 0):
                    bits = self._bits[id(t)] = self.act.alloc(self._npix(yv) * yv.c // 8)
                k_src = x.shape[-1] if (x.producer.kind == "input" and self.x_pad) else 0
                post = [None, None]
                bn_ld = (concat_buf[id(home[id(cons[0].output)][0])].ld
                         if len(cons) == 1 and id(cons[0].output) in home else yv.c)
                if (self.fuse_bn_infer and self.prepack and len(cons) == 1 and cons[0].kind == "batch_normalization" and
                        xv.c % 16 == 0 and yv.c % 16 == 0 and yv.c <= 1024 and xv.ld % 8 == 0 and bn_ld % 8 == 0):
                    # the BN's output takes the conv's place: the conv writes normalised values, the BN layer emits nothing
                    bn = cons[0]
                    aux = bn_aux.setdefault(id(bn), {})
                    self._bn_finalize(bn, aux, yv.c, self._npix(yv))
                    aux["folded"] = True
                    place(bn.output, dt)
                    yv = self.views[id(t)] = self.views[id(bn.output)]
                    post = [aux["scale"], aux["shift"]]
                    self.folded_into_next.add(l.name)
                self.fwd.append(Op(OP_CONV3X3_FWD, dt, [xv.ref, self._w(l, "kernel"), self._w(l, "bias"), yv.ref, stats,
                                                        self._packed(l, 0, 9, yv.c, xv.c, k_src), bits] + post,
                                   [xv.ld, xv.c, ACT[l.activation], yv.ld, yv.c, n, xv.h, xv.w, 1 if self.training else 0,
                                    k_src], tag=l.name))
                # i[8]: training-mode op (kernel selection may trade a rounding for speed); i[9]: the kernel's real input
                # channel count when the input tensor is zero-padded (0 = not padded)
            elif l.kind == "conv2d":      # 1x1 output head
                if l.activation != "sigmoid" or l.filters != 1 or t is not g.output:
                    raise NotImplementedError("1x1 conv is supported as the sigmoid output head only")
                self.prob = self.f32.alloc(self._npix(xv) * 4)
                self.prob_shape = (n, xv.h, xv.w, 1)
                self.views[id(t)] = View(self.prob, 1, 1, xv.h, xv.w, F32)
                self.fwd.append(Op(OP_HEAD_FWD, dt, [xv.ref, self._w(l, "kernel"), self._w(l, "bias"), self.prob],
                                   [xv.ld, xv.c, self._npix(xv)], tag=l.name))
            elif l.kind == "conv2d_transpose":
                place(t, dt)
                yv = self.views[id(t)]
                self.fwd.append(Op(OP_CONVT_FWD, dt, [xv.ref, self._w(l, "kernel"), self._w(l, "bias"), yv.ref, None,
                                                      self._packed(l, 2, 1, 4 * yv.c, xv.c)],
                                   [xv.ld, xv.c, yv.ld, yv.c, n, xv.h, xv.w, 0], tag=l.name))
                prod_op[id(t)] = self.fwd[-1]
            elif l.kind == "batch_normalization":
                aux = bn_aux.setdefault(id(l), {})
                if aux.get("folded"):                          # inference: applied by the producing conv's epilogue
                    self.layer_out[l.name] = self.views[id(t)]
                    continue
                place(t, xv.dt)
                yv = self.views[id(t)]
                c = xv.c
                count = self._npix(xv)
                if not self.training:
                    self._bn_finalize(l, aux, c, count)
                    self.fwd.append(self._bn_apply_op(xv, yv, aux, c, l.name))
                    self.layer_out[l.name] = yv
                    continue
                for nm in ("scale", "shift", "mean", "invstd"):
                    aux[nm] = self.f32.alloc(c * 4)
                if self.training:
                    if "sums" not in aux:
                        aux["sums"] = self.zero.alloc(2 * c * 8)
                        src = x.producer
                        fusable = (self.fuse_bn_stats and src.kind == "concatenate" and
                                   all((id(src) in split or home[id(ti)][0] is src) and id(ti) in prod_op and
                                       prod_op[id(ti)].p[4] is None for ti in src.inputs))
                        if fusable:
                            # BN over a concat buffer: every producer (transposed conv epilogue, encoder BN apply)
                            # accumulates the statistics of the channel slice it writes -- no pass over the buffer
                            off = 0
                            for ti in src.inputs:
                                po = prod_op[id(ti)]
                                po.p[4] = aux["sums"] + off * 8
                                po.i[7 if po.kind == OP_CONVT_FWD else 4] = c
                                off += ti.channels
                        elif isinstance(xv, SplitView):            # one statistics pass per dense half
                            off = 0
                            for pv in xv.parts:
                                self.fwd.append(Op(OP_BN_STATS, pv.dt, [pv.ref, aux["sums"] + off * 8], [pv.ld, pv.c, count, c],
                                                   tag=l.name))
                                off += pv.c
                        elif id(x) in self._lazy_drop:             # statistics of dropout(x), the mask regenerated
                            rate, op_id, bits = self._lazy_drop[id(x)]
                            self.fwd.append(Op(OP_BN_STATS, xv.dt, [xv.ref, aux["sums"], self.step_ref, bits],
                                               [xv.ld, c, count, 0, op_id], [rate], tag=l.name))
                        else:
                            self.fwd.append(Op(OP_BN_STATS, xv.dt, [xv.ref, aux["sums"]], [xv.ld, c, count], tag=l.name))
                    if self.sync_stats:
                        self.fwd.append(Op(OP_ALLREDUCE_F64, 0, [aux["sums"]], [2 * c], tag=l.name))
                        count *= self.world
                aux["count"] = count
                self.fwd.append(Op(OP_BN_FINALIZE, 0,
                                   [aux.get("sums"), self._w(l, "gamma"), self._w(l, "beta"), self._w(l, "moving_mean"),
                                    self._w(l, "moving_variance"), aux["scale"], aux["shift"], aux["mean"], aux["invstd"]],
                                   [count, 1 if self.training else 0, c], [l.momentum, l.epsilon], tag=l.name))
                self.fwd.append(self._bn_apply_op(xv, yv, aux, c, l.name, self._lazy_drop.get(id(x))))
                if self.training:
                    prod_op[id(t)] = self.fwd[-1]
            elif l.kind == "max_pooling2d":
                place(t, xv.dt)
                yv = self.views[id(t)]
                p_drop, op_id = 0.0, 0
                cons = t.consumers
                if len(cons) == 1 and cons[0].kind == "dropout" and id(t) not in home:
                    fused_drop.add(id(cons[0]))
                    if self.dropout:
                        p_drop, op_id = cons[0].rate, drop_index[id(cons[0])]
                l._pool_drop = (p_drop, op_id)
                prev = self.fwd[-1] if self.fwd else None
                if (self.fuse_bn_pool and prev is not None and prev.kind == OP_BN_APPLY and
                        x.producer.kind == "batch_normalization" and prev.tag == x.producer.name and
                        xv.h % 2 == 0 and xv.w % 2 == 0):
                    # the BN apply that has just written x becomes one pass that also writes its 2x2 max (+ dropout):
                    # x is not read back (same Op object: later concat-BN statistics fusion patches p[4] / i[4])
                    prev.kind = OP_BN_APPLY_POOL
                    prev.p = prev.p[:5] + [yv.ref, self.step_ref if p_drop > 0 else None]
                    prev.i = prev.i[:5] + [n, xv.h, xv.w, yv.ld, op_id]
                    prev.f = [float(p_drop)]
                    prev.tag = prev.tag + "+" + l.name
                else:
                    self.fwd.append(Op(OP_MAXPOOL_FWD, xv.dt, [xv.ref, yv.ref, self.step_ref if p_drop > 0 else None],
                                       [xv.ld, yv.ld, xv.c, n, xv.h, xv.w, op_id], [p_drop], tag=l.name))
            elif l.kind == "dropout":
                if id(l) in fused_drop or not self.dropout:
                    if id(t) in home:
                        raise NotImplementedError("identity dropout feeding a concatenate")
                    self.views[id(t)] = xv                     # alias: identity (fused or inference)
                    if self.training:
                        self.gviews[id(t)] = self.gviews[id(x)]
                elif (self.fuse_dropout_bn and self.training and id(t) not in home and len(t.consumers) == 1 and
                      t.consumers[0].kind == "batch_normalization" and len(x.consumers) == 1 and
                      not isinstance(xv, SplitView) and xv.c % 8 == 0 and
                      not (len(t.consumers[0].output.consumers) == 1 and
                           t.consumers[0].output.consumers[0].kind == "max_pooling2d")):
                    # not materialised: the BatchNormalization behind it applies the mask on the fly (aliases like an
                    # identity dropout; `_lazy_drop` tells the BN ops)
                    self.views[id(t)] = xv
                    self.gviews[id(t)] = self.gviews[id(x)]
                    # (rate, dropout op index, packed keep mask: written by the BN statistics pass, read by the others)
                    self._lazy_drop[id(t)] = (float(l.rate), drop_index[id(l)], self.act.alloc(self._npix(xv) * xv.c // 8))
                else:
                    place(t, xv.dt)
                    yv = self.views[id(t)]
                    self.fwd.append(Op(OP_DROPOUT_FWD, xv.dt, [xv.ref, yv.ref, self.step_ref],
                                       [xv.ld, yv.ld, xv.c, self._npix(xv), drop_index[id(l)]], [l.rate], tag=l.name))
            elif l.kind == "concatenate" and id(l) in split:
                self.views[id(t)] = SplitView(*(self.views[id(ti)] for ti in l.inputs))
                if self.training:
                    self.gviews[id(t)] = SplitView(*(self.gviews[id(ti)] for ti in l.inputs))
            elif l.kind == "concatenate":
                buf = concat_buf[id(l)]
                self.views[id(t)] = buf
                if self.training:
                    self.gviews[id(t)] = concat_gbuf[id(l)]
                off = 0
                for ti in l.inputs:
                    if home[id(ti)][0] is not l:               # lives in a later concat: copy in
                        sv = self.views[id(ti)]
                        dv = buf.slice(off, ti.channels)
                        self.fwd.append(Op(OP_COPY_SLICE, dt, [sv.ref, dv.ref], [sv.ld, dv.ld, sv.c, self._npix(sv), 0],
                                           tag=l.name))
                    off += ti.channels
            elif l.kind == "flatten":
                if xv.ld != xv.c:
                    raise NotImplementedError("flatten of a strided view")
                kf = xv.h * xv.w * xv.c
                self.views[id(t)] = View(xv.ref, kf, kf, 1, 1, xv.dt)
                if self.training:
                    gx = self.gviews[id(x)]
                    self.gviews[id(t)] = View(gx.ref, kf, kf, 1, 1, gx.dt)
            elif l.kind == "dense":
                yv = self._alloc_view((l.units,), F32, self.f32)
                self.views[id(t)] = yv
                if self.training:
                    self.gviews[id(t)] = self._alloc_view((l.units,), F32, self.f32)
                if t is g.output:
                    if l.activation != "sigmoid" or l.units != 1:
                        raise NotImplementedError("dense output head must be Dense(1, sigmoid)")
                    self.prob, self.prob_shape = yv.ref, (n, 1)
                self.fwd.append(Op(OP_DENSE_FWD, xv.dt, [xv.ref, self._w(l, "kernel"), self._w(l, "bias"), yv.ref],
                                   [xv.c, ACT[l.activation], l.units, n], tag=l.name))
            else:
                raise NotImplementedError(l.kind)
            self.layer_out[l.name] = self.views[id(t)]

        # ---- loss / metric ------------------------------------------------------------------------
        out_elems = int(math.prod(self.prob_shape))
        self.target = self.f32.alloc(out_elems * 4)
        self.sample_w = self.f32.alloc(n * 4)
        self.loss_out = self.f32.alloc(16)
        self.loss_ops = []
        count = out_elems
        if self.loss == "bce_dice":
            self.loss_sums = self.zero.alloc(4 * 8)
            self.loss_ops.append(Op(OP_BCE_DICE_SUMS, 0, [self.prob, self.target, self.loss_sums], [out_elems], tag="loss"))
            if self.sync_stats:
                self.loss_ops.append(Op(OP_ALLREDUCE_F64, 0, [self.loss_sums], [4], tag="loss"))
                count *= self.world
            self.loss_ops.append(Op(OP_BCE_DICE_FINALIZE, 0, [self.loss_sums, self.loss_out], [count], tag="loss"))
        elif self.loss == "bce":
            self.loss_ops.append(Op(OP_BCE_FWD, 0, [self.prob, self.target, self.sample_w, self.loss_out], [n], tag="loss"))
        else:
            raise NotImplementedError(self.loss)
        self.loss_count = count
        if not self.training:
            return

        # ======================================= backward =========================================
        grads = lambda l, s: self._w(l, s, "grads")
        yv_c = lambda l: self.views[id(l.output)].c
        for l in reversed(layers):
            t = l.output
            if l.kind == "input":
                continue
            x = l.inputs[0]
            xv = self.views[id(x)]
            x_is_input = x.producer.kind == "input"
            if l.kind == "conv2d" and l.kernel_size == (1, 1):
                gx = self.gviews[id(x)]
                mv, ma = self._mask_for(x)
                self.bwd.append(Op(OP_HEAD_BWD, dt, [self.prob, self.target, self.loss_sums, self.step_ref, xv.ref,
                                                     self._w(l, "kernel"), gx.ref, grads(l, "kernel"), grads(l, "bias"),
                                                     self._bias_sink(x) if mv is not None else None],
                                   [self.loss_count, xv.ld, xv.c, gx.ld, ma, self._npix(xv)], tag=l.name))
                written.add(id(x))
                continue
            if l.kind == "dense" and t is g.output:
                gy = self.gviews[id(t)]
                self.bwd.append(Op(OP_BCE_SIGMOID_BWD, F32, [self.prob, self.target, self.sample_w, self.step_ref, gy.ref],
                                   [n], tag=l.name))
                written.add(id(t))
            if id(t) not in written:
                if l.kind in ("max_pooling2d",) and not t.consumers:
                    continue                                    # dead branch (UPP:912 p4)
                raise RuntimeError("no gradient reached %s" % l.name)
            gy = self.gviews[id(t)]
            if l.kind == "conv2d":
                self.bwd.append(Op(OP_CONV3X3_WGRAD, dt | OPF_SIDE, [xv.ref, gy.ref, grads(l, "kernel"),
                                                          None if id(l) in self._bias_done else grads(l, "bias")],
                                   [xv.ld, xv.c, gy.ld, gy.c, n, xv.h, xv.w], tag=l.name))
                if not x_is_input:
                    gx = self.gviews[id(x)]
                    mv, ma = self._mask_for(x)
                    acc = 1 if id(x) in written else 0
                    sink = self._bias_sink(x) if (mv is not None and not acc) else None
                    mref, mld = (mv.ref, mv.ld) if mv else (None, 0)
                    xm = x                                    # the activated tensor behind identity dropouts
                    while xm.producer.kind == "dropout" and self.views[id(xm)] is self.views[id(xm.producer.inputs[0])]:
                        xm = xm.producer.inputs[0]
                    if mv is not None and not acc and id(xm) in self._bits:
                        mref, mld, ma = self._bits[id(xm)], gx.c, ACT_RELU_BITS
                    bn = x.producer
                    if (self.fuse_bn_bwd_wgrad and sink is None and not acc and bn.kind == "batch_normalization" and
                            len(x.consumers) == 1 and "bwd_sums" not in bn_aux[id(bn)] and x.channels % 8 == 0):
                        # x is a BatchNorm output read by this conv only: the BN backward statistics follow from this
                        # conv's weight gradient (<W, dW> per input channel) and the column sums of the gradient written
                        # here -- no pass over dy and the BN input (include/b200unet.h b2u_bn_bwd_sums_from_wgrad)
                        sink = self.zero.alloc(x.channels * 4)
                        bn_aux[id(bn)]["wgrad_sums"] = (sink, self._w(l, "kernel"), grads(l, "kernel"), yv_c(l))
                    self.bwd.append(Op(OP_CONV3X3_DGRAD, dt, [gy.ref, self._w(l, "kernel"), gx.ref, mref, sink,
                                                              self._packed(l, 1, 9, gx.c, gy.c)],
                                       [gy.ld, gy.c, gx.ld, gx.c, mld, ma, acc,
                                        n, xv.h, xv.w], tag=l.name))
                    written.add(id(x))
            elif l.kind == "conv2d_transpose":
                # a transposed conv whose output only feeds (through a concatenate) a training-mode BatchNorm has an
                # analytically zero bias gradient: the BN backward output sums to zero over every channel
                zero_db = (self.fuse_bias_grad and len(t.consumers) == 1 and t.consumers[0].kind == "concatenate" and
                           len(t.consumers[0].output.consumers) == 1 and
                           t.consumers[0].output.consumers[0].kind == "batch_normalization")
                self.bwd.append(Op(OP_CONVT_WGRAD, dt | OPF_SIDE, [xv.ref, gy.ref, grads(l, "kernel"),
                                                        None if zero_db else grads(l, "bias")],
                                   [xv.ld, xv.c, gy.ld, gy.c, n, xv.h, xv.w], tag=l.name))
                gx = self.gviews[id(x)]
                mv, ma = self._mask_for(x)
                acc = 1 if id(x) in written else 0
                sink = self._bias_sink(x) if (mv is not None and not acc) else None
                self.bwd.append(Op(OP_CONVT_DGRAD, dt, [gy.ref, self._w(l, "kernel"), gx.ref, mv.ref if mv else None, sink,
                                                        self._packed(l, 3, 4, gx.c, gy.c)],
                                   [gy.ld, gy.c, gx.ld, gx.c, mv.ld if mv else 0, ma, acc,
                                    n, xv.h, xv.w], tag=l.name))
                written.add(id(x))
            elif l.kind == "batch_normalization":
                aux = bn_aux[id(l)]
                c = xv.c
                if id(x) in written:
                    raise NotImplementedError("BN backward accumulate")
                gx = self.gviews[id(x)]
                mv, ma = self._mask_for(x)
                if "bwd_sums" in aux:                       # produced by the max-pool backward that completed gy
                    bsums = aux["bwd_sums"]
                elif "wgrad_sums" in aux:                   # from the consumer conv's weight gradient
                    csum, wref, dwref, cout = aux["wgrad_sums"]
                    bsums = self.zero.alloc(2 * c * 8)
                    self.bwd.append(Op(OP_BN_BWD_SUMS_WGRAD, OPF_JOIN, [wref, dwref, csum, self._w(l, "gamma"), self._w(l, "beta"),
                                                                 bsums], [c, cout, 9], tag=l.name))
                    if self.sync_stats:
                        self.bwd.append(Op(OP_ALLREDUCE_F64, 0, [bsums], [2 * c], tag=l.name))
                elif isinstance(xv, SplitView):           # one reduction pass per dense half, into the shared sums
                    bsums = self.zero.alloc(2 * c * 8)
                    off = 0
                    for pv in xv.parts:
                        gs = View(gy.ref + off * ELEM[gy.dt], gy.ld, pv.c, gy.h, gy.w, gy.dt)
                        self.bwd.append(Op(OP_BN_BWD_REDUCE, pv.dt, [gs.ref, pv.ref, aux["mean"] + off * 4, aux["invstd"] + off * 4,
                                                                     bsums + off * 8], [gs.ld, pv.ld, pv.c, self._npix(pv), c],
                                           tag=l.name))
                        off += pv.c
                    if self.sync_stats:
                        self.bwd.append(Op(OP_ALLREDUCE_F64, 0, [bsums], [2 * c], tag=l.name))
                else:
                    bsums = self.zero.alloc(2 * c * 8)
                    lazy = self._lazy_drop.get(id(x))
                    self.bwd.append(Op(OP_BN_BWD_REDUCE, xv.dt, [gy.ref, xv.ref, aux["mean"], aux["invstd"], bsums] +
                                       ([lazy[2]] if lazy else []),
                                       [gy.ld, xv.ld, c, self._npix(xv)], [lazy[0]] if lazy else [], tag=l.name))
                    if self.sync_stats:
                        self.bwd.append(Op(OP_ALLREDUCE_F64, 0, [bsums], [2 * c], tag=l.name))
                # npix is the LOCAL pixel count; the divisor (count) is the global one under sync_stats,
                # where the all-reduced sums must enter dgamma/dbeta on one rank only (grads are summed)
                own = (not self.sync_stats) or self.rank == 0
                if isinstance(xv, SplitView):               # two dense (input, gradient) pairs: p[11], p[12], i[8..10]
                    (xa, xb), (ga, gb) = xv.parts, gx.parts
                    assert mv is None
                    self.bwd.append(Op(OP_BN_BWD_APPLY, xv.dt,
                                       [gy.ref, xa.ref, ga.ref, self._w(l, "gamma"), aux["mean"], aux["invstd"], bsums,
                                        grads(l, "gamma") if own else None, grads(l, "beta") if own else None, None, None,
                                        xb.ref, gb.ref],
                                       [gy.ld, xa.ld, ga.ld, c, self._npix(xv), 0, 0, aux["count"], xa.c, xb.ld, gb.ld],
                                       tag=l.name))
                    written.add(id(x))
                    continue
                lazy = self._lazy_drop.get(id(x))     # unmaterialised dropout: p[13] = its keep bits, f[0] = rate
                self.bwd.append(Op(OP_BN_BWD_APPLY, xv.dt,
                                   [gy.ref, xv.ref, gx.ref, self._w(l, "gamma"), aux["mean"], aux["invstd"], bsums,
                                    grads(l, "gamma") if own else None, grads(l, "beta") if own else None,
                                    mv.ref if mv else None, self._bias_sink(x) if mv is not None else None] +
                                   ([None, None, lazy[2]] if lazy else []),
                                   [gy.ld, xv.ld, gx.ld, c, self._npix(xv), mv.ld if mv else 0, ma, aux["count"]],
                                   [lazy[0]] if lazy else [],
                                   tag=l.name))
                written.add(id(x))
            elif l.kind == "max_pooling2d":
                gx = self.gviews[id(x)]
                p_drop, op_id = l._pool_drop
                if self._act_of(x) is not None:
                    raise NotImplementedError("max-pool directly after an activated conv")
                # when this pool is the LAST contribution to the gradient of a BatchNorm output, the kernel also
                # produces that BN's backward statistics (sum dx, sum dx*xhat with xhat = (x - beta)/gamma) from the
                # values it already holds: no separate BN_BWD_REDUCE pass over dx and the BN input
                fuse = None
                bn = x.producer
                if bn.kind == "batch_normalization" and all(cn._seq >= l._seq for cn in x.consumers) and self.fuse_bn_bwd:
                    fuse = self.zero.alloc(2 * xv.c * 8)
                    bn_aux[id(bn)]["bwd_sums"] = fuse
                self.bwd.append(Op(OP_MAXPOOL_BWD, xv.dt,
                                   [xv.ref, gy.ref, gx.ref, self.step_ref if p_drop > 0 else None, fuse,
                                    self._w(bn, "gamma") if fuse else None, self._w(bn, "beta") if fuse else None],
                                   [xv.ld, gy.ld, gx.ld, xv.c, n, xv.h, xv.w, op_id, 1 if id(x) in written else 0],
                                   [p_drop], tag=l.name))
                if fuse is not None and self.sync_stats:
                    self.bwd.append(Op(OP_ALLREDUCE_F64, 0, [fuse], [2 * xv.c], tag=l.name))
                written.add(id(x))
            elif l.kind == "dropout":
                if self.views[id(t)] is xv:                     # identity alias
                    written.add(id(x))
                    continue
                if id(x) in written:
                    raise NotImplementedError("dropout backward accumulate")
                gx = self.gviews[id(x)]
                mv, ma = self._mask_for(x)
                self.bwd.append(Op(OP_DROPOUT_BWD, xv.dt, [gy.ref, gx.ref, self.step_ref, mv.ref if mv else None],
                                   [gy.ld, gx.ld, xv.c, self._npix(xv), drop_index[id(l)], mv.ld if mv else 0, ma],
                                   [l.rate], tag=l.name))
                written.add(id(x))
            elif l.kind == "concatenate" and id(l) in split:
                for ti in l.inputs:                             # the BN backward wrote both gradient tensors in place
                    if id(ti) in written:
                        raise RuntimeError("split concat must be the last consumer of %r" % ti)
                    written.add(id(ti))
            elif l.kind == "concatenate":
                off = 0
                for ti in l.inputs:
                    if home[id(ti)][0] is l:
                        if id(ti) in written:
                            raise RuntimeError("home concat must be the last consumer of %r" % ti)
                        written.add(id(ti))                     # gradient slice already in place
                    else:
                        sv = gy.slice(off, ti.channels)
                        dv = self.gviews[id(ti)]
                        if id(ti) not in written:
                            raise RuntimeError("copy-concat gradient before home concat for %r" % ti)
                        if self._act_of(ti) is not None:
                            raise NotImplementedError("activated tensor feeding several concats")
                        self.bwd.append(Op(OP_COPY_SLICE, dt, [sv.ref, dv.ref], [sv.ld, dv.ld, sv.c, self._npix(sv), 1],
                                           tag=l.name))
                    off += ti.channels
            elif l.kind == "flatten":
                written.add(id(x))
            elif l.kind == "dense":
                gx = self.gviews[id(x)]
                mv, ma = self._mask_for(x)
                if id(x) in written:
                    raise NotImplementedError("dense backward accumulate")
                yv = self.views[id(t)]
                # gy already holds the pre-activation gradient (mask applied by the consumer / loss)
                self.bwd.append(Op(OP_DENSE_BWD, xv.dt, [xv.ref, self._w(l, "kernel"), yv.ref, gy.ref, gx.ref,
                                                         mv.ref if mv else None, grads(l, "kernel"), grads(l, "bias")],
                                   [xv.c, 0, ma, l.units, n], tag=l.name))
                written.add(id(x))
            else:
                raise NotImplementedError(l.kind)

        # ======================================= optimizer ========================================
        npar = self.layout.n_params
        if self.world > 1:
            if self.grad_bucket_bytes > 0:
                self._bucket_allreduce(npar)
            else:
                self.opt.append(Op(OP_ALLREDUCE_F32, OPF_JOIN, [Ref("grads", 0)], [npar], tag="allreduce"))
        self.opt.append(Op(OP_ADAM, OPF_JOIN, [Ref("params", 0), Ref("grads", 0), Ref("adam_m", 0), Ref("adam_v", 0), self.step_ref],
                           [npar], tag="adam"))
        self.opt.append(Op(OP_STATE_ADVANCE, 0, [self.step_ref], tag="advance"))

    # ---------------------------------------------------------------------------------------
    def _bucket_allreduce(self, npar):
        """Gradient exchange overlapped with the backward pass (VERDICT r1 missing 5).

        The flat gradient buffer is in Keras weight order = forward layer order, and the backward pass completes it from
        the END: the buffer is cut, walking backwards, into contiguous buckets of >= grad_bucket_bytes, and the all-reduce
        of a bucket is placed right behind the LAST backward op that writes any gradient inside it (weight gradients,
        bias column sums written by a consumer's data gradient, BatchNorm affine gradients, the head).  The executor
        issues B2U_OPF_COMM ops on its side stream, so the exchange of the deep layers (c5a..c6a hold 61 % of the U-Net's
        parameters and are complete mid-backward) runs under the remaining data / weight gradient kernels; only the
        small bucket of the first layers is exposed.  Adam (B2U_OPF_JOIN) waits for all of them."""
        # last writer (index into self.bwd) of every 4-byte word range of the gradient buffer that some op writes
        tensors = sorted((off, int(math.prod(shape))) for a, off, shape in self.layout.offsets.values() if a == "params")
        starts = [t[0] for t in tensors]
        last = [-1] * len(tensors)
        import bisect
        for k, op in enumerate(self.bwd):
            for ref in op.p:
                if isinstance(ref, Ref) and ref.arena == "grads":
                    j = bisect.bisect_right(starts, ref.off // 4) - 1
                    last[j] = max(last[j], k)
        buckets, hi, ready = [], npar, -1          # [lo, hi) element ranges, walking from the end of the buffer
        for j in range(len(tensors) - 1, -1, -1):
            ready = max(ready, last[j])
            lo = tensors[j][0]
            if (hi - lo) * 4 >= self.grad_bucket_bytes or j == 0:
                buckets.append((0 if j == 0 else lo, hi, ready))
                hi, ready = lo, -1
        self.grad_buckets = buckets
        inserts = {}
        for lo, hi, ready in buckets:
            inserts.setdefault(ready, []).append(Op(OP_ALLREDUCE_F32, OPF_COMM, [Ref("grads", lo * 4)], [hi - lo],
                                                    tag="allreduce[%d:%d]" % (lo, hi)))
        out = list(inserts.get(-1, []))            # (a bucket nothing writes: exchanged up front, harmless)
        for k, op in enumerate(self.bwd):
            out.append(op)
            out.extend(inserts.get(k, []))
        self.bwd = out

    def _pack_op(self):
        if not self.pack_entries:
            return []
        total = sum(self._pack_tiles(t, j, k) for _, _, _, t, j, k, _ks in self.pack_entries)
        return [Op(OP_PACK_WEIGHTS, 0, [Ref("wtab", 0), Ref("params", 0), Ref("wpack", 0)],
                   [len(self.pack_entries), total], tag="pack-weights")]

    def prep_ops(self):
        """inference plans: the ops that depend on the weights only (fp16 operand copies, BN scale / shift from the moving
        statistics); they must run again whenever the weights -- or another plan sharing the arenas -- have changed."""
        return (self._pack_op() + self.prep) if self.hoist_prep else []

    def prologue(self, prep=True):
        """ops that must run before every step: clear the statistics arena (and gradients)."""
        ops = []
        if self.hoist_prep:
            ops += self.prep_ops() if prep else []
        else:
            ops += self._pack_op()
        if self.zero.size:
            ops.append(Op(OP_MEMSET, 0, [Ref("zero", 0)], [self.zero.size], tag="zero-sums"))
        if self.training:
            ops.append(Op(OP_MEMSET, 0, [Ref("grads", 0)], [self.layout.n_params * 4], tag="zero-grads"))
        return ops

    def train_ops(self):
        return self.prologue() + self.fwd + self.loss_ops + self.bwd + self.opt

    def forward_ops(self, with_loss=False, prep=True):
        """prep=False: without the weight-only ops of an inference plan (the caller has run `prep_ops()`)"""
        return self.prologue(prep) + self.fwd + (self.loss_ops if with_loss else [])

    def arena_sizes(self):
        return {"act": self.act.size, "f32": self.f32.size, "zero": max(self.zero.size, 8),
                "wpack": max(self.wpack.size, 8),
                "params": self.layout.n_params * 4, "state": max(self.layout.n_state * 4, 4),
                "step": STEP_STATE_BYTES}
