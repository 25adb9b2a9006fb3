"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the reference's Keras hot path.

PARITY UNPINNED: the reference (deadskull7/One-Stop-for-COVID-19-...) ships no tests, no golden
vectors and cannot run offline (TensorFlow 2.2 / Keras 2.3 / segmentation_models are absent, see
SURVEY.md section 8c).  The arithmetic lives in those un-vendored dependencies, so this file restates
their *published* semantics and is anchored on the only facts the reference
does pin: parameter counts and per-layer output shapes of its stored `model.summary()` tables
(tests/test_oracle.py), plus self-checks (finite differences, fp64-vs-fp32 agreement).

Reference call sites restated here (paths relative to /root/reference/Scripts):
  * U-Net graph ............ task1_preprocessing_plus_unet_with_comments.py:853-915 (= task3:850-912)
  * U-Net++ graph .......... task1_unet_plus_plus.py:860-949
  * classifier graph ....... task2_covid19_classifcation.py:747-778
  * dice_coeff / bce_dice .. task1_preprocessing_plus_unet_with_comments.py:784-799
  * recall/precision/f1 .... task2_covid19_classifcation.py:688-703
  * Adam(lr=5e-4) .......... task1_preprocessing_plus_unet_with_comments.py:1053
  * CosineAnnealing ........ task1_preprocessing_plus_unet_with_comments.py:970-996
  * sm.metrics thresholds .. task1_preprocessing_plus_unet_with_comments.py:1206-1211

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  It is never on the product path.

Layouts are Keras-native: activations NHWC, Conv2D kernels (kh,kw,Cin,Cout), Conv2DTranspose kernels
(kh,kw,Cout,Cin), Dense kernels (in,out).  Weights are an ordered dict  name -> np.ndarray  with
Keras auto-names (conv2d_1/kernel, batch_normalization_1/gamma, ...).
"""
from collections import OrderedDict
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import philox

BN_EPS = 1e-3          # keras BatchNormalization default epsilon
BN_MOMENTUM = 0.99     # keras BatchNormalization default momentum
K_EPSILON = 1e-7       # keras.backend.epsilon()


# --------------------------------------------------------------------------------------------
# initialisers (Keras VarianceScaling semantics; only the distribution can match TF, not the stream)
# --------------------------------------------------------------------------------------------
def he_normal(rng, shape, fan_in):
    std = math.sqrt(2.0 / fan_in) / 0.87962566103423978
    out = rng.standard_normal(shape)
    bad = np.abs(out) > 2.0
    while bad.any():                      # truncated normal at +-2 sigma (resample)
        out[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(out) > 2.0
    return (out * std).astype(np.float32)


def glorot_uniform(rng, shape, fan_in, fan_out):
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


# --------------------------------------------------------------------------------------------
# a tiny tape: the three graphs are written once as python functions over a `Ctx`
# that either *creates* parameters (init) or *consumes* them (forward)
# --------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self, params=None, rng=None, training=False, dtype=torch.float64,
                 dropout=None, taps=None):
        self.creating = params is None
        self.params = OrderedDict() if params is None else params
        self.rng = rng
        self.training = training
        self.dtype = dtype
        self.dropout = dropout          # None (off) or dict(seed=, step=)
        self.counts = {}
        self.new_moving = OrderedDict()  # BN moving-stat updates produced by a training forward
        self.taps = taps                # optional dict: layer name -> activation (NHWC numpy)
        self.shapes = []                # (layer name, output shape NHWC) like model.summary()
        self.n_dropout = 0

    def name(self, kind):
        self.counts[kind] = self.counts.get(kind, 0) + 1
        return "%s_%d" % (kind, self.counts[kind])

    def get(self, name, make):
        if self.creating and name not in self.params:
            self.params[name] = make()
        p = self.params[name]
        if isinstance(p, np.ndarray):
            p = torch.from_numpy(p)
        return p.to(self.dtype)

    def record(self, name, t):
        shp = tuple(t.shape)
        if len(shp) == 4:
            shp = (shp[0], shp[2], shp[3], shp[1])
        self.shapes.append((name, shp))
        if self.taps is not None:
            a = t.detach()
            self.taps[name] = (a.permute(0, 2, 3, 1) if a.dim() == 4 else a).contiguous().numpy()


def _act(x, act):
    if act == "relu":
        return torch.relu(x)
    if act == "elu":
        return F.elu(x, alpha=1.0)
    if act == "sigmoid":
        return torch.sigmoid(x)
    assert act in (None, "linear")
    return x


def conv2d(ctx, x, cout, k, act, init="he_normal"):
    """keras Conv2D(cout,(k,k),activation=act,padding='same'); x is NCHW inside the oracle."""
    name = ctx.name("conv2d")
    cin = x.shape[1]
    fan_in, fan_out = k * k * cin, k * k * cout
    mk = (lambda: he_normal(ctx.rng, (k, k, cin, cout), fan_in)) if init == "he_normal" else \
         (lambda: glorot_uniform(ctx.rng, (k, k, cin, cout), fan_in, fan_out))
    w = ctx.get(name + "/kernel", mk)
    b = ctx.get(name + "/bias", lambda: np.zeros(cout, np.float32))
    y = F.conv2d(x, w.permute(3, 2, 0, 1), b, padding=k // 2)      # HWIO -> OIHW, cross-correlation
    y = _act(y, act)
    ctx.record(name, y)
    return y


def conv2d_transpose(ctx, x, cout):
    """keras Conv2DTranspose(cout,(2,2),strides=(2,2),padding='same'), glorot_uniform, linear."""
    name = ctx.name("conv2d_transpose")
    cin = x.shape[1]
    # keras fan computation for a (kh,kw,Cout,Cin) kernel uses shape[-2] as fan_in, shape[-1] as fan_out
    fan_in, fan_out = 4 * cout, 4 * cin
    w = ctx.get(name + "/kernel", lambda: glorot_uniform(ctx.rng, (2, 2, cout, cin), fan_in, fan_out))
    b = ctx.get(name + "/bias", lambda: np.zeros(cout, np.float32))
    # out[n,co,2i+a,2j+b] = sum_ci x[n,ci,i,j] * W[a,b,co,ci]  -> torch weight (Cin,Cout,kh,kw)
    y = F.conv_transpose2d(x, w.permute(3, 2, 0, 1), b, stride=2)
    ctx.record(name, y)
    return y


def batchnorm(ctx, x):
    """keras BatchNormalization() defaults: axis=-1, momentum .99, eps 1e-3."""
    name = ctx.name("batch_normalization")
    c = x.shape[1]
    g = ctx.get(name + "/gamma", lambda: np.ones(c, np.float32))
    bta = ctx.get(name + "/beta", lambda: np.zeros(c, np.float32))
    mm = ctx.get(name + "/moving_mean", lambda: np.zeros(c, np.float32))
    mv = ctx.get(name + "/moving_variance", lambda: np.ones(c, np.float32))
    shp = (1, c) + (1,) * (x.dim() - 2)
    if ctx.training:
        red = [d for d in range(x.dim()) if d != 1]
        mean = x.mean(dim=red)
        var = x.var(dim=red, unbiased=False)
        n = x.numel() // c
        # Keras 2.3 normalization.py: variance *= n / (n - (1 + eps)) before the moving update
        unb = var.detach() * (n / (n - (1.0 + BN_EPS)))
        ctx.new_moving[name + "/moving_mean"] = (BN_MOMENTUM * mm + (1 - BN_MOMENTUM) * mean.detach())
        ctx.new_moving[name + "/moving_variance"] = (BN_MOMENTUM * mv + (1 - BN_MOMENTUM) * unb)
    else:
        mean, var = mm, mv
    y = (x - mean.view(shp)) / torch.sqrt(var.view(shp) + BN_EPS) * g.view(shp) + bta.view(shp)
    ctx.record(name, y)
    return y


def maxpool(ctx, x):
    name = ctx.name("max_pooling2d")
    y = F.max_pool2d(x, 2, 2)
    ctx.record(name, y)
    return y


def dropout(ctx, x, p):
    name = ctx.name("dropout")
    op_id = ctx.n_dropout
    ctx.n_dropout += 1
    if ctx.training and ctx.dropout is not None:
        shp = x.shape
        if x.dim() == 4:
            n, c, h, w = shp
            keep = philox.dropout_keep_mask(n * h * w * c, p, ctx.dropout["seed"], ctx.dropout["step"], op_id)
            keep = torch.from_numpy(keep.reshape(n, h, w, c)).permute(0, 3, 1, 2)
        else:
            keep = torch.from_numpy(philox.dropout_keep_mask(x.numel(), p, ctx.dropout["seed"],
                                                             ctx.dropout["step"], op_id).reshape(shp))
        x = x * keep.to(x.dtype) * (1.0 / (1.0 - p))
    ctx.record(name, x)
    return x


def concatenate(ctx, xs):
    name = ctx.name("concatenate")
    y = torch.cat(xs, dim=1)
    ctx.record(name, y)
    return y


def flatten(ctx, x):
    name = ctx.name("flatten")
    y = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)      # NHWC order (h,w,c) like Keras
    ctx.record(name, y)
    return y


def dense(ctx, x, units, act):
    name = ctx.name("dense")
    fin = x.shape[1]
    w = ctx.get(name + "/kernel", lambda: glorot_uniform(ctx.rng, (fin, units), fin, units))
    b = ctx.get(name + "/bias", lambda: np.zeros(units, np.float32))
    y = _act(x @ w + b, act)
    ctx.record(name, y)
    return y


# --------------------------------------------------------------------------------------------
# the three graphs
# --------------------------------------------------------------------------------------------
def unet(ctx, x):
    """T1H:853-915. conv+ReLU, conv+ReLU, BN, pool, dropout(.25); decoder convT, concat, BN, conv, conv."""
    skips = []
    for ch in (32, 64, 128, 256):
        c = conv2d(ctx, x, ch, 3, "relu")
        c = conv2d(ctx, c, ch, 3, "relu")
        c = batchnorm(ctx, c)
        skips.append(c)
        x = dropout(ctx, maxpool(ctx, c), 0.25)
    c = conv2d(ctx, x, 512, 3, "relu")
    c = conv2d(ctx, c, 512, 3, "relu")
    for ch, skip in zip((256, 128, 64, 32), reversed(skips)):
        u = conv2d_transpose(ctx, c, ch)
        u = concatenate(ctx, [u, skip])
        u = batchnorm(ctx, u)
        c = conv2d(ctx, u, ch, 3, "relu")
        c = conv2d(ctx, c, ch, 3, "relu")
    return conv2d(ctx, c, 1, 1, "sigmoid", init="glorot_uniform")


def _upp_conv_block(ctx, x, ch):
    """UPP:862-869: conv+ELU, Dropout .4, BN, conv+ELU, Dropout .4, BN."""
    for _ in range(2):
        x = conv2d(ctx, x, ch, 3, "elu")
        x = dropout(ctx, x, 0.4)
        x = batchnorm(ctx, x)
    return x


def _upp_backbone(ctx, x, ch):
    """UPP:878-882: conv+ELU, Dropout .2, conv+ELU, BN (pool applied by caller)."""
    c = conv2d(ctx, x, ch, 3, "elu")
    c = dropout(ctx, c, 0.2)
    c = conv2d(ctx, c, ch, 3, "elu")
    return batchnorm(ctx, c)


def unetpp(ctx, x):
    """UPP:875-949, layer creation order preserved (it fixes the Keras auto-names)."""
    c1 = _upp_backbone(ctx, x, 32); p1 = maxpool(ctx, c1)
    c2 = _upp_backbone(ctx, p1, 64); p2 = maxpool(ctx, c2)
    up1_2 = conv2d_transpose(ctx, c2, 32)
    conv1_2 = _upp_conv_block(ctx, concatenate(ctx, [up1_2, c1]), 32)
    c3 = _upp_backbone(ctx, p2, 128); p3 = maxpool(ctx, c3)
    up2_2 = conv2d_transpose(ctx, c3, 64)
    conv2_2 = _upp_conv_block(ctx, concatenate(ctx, [up2_2, c2]), 64)
    up1_3 = conv2d_transpose(ctx, conv2_2, 32)
    conv1_3 = _upp_conv_block(ctx, concatenate(ctx, [up1_3, c1, conv1_2]), 32)
    c4 = _upp_backbone(ctx, p3, 256); _p4 = maxpool(ctx, c4)      # p4 is dead in the reference (UPP:912)
    up3_2 = conv2d_transpose(ctx, c4, 128)
    conv3_2 = _upp_conv_block(ctx, concatenate(ctx, [up3_2, c3]), 128)
    up2_3 = conv2d_transpose(ctx, conv3_2, 64)
    conv2_3 = _upp_conv_block(ctx, concatenate(ctx, [up2_3, c2, conv2_2]), 64)
    up1_4 = conv2d_transpose(ctx, conv2_3, 32)
    conv1_4 = _upp_conv_block(ctx, concatenate(ctx, [up1_4, c1, conv1_2, conv1_3]), 32)
    return conv2d(ctx, conv1_4, 1, 1, "sigmoid", init="he_normal")


def classifier(ctx, x):
    """T2:747-778: [conv+ReLU, BN]x2, pool  x3 (16,32,64); Flatten; Dense32 ReLU; Dropout .4; Dense1 sigmoid."""
    for ch in (16, 32, 64):
        x = batchnorm(ctx, conv2d(ctx, x, ch, 3, "relu"))
        x = batchnorm(ctx, conv2d(ctx, x, ch, 3, "relu"))
        x = maxpool(ctx, x)
    x = flatten(ctx, x)
    x = dense(ctx, x, 32, "relu")
    x = dropout(ctx, x, 0.4)
    return dense(ctx, x, 1, "sigmoid")


GRAPHS = {"unet": unet, "unetpp": unetpp, "classifier": classifier}


# --------------------------------------------------------------------------------------------
# public helpers
# --------------------------------------------------------------------------------------------
def _to_nchw(x, dtype):
    t = torch.as_tensor(np.asarray(x)).to(dtype)
    return t.permute(0, 3, 1, 2).contiguous() if t.dim() == 4 else t


def _to_nhwc_np(t):
    t = t.detach()
    return (t.permute(0, 2, 3, 1) if t.dim() == 4 else t).contiguous().numpy()


def init_params(graph, input_shape, seed=42):
    """Create the ordered weight dict (Keras get_weights() order within each layer)."""
    ctx = Ctx(params=None, rng=np.random.default_rng(seed), training=False, dtype=torch.float32)
    h, w, c = input_shape
    with torch.no_grad():
        GRAPHS[graph](ctx, torch.zeros(1, c, h, w))
    return ctx.params, ctx.shapes


def is_trainable(name):
    return not (name.endswith("moving_mean") or name.endswith("moving_variance"))


def forward(graph, params, x, training=False, dtype=torch.float64, dropout=None, taps=None):
    """model.predict (training=False) or the training-mode forward. Returns (probs NHWC numpy, ctx)."""
    ctx = Ctx(params=params, training=training, dtype=dtype, dropout=dropout, taps=taps)
    with torch.no_grad():
        y = GRAPHS[graph](ctx, _to_nchw(x, dtype))
    return _to_nhwc_np(y), ctx


def binary_crossentropy_map(t, p):
    """keras.losses.binary_crossentropy: clip to [eps, 1-eps], mean over the last (channel) axis."""
    ph = torch.clamp(p, K_EPSILON, 1.0 - K_EPSILON)
    return -(t * torch.log(ph) + (1.0 - t) * torch.log(1.0 - ph))


def dice_coeff(t, p):
    """T1H:784-790 -- over the whole batch flattened, smooth = 1."""
    inter = (t * p).sum()
    return (2.0 * inter + 1.0) / (t.sum() + p.sum() + 1.0)


def bce_dice_loss(t, p):
    """T1H:797-799 after Keras' mean reduction: 0.5*mean(BCE) + 0.5*(1-dice)."""
    return 0.5 * binary_crossentropy_map(t, p).mean() + 0.5 * (1.0 - dice_coeff(t, p))


def weighted_bce(t, p, sample_w):
    """T2:828/836: 'binary_crossentropy' with class_weight -> per-sample weights, Keras weighted mean
    (sum(w*l)/N, Keras 2.3 'sum_over_batch_size' on the weighted losses)."""
    l = binary_crossentropy_map(t, p).mean(dim=-1)
    return (l * sample_w).mean()


def loss_and_grads(graph, params, x, t, dtype=torch.float64, dropout=None, loss="bce_dice",
                   sample_weight=None, taps=None):
    """One training-mode forward + backward. Returns dict(loss, metric, grads{name}, probs, new_moving)."""
    tp = OrderedDict()
    for k, v in params.items():
        tv = torch.from_numpy(np.asarray(v)).to(dtype)
        tp[k] = tv.requires_grad_(is_trainable(k))
    ctx = Ctx(params=tp, training=True, dtype=dtype, dropout=dropout, taps=taps)
    p = GRAPHS[graph](ctx, _to_nchw(x, dtype))
    tt = _to_nchw(t, dtype)
    if loss == "bce_dice":
        L = bce_dice_loss(tt, p)
        metric = dice_coeff(tt, p)
    else:
        sw = torch.ones(p.shape[0], dtype=dtype) if sample_weight is None else torch.as_tensor(sample_weight).to(dtype)
        L = weighted_bce(tt, p, sw)
        metric = L
    names = [k for k in tp if is_trainable(k)]
    gs = torch.autograd.grad(L, [tp[k] for k in names])
    grads = OrderedDict((k, g.detach().numpy()) for k, g in zip(names, gs))
    new_moving = OrderedDict((k, v.detach().numpy()) for k, v in ctx.new_moving.items())
    return dict(loss=float(L.detach()), metric=float(metric.detach()), grads=grads, probs=_to_nhwc_np(p),
                new_moving=new_moving, ctx=ctx)


class Adam:
    """keras.optimizers.Adam(lr) defaults b1 .9, b2 .999, eps 1e-7 (T1H:1053):
    p -= lr*sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps)."""

    def __init__(self, lr=5e-4, b1=0.9, b2=0.999, eps=K_EPSILON):
        self.lr, self.b1, self.b2, self.eps, self.t = lr, b1, b2, eps, 0
        self.m, self.v = {}, {}

    def step(self, params, grads):
        self.t += 1
        lr_t = self.lr * math.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        for k, g in grads.items():
            g = np.asarray(g, dtype=np.float64)
            m = self.m.get(k, 0.0) * self.b1 + (1 - self.b1) * g
            v = self.v.get(k, 0.0) * self.b2 + (1 - self.b2) * g * g
            self.m[k], self.v[k] = m, v
            params[k] = (np.asarray(params[k], np.float64) - lr_t * m / (np.sqrt(v) + self.eps)).astype(params[k].dtype)
        return params


def train_step(graph, params, opt, x, t, dtype=torch.float32, dropout=None, **kw):
    """forward + backward + Adam + BN moving update, in place on `params`; returns (loss, metric)."""
    r = loss_and_grads(graph, params, x, t, dtype=dtype, dropout=dropout, **kw)
    opt.step(params, r["grads"])
    for k, v in r["new_moving"].items():
        params[k] = v.astype(params[k].dtype)
    return r["loss"], r["metric"]


def cosine_annealing_lr(epoch, T_max=7, eta_max=5e-4, eta_min=1e-4):
    """T1H:980 CosineAnnealingScheduler.on_epoch_begin."""
    return eta_min + (eta_max - eta_min) * (1 + math.cos(math.pi * epoch / T_max)) / 2


def sm_threshold_metrics(t, p, threshold, smooth=1e-5):
    """segmentation_models 1.0.1 FScore/IOUScore/Precision/Recall(threshold) on one batch
    (T1H:1206-1211): pr = float(p > thr); gt not thresholded; reduced over the whole batch."""
    t = np.asarray(t, np.float64)
    pr = (np.asarray(p, np.float64) > threshold).astype(np.float64)
    tp = (t * pr).sum()
    fp = pr.sum() - tp
    fn = t.sum() - tp
    return dict(f1=(2 * tp + smooth) / (2 * tp + fn + fp + smooth),
                iou=(tp + smooth) / (t.sum() + pr.sum() - tp + smooth),
                precision=(tp + smooth) / (tp + fp + smooth),
                recall=(tp + smooth) / (tp + fn + smooth),
                tp=tp, sum_pr=pr.sum(), sum_gt=t.sum())


def task2_batch_metrics(t, p):
    """T2:688-703 recall / precision / f1 with K.round(K.clip(.,0,1)) (round-half-even) and eps 1e-7."""
    t = np.asarray(t, np.float64); p = np.asarray(p, np.float64)
    tp = np.rint(np.clip(t * p, 0, 1)).sum()
    poss = np.rint(np.clip(t, 0, 1)).sum()
    pred = np.rint(np.clip(p, 0, 1)).sum()
    rec = tp / (poss + K_EPSILON)
    prec = tp / (pred + K_EPSILON)
    return dict(recall=rec, precision=prec, f1=2 * (prec * rec) / (prec + rec + K_EPSILON))


def count_params(params):
    tot = sum(int(np.prod(v.shape)) for v in params.values())
    tr = sum(int(np.prod(v.shape)) for k, v in params.items() if is_trainable(k))
    return tot, tr, tot - tr
