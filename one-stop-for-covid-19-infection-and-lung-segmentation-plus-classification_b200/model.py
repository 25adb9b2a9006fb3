"""Keras-shaped facade: the subset of `keras.models.Model` the reference runners call.

  compile / fit / evaluate / predict / predict_proba / save_weights / load_weights / to_json /
  summary / get_layer / optimizer.lr / callbacks (on_epoch_begin, on_epoch_end)

Call sites replaced: /root/reference/Scripts/task1_preprocessing_plus_unet_with_comments.py:1053
(compile), :1059-1061 (fit), :1101 (evaluate), :1137 (predict), :1073-1093 (weights / json),
:1196-1343 (threshold sweeps); task2_covid19_classifcation.py:726-728 (predict_proba), :828-836.

Inputs are host numpy arrays (NHWC, float64/float32 in [0,1]) exactly as the reference passes them;
outputs are fresh host arrays.  All arithmetic runs in libb200unet.so on the GPU.
"""
import json
import math
import time

import numpy as np
import torch

from . import dist as D
from . import engine as E
from . import hdf5 as H5
from . import layers as L
from . import losses as LS
from . import plan as P


class Adam:
    """keras.optimizers.Adam(lr=...) -- T1H:1053"""

    def __init__(self, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, learning_rate=None, **kw):
        self.lr = float(learning_rate if learning_rate is not None else lr)
        self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon


class _OptimizerHandle:
    """`model.optimizer.lr` get/set (K.set_value / K.get_value in the reference, T1H:982-990)."""

    def __init__(self, eng):
        self._eng = eng

    @property
    def lr(self):
        return self._eng.lr

    @lr.setter
    def lr(self, v):
        self._eng.lr = v


class History:
    def __init__(self):
        self.history = {}
        self.epoch = []


class Callback:
    def set_model(self, model):
        self.model = model

    def on_train_begin(self, logs=None):
        pass

    def on_train_end(self, logs=None):
        pass

    def on_epoch_begin(self, epoch, logs=None):
        pass

    def on_epoch_end(self, epoch, logs=None):
        pass


class CosineAnnealingScheduler(Callback):
    """T1H:970-990: lr(e) = eta_min + (eta_max - eta_min) * (1 + cos(pi * e / T_max)) / 2 at epoch begin."""

    def __init__(self, T_max, eta_max, eta_min=0, verbose=1):
        self.T_max, self.eta_max, self.eta_min, self.verbose = T_max, eta_max, eta_min, verbose

    def on_epoch_begin(self, epoch, logs=None):
        lr = self.eta_min + (self.eta_max - self.eta_min) * (1 + math.cos(math.pi * epoch / self.T_max)) / 2
        self.model.optimizer.lr = lr
        if self.verbose > 0:
            print('\nEpoch %05d: CosineAnnealingScheduler setting learning rate to %s.' % (epoch + 1, lr))

    def on_epoch_end(self, epoch, logs=None):
        if logs is not None:
            logs['lr'] = self.model.optimizer.lr


class PendingBatch:
    """result of Model.train_on_batch(..., wait=False): `.get()` blocks until that step has finished and returns
    [loss, metric] from the pinned host buffer the step wrote them to"""

    def __init__(self, event, out):
        self._event, self._out, self._val = event, out, None

    def get(self):
        if self._val is None:
            self._event.synchronize()
            self._val = [float(self._out[0]), float(self._out[1])]
        return self._val


class ModelCheckpoint(Callback):
    """keras ModelCheckpoint(filepath, monitor, save_best_only, mode) -- T1H:1044-1047 (weights as .npz)."""

    def __init__(self, filepath, monitor='val_loss', verbose=0, save_best_only=False, mode='auto', **kw):
        self.filepath, self.monitor, self.verbose, self.save_best_only = filepath, monitor, verbose, save_best_only
        if mode == 'auto':
            mode = 'max' if ('acc' in monitor or 'dice' in monitor or 'auc' in monitor or monitor.startswith('val_f')) else 'min'
        self.mode = mode
        self.best = -np.inf if mode == 'max' else np.inf

    def on_epoch_end(self, epoch, logs=None):
        cur = (logs or {}).get(self.monitor)
        if not self.save_best_only:
            self.model.save_weights(self.filepath)
            return
        if cur is None:
            return
        better = cur > self.best if self.mode == 'max' else cur < self.best
        if better:
            if self.verbose:
                print('Epoch %05d: %s improved from %0.5f to %0.5f, saving model to %s'
                      % (epoch + 1, self.monitor, self.best, cur, self.filepath))
            self.best = cur
            self.model.save_weights(self.filepath)


class RocCallback(Callback):
    """T2:706-741: ROC-AUC on train and validation after every epoch; saves weights on best val AUC."""

    def __init__(self, training_data, validation_data, filepath=None):
        self.x, self.y = training_data
        self.x_val, self.y_val = validation_data
        self.filepath = filepath
        self.best = -1.0

    def on_epoch_end(self, epoch, logs=None):
        from sklearn.metrics import roc_auc_score
        roc = roc_auc_score(self.y, self.model.predict_proba(self.x))
        roc_val = roc_auc_score(self.y_val, self.model.predict_proba(self.x_val))
        if logs is not None:
            logs['roc_auc'], logs['val_roc_auc'] = roc, roc_val
        print('\rroc-auc: %s - roc-auc_val: %s' % (str(round(roc, 5)), str(round(roc_val, 5))), end=100 * ' ' + '\n')
        if roc_val > self.best and self.filepath:
            self.best = roc_val
            self.model.save_weights(self.filepath)


def _metric_name(m):
    return m if isinstance(m, str) else getattr(m, "__name__", m.__class__.__name__)


class Model:
    """`Model(inputs=[inputs], outputs=[outputs])` (T1H:915) over a layers.Graph."""

    def __init__(self, inputs=None, outputs=None, graph=None, precision="float16", seed=42, comm=None,
                 sync_stats=False, use_graph=True, device=None, dropout_seed=7, loss_scale=None, plan_options=None):
        self.graph = graph if graph is not None else L.Graph(inputs, outputs)
        self._eng_kw = dict(precision=precision, seed=seed, comm=comm, sync_stats=sync_stats, use_graph=use_graph,
                            device=device, dropout_seed=dropout_seed, loss_scale=loss_scale, plan_options=plan_options)
        self._eng = None
        self.loss_kind = "bce_dice" if len(self.graph.output.shape) == 3 else "bce"
        self.metrics = []
        self.metrics_names = ["loss"]
        self.stop_training = False
        self.pipeline_predict = True     # predict(): overlap the host->device copy of a host input with the forward passes
        self._shuffle_rng = np.random.RandomState(1234)

    # ---- structure --------------------------------------------------------------------------
    @property
    def layers(self):
        return self.graph.layers

    @property
    def input(self):
        return self.graph.input

    def get_layer(self, name):
        return self.graph.get_layer(name)

    def count_params(self):
        return self.graph.count_params()[0]

    def summary(self, print_fn=print):
        print_fn('Model: "model"')
        print_fn("_" * 98)
        print_fn("%-32s %-26s %-10s %s" % ("Layer (type)", "Output Shape", "Param #", "Connected to"))
        print_fn("=" * 98)
        for l in self.graph.layers:
            shp = "(None, " + ", ".join(str(s) for s in l.output.shape) + ")"
            conn = ", ".join(t.producer.name for t in l.inputs)
            print_fn("%-32s %-26s %-10d %s" % ("%s (%s)" % (l.name, type(l).__name__), shp, l.count_params(), conn))
        tot, tr, ntr = self.graph.count_params()
        print_fn("=" * 98)
        print_fn("Total params: {:,}".format(tot))
        print_fn("Trainable params: {:,}".format(tr))
        print_fn("Non-trainable params: {:,}".format(ntr))

    def to_json(self):
        """keras `model.to_json()` (T1H:1091, UPP:1113): the functional-API architecture description Keras 2.3 writes --
        class_name Model, config.layers[{name, class_name, config, inbound_nodes}], input_layers, output_layers --
        so `keras.models.model_from_json` can rebuild the network, and `model_from_json` below reads Keras' own files."""
        return json.dumps(keras_config(self.graph))

    # ---- engine ------------------------------------------------------------------------------
    @property
    def engine(self):
        if self._eng is None:
            self._eng = E.Engine(self.graph, loss=self.loss_kind, **self._eng_kw)
        return self._eng

    @property
    def optimizer(self):
        return _OptimizerHandle(self.engine)

    def compile(self, optimizer=None, loss=None, metrics=None, **kw):
        """Fresh optimizer state, weights kept (the reference re-compiles per fold / per threshold)."""
        name = _metric_name(loss) if loss is not None else None
        if name in ("bce_dice_loss",):
            kind = "bce_dice"
        elif name in ("binary_crossentropy",):
            kind = "bce" if len(self.graph.output.shape) == 1 else "bce_dice_unsupported"
        elif name is None:
            kind = self.loss_kind
        else:
            raise ValueError("unsupported loss %r (engine implements bce_dice_loss and binary_crossentropy)" % name)
        if kind == "bce_dice_unsupported":
            raise ValueError("binary_crossentropy alone is implemented for the Dense(1) classifier head only")
        if self._eng is not None and kind != self.loss_kind:
            self._eng.loss = kind
        self.loss_kind = kind
        self.metrics = list(metrics or [])
        self.metrics_names = ["loss"] + [_metric_name(m) for m in self.metrics]
        opt = optimizer if optimizer is not None else Adam()
        if isinstance(opt, str):
            opt = Adam()
        self.engine.reset_optimizer(lr=opt.lr, beta1=opt.beta_1, beta2=opt.beta_2, eps=opt.epsilon)

    # ---- weights -----------------------------------------------------------------------------
    def get_weights(self):
        return list(self.engine.get_weights().values())

    def set_weights(self, ws):
        names = [n for n, _, _, _ in self.graph.weight_specs()]
        self.engine.set_weights(dict(zip(names, ws)))

    def get_weights_dict(self):
        return self.engine.get_weights()

    def set_weights_dict(self, d):
        self.engine.set_weights(d)

    def _keras_layers(self, d):
        """[(layer name, [(weight name 'layer/kernel:0', array), ...])] in model.layers order (Keras' HDF5 layout)"""
        out = []
        for l in self.graph.layers:
            out.append((l.name, [("%s/%s:0" % (l.name, k), d["%s/%s" % (l.name, k)]) for k in l.weights]))
        return out

    def save_weights(self, path):
        """keras `model.save_weights(path)` (T1H:1079, CV4:1105-1108): '.h5' / '.hdf5' / '.keras' paths are written as Keras
        2.3 HDF5 weight files (hdf5.py: layer_names / weight_names attributes, <layer>/<layer>/kernel:0 datasets in Keras
        layouts), anything else as a numpy .npz archive."""
        if self.engine.rank != 0:           # data parallel: replicas are identical, rank 0 owns the file
            return
        d = self.engine.get_weights()
        if str(path).lower().endswith((".h5", ".hdf5", ".keras")):
            H5.save_keras_weights(path, self._keras_layers(d))
            return
        with open(path, "wb") as f:
            np.savez(f, **{k.replace("/", "__"): v for k, v in d.items()})

    def load_weights(self, path):
        """keras `model.load_weights(path)` (T1H:1073, 1190): Keras HDF5 weight files -- also the full-model files
        ModelCheckpoint writes (weights under /model_weights, T1H:1044-1047) -- or the .npz archives of save_weights.
        Layers are matched by name like Keras' topological loading does for identical architectures; shapes are checked."""
        D.barrier(self.engine.world)        # data parallel: rank 0 may still be writing the file
        with open(path, "rb") as f:
            magic = f.read(8)
        if magic == H5.SIGNATURE or str(path).lower().endswith((".h5", ".hdf5")):
            layers = H5.load_keras_weights(path)
            d = {}
            for l in self.graph.layers:
                if not l.weights:
                    continue
                if l.name not in layers:
                    raise ValueError("weight file %s has no layer %r" % (path, l.name))
                vals = list(layers[l.name].values())
                if len(vals) != len(l.weights):
                    raise ValueError("layer %s: file has %d weight tensors, the model %d" % (l.name, len(vals), len(l.weights)))
                for k, v in zip(l.weights, vals):          # Keras order within a layer: kernel, bias / gamma, beta, mean, var
                    d["%s/%s" % (l.name, k)] = v
            self.engine.set_weights(d)
            return
        with np.load(path) as z:
            self.engine.set_weights({k.replace("__", "/"): z[k] for k in z.files})

    # ---- data --------------------------------------------------------------------------------
    def _to_dev(self, a, flat_out=False):
        eng = self.engine
        if isinstance(a, torch.Tensor):           # e.g. a pinned host batch: asynchronous copy on the engine's stream
            with torch.cuda.stream(eng.stream):
                return a.to(eng.device, dtype=torch.float32, non_blocking=True).contiguous()
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
        with torch.cuda.stream(eng.stream):
            return torch.from_numpy(a).to(eng.device, non_blocking=False)

    def _check_x(self, x):
        if not isinstance(x, torch.Tensor):
            x = np.asarray(x)
        if tuple(x.shape[1:]) != tuple(self.graph.input.shape):
            raise ValueError("expected input of shape (N,%s), got %s" % (",".join(map(str, self.graph.input.shape)), x.shape))
        return x

    def _sample_weights(self, y, class_weight):
        if class_weight is None:
            return np.ones(len(y), np.float32)
        yl = np.asarray(y).reshape(len(y), -1)[:, 0]
        if isinstance(class_weight, dict):
            return np.asarray([class_weight[int(round(v))] for v in yl], np.float32)
        cw = np.asarray(class_weight, np.float32)           # the reference passes an ndarray (T2:801-803, 836)
        return cw[np.rint(yl).astype(np.int64)]

    # ---- metric plumbing -----------------------------------------------------------------------
    def _host_metrics(self, metrics, t, p):
        out = []
        for m in metrics:
            if isinstance(m, str):
                if m in ("accuracy", "acc"):
                    out.append(float((np.rint(p) == np.rint(t)).mean()))
                else:
                    raise ValueError("unknown metric %r" % m)
            else:
                out.append(float(m(t, p)))
        return out

    def _run_eval(self, x, y, batch_size, metrics, class_weight=None):
        """forward (inference mode) over x in batches; returns [loss, metrics...] with Keras'
        batch-size-weighted averaging of per-batch values."""
        eng = self.engine
        x = self._check_x(x)
        n_tot = len(x)
        xd, yd = self._to_dev(x), self._to_dev(np.asarray(y).reshape(n_tot, -1))
        swd = self._to_dev(self._sample_weights(y, class_weight)) if self.loss_kind == "bce" else None
        sm = [m for m in metrics if isinstance(m, LS._SMMetric)]
        dice_idx = [k for k, m in enumerate(metrics) if _metric_name(m) == "dice_coeff"]
        host_ms = [m for m in metrics if not isinstance(m, LS._SMMetric) and _metric_name(m) != "dice_coeff"]
        nb = (n_tot + batch_size - 1) // batch_size
        with torch.cuda.stream(eng.stream):
            lossbuf = torch.zeros(nb, 2, dtype=torch.float32, device=eng.device)
            thr = torch.tensor([m.threshold for m in sm] or [0.5], dtype=torch.float32, device=eng.device)
            tp = torch.zeros(nb, thr.numel(), dtype=torch.float64, device=eng.device)
            spr = torch.zeros_like(tp)
            sgt = torch.zeros(nb, dtype=torch.float64, device=eng.device)
            probs_all = torch.empty((n_tot,) + tuple(self.graph.output.shape), dtype=torch.float32,
                                    device=eng.device) if host_ms else None
            sizes = []
            for bi in range(nb):
                lo = bi * batch_size
                n = min(batch_size, n_tot - lo)
                sizes.append(n)
                b = eng.forward_batch(xd[lo:lo + n], None, n, t_src=yd[lo:lo + n],
                                      sw_src=swd[lo:lo + n] if swd is not None else None)
                lossbuf[bi].copy_(eng.loss_dev(b))
                if sm:
                    eng.threshold_counts(b, thr, tp[bi], spr[bi], sgt[bi:bi + 1])
                if probs_all is not None:
                    probs_all[lo:lo + n].copy_(eng.probs(b))
        eng.stream.synchronize()
        w = np.asarray(sizes, np.float64)
        lb = lossbuf.cpu().numpy().astype(np.float64)
        res = {"loss": float((lb[:, 0] * w).sum() / w.sum())}
        vals = []
        tpn, sprn, sgtn = tp.cpu().numpy(), spr.cpu().numpy(), sgt.cpu().numpy()
        hm = None
        if host_ms:
            pa = probs_all.cpu().numpy()
            ya = np.asarray(y, np.float64).reshape(pa.shape)
            hm = np.zeros((nb, len(host_ms)))
            for bi in range(nb):
                lo = bi * batch_size
                hm[bi] = self._host_metrics(host_ms, ya[lo:lo + sizes[bi]], pa[lo:lo + sizes[bi]])
        ks = kh = 0
        for k, m in enumerate(metrics):
            if k in dice_idx:
                vals.append(float((lb[:, 1] * w).sum() / w.sum()))
            elif isinstance(m, LS._SMMetric):
                per = np.array([m.from_counts(tpn[bi, ks], sprn[bi, ks], sgtn[bi]) for bi in range(nb)])
                vals.append(float((per * w).sum() / w.sum()))
                ks += 1
            else:
                vals.append(float((hm[:, kh] * w).sum() / w.sum()))
                kh += 1
        return [res["loss"]] + vals

    # ---- public API ----------------------------------------------------------------------------
    def evaluate(self, x, y, batch_size=32, verbose=0, **kw):
        out = self._run_eval(x, y, batch_size, self.metrics)
        return out if len(out) > 1 else out[0]

    def predict(self, x, batch_size=32, verbose=0, **kw):
        """keras Model.predict (T1H:1113, T2:726).  A HOST input is copied in chunks on a separate copy stream while the
        forward of the previous chunk runs (inference is per-sample independent, so the chunk size only changes speed:
        `batch_size`, or a quarter of the input when the whole input is less than four batches); the probabilities come
        back in one device->host copy at the end."""
        eng = self.engine
        x = self._check_x(x)
        n_tot = len(x)
        on_dev = isinstance(x, torch.Tensor) and x.is_cuda
        chunk = int(batch_size)
        if not on_dev and self.pipeline_predict and n_tot < 4 * chunk:
            chunk = max(1, min(chunk, -(-n_tot // 4)))
        with torch.cuda.stream(eng.stream):
            out = torch.empty((n_tot,) + tuple(self.graph.output.shape), dtype=torch.float32, device=eng.device)
        if on_dev or not self.pipeline_predict or n_tot <= chunk:
            xd, xt, cs = self._to_dev(x), None, None
        else:
            xt = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
            cs = getattr(self, "_predict_copy_stream", None)
            if cs is None:
                cs = self._predict_copy_stream = torch.cuda.Stream(eng.device)
            with torch.cuda.stream(eng.stream):
                xd = torch.empty(tuple(xt.shape), dtype=torch.float32, device=eng.device)
            cs.wait_stream(eng.stream)                      # the buffer exists (and its previous user is done) before copying
            xd.record_stream(cs)
        for lo in range(0, n_tot, chunk):
            n = min(chunk, n_tot - lo)
            if cs is not None:
                # copy, then enqueue the forward: with pageable memory the copy blocks the host while the device still
                # runs the previous chunk; with pinned memory everything is queued at once and the two streams overlap
                with torch.cuda.stream(cs):
                    xd[lo:lo + n].copy_(xt[lo:lo + n], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(cs)
                eng.stream.wait_event(ev)
            with torch.cuda.stream(eng.stream):
                b = eng.forward_batch(xd[lo:lo + n], None, n)
                out[lo:lo + n].copy_(eng.probs(b))
        eng.stream.synchronize()
        return out.cpu().numpy()

    def predict_proba(self, x, batch_size=32, verbose=0):
        return self.predict(x, batch_size=batch_size)

    def intermediate(self, x, layer_name):
        """Model(inputs=model.input, outputs=model.get_layer(name).output).predict(x) -- T1H:1386-1405"""
        eng = self.engine
        x = self._check_x(x)
        outs = []
        xd = self._to_dev(x)
        for lo in range(0, len(x), 32):
            n = min(32, len(x) - lo)
            b = eng.forward_batch(xd[lo:lo + n], None, n, tap=True)
            outs.append(eng.layer_output(b, layer_name))
        return np.concatenate(outs, 0)

    def threshold_sweep(self, x, y, thresholds, batch_size=32):
        """All thresholds x {F1, IoU, precision, recall} from ONE forward pass (replaces the reference's
        re-compile + evaluate loop, T1H:1196-1343).  Values are Keras-style batch-weighted means."""
        ms = []
        for t in thresholds:
            ms += [LS.FScore(threshold=t), LS.IOUScore(threshold=t), LS.Precision(threshold=t), LS.Recall(threshold=t)]
        # one device threshold per metric object is wasteful but tiny; counts are shared by value
        vals = self._run_eval(x, y, batch_size, ms)[1:]
        v = np.asarray(vals).reshape(len(thresholds), 4)
        return {"threshold": np.asarray(thresholds), "f1": v[:, 0], "iou": v[:, 1], "precision": v[:, 2], "recall": v[:, 3]}

    def train_on_batch(self, x, y, sample_weight=None, dropout=True, wait=True):
        """keras Model.train_on_batch: one optimisation step on a HOST batch; returns [loss, metric].
        The host->device copy of the batch and the device->host read of the loss are part of the call
        (this is the end-to-end path bench.py times).  x / y may be numpy arrays or pinned torch tensors.

        The batch is copied on a separate copy stream into one of two staging slots, so with `wait=False` (the call
        then returns a `PendingBatch` whose `.get()` gives [loss, metric]) the copy of the next batch overlaps the
        training step of the current one: call, then `.get()` the handle of the PREVIOUS call."""
        eng = self.engine
        n = len(x)
        xt = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
        yt = y if isinstance(y, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(y, dtype=np.float32))
        key = (n, tuple(xt.shape[1:]))
        st = getattr(self, "_stage", None)
        if st is None or st[0] != key:
            eng.stream.synchronize()
            slots = []
            with torch.cuda.stream(eng.stream):
                for _ in range(2):
                    slots.append(dict(x=torch.empty(xt.shape, dtype=torch.float32, device=eng.device),
                                      y=torch.empty((n, int(yt.numel() // n)), dtype=torch.float32, device=eng.device),
                                      sw=torch.ones(n, dtype=torch.float32, device=eng.device),
                                      out=torch.empty(2, dtype=torch.float32).pin_memory(),
                                      copied=torch.cuda.Event(), consumed=None, done=torch.cuda.Event()))
            eng.stream.synchronize()
            st = self._stage = (key, slots, [0], torch.cuda.Stream(eng.device))
        _, slots, turn, copy_stream = st
        slot = slots[turn[0]]
        turn[0] ^= 1
        with torch.cuda.stream(copy_stream):
            if slot["consumed"] is not None:
                copy_stream.wait_event(slot["consumed"])        # the step that last used this slot has finished
            slot["x"].copy_(xt, non_blocking=True)
            slot["y"].copy_(yt.reshape(n, -1), non_blocking=True)
            if sample_weight is not None:
                slot["sw"].copy_(torch.as_tensor(sample_weight, dtype=torch.float32), non_blocking=True)
            slot["copied"].record(copy_stream)
        with torch.cuda.stream(eng.stream):
            eng.stream.wait_event(slot["copied"])
            b = eng.train_batch(slot["x"], slot["y"], None, n, dropout=dropout, sw_src=slot["sw"])
            slot["out"].copy_(eng.loss_dev(b), non_blocking=True)
            slot["done"].record(eng.stream)
            slot["consumed"] = slot["done"]
        pending = PendingBatch(slot["done"], slot["out"])
        return pending.get() if wait else pending

    def fit(self, x, y, batch_size=32, epochs=1, validation_data=None, callbacks=None, class_weight=None,
            shuffle=True, verbose=1, initial_epoch=0, dropout=True, **kw):
        eng = self.engine
        x = self._check_x(x)
        n_tot = len(x)
        world, rank = eng.world, eng.rank
        xd, yd = self._to_dev(x), self._to_dev(np.asarray(y).reshape(n_tot, -1))
        swd = self._to_dev(self._sample_weights(y, class_weight)) if self.loss_kind == "bce" else None
        hist = History()
        cbs = list(callbacks or [])
        for cb in cbs:
            cb.set_model(self)
            cb.on_train_begin({})
        metric_names = [_metric_name(m) for m in self.metrics]
        if verbose:
            if validation_data is not None:
                print("Train on %d samples, validate on %d samples" % (n_tot, len(validation_data[0])))
            else:
                print("Train on %d samples" % n_tot)
        for epoch in range(initial_epoch, epochs):
            if self.stop_training:
                break
            logs = {}
            for cb in cbs:
                cb.on_epoch_begin(epoch, logs)
            t0 = time.time()
            # the shuffle stream is seeded identically on every rank, so all ranks see the same permutation and take
            # disjoint, equally sized shares of it (dist.epoch_batches): same step count and batch sizes everywhere
            perm = self._shuffle_rng.permutation(n_tot) if shuffle else np.arange(n_tot)
            batches = D.epoch_batches(perm, batch_size, rank, world)
            with torch.cuda.stream(eng.stream):
                lossbuf = torch.zeros(len(batches), 2, dtype=torch.float32, device=eng.device)
                sizes = []
                lo = 0
                mine = np.concatenate(batches) if batches else np.zeros(0, np.int64)
                perm_d = torch.from_numpy(mine.astype(np.int32)).to(eng.device)
                for bi, idx in enumerate(batches):
                    n = len(idx)
                    sizes.append(n)
                    b = eng.train_batch(xd, yd, perm_d[lo:lo + n], n, dropout=dropout, sw_src=swd)
                    lossbuf[bi].copy_(eng.loss_dev(b))
                    lo += n
            eng.stream.synchronize()
            eng.guard_small_gamma()
            if eng.overflowed():
                print("warning: steps with non-finite gradients were skipped this epoch (loss scale now %g)" % eng.loss_scale)
                eng._set_fields(overflow=0)
            w = np.asarray(sizes, np.float64)
            lb = lossbuf.cpu().numpy().astype(np.float64)
            logs["loss"] = float((lb[:, 0] * w).sum() / w.sum())
            if "dice_coeff" in metric_names:
                logs["dice_coeff"] = float((lb[:, 1] * w).sum() / w.sum())
            if world > 1:                      # epoch means over all ranks (equal sample counts per rank)
                for k in ("loss", "dice_coeff"):
                    if k in logs:
                        logs[k] = D.mean_over_ranks(logs[k], world)
            if world > 1 and not eng.sync_stats:
                # BatchNorm moving statistics follow each rank's own batches: average them once per epoch so that validation,
                # checkpoints and every callback decision (best-only saves, early stopping) are the same on all ranks
                eng.average_moving_statistics()
            if validation_data is not None:
                xv, yv = validation_data[0], validation_data[1]
                vals = self._run_eval(xv, yv, batch_size, self.metrics)
                logs["val_loss"] = vals[0]
                for nm, v in zip(metric_names, vals[1:]):
                    logs["val_" + nm] = v
                if world > 1:                  # identical weights and statistics: the mean only removes last-bit noise
                    for k in list(logs):
                        if k.startswith("val_"):
                            logs[k] = D.mean_over_ranks(logs[k], world)
            dt_ = time.time() - t0
            for cb in cbs:
                cb.on_epoch_end(epoch, logs)
            hist.epoch.append(epoch)
            for k, v in logs.items():
                hist.history.setdefault(k, []).append(v)
            if verbose:
                msg = " - ".join("%s: %.4f" % (k, v) for k, v in logs.items())
                print("Epoch %d/%d\n%d/%d - %ds %dms/sample - %s" % (epoch + 1, epochs, n_tot, n_tot, int(dt_),
                                                                      int(1000 * dt_ / max(n_tot, 1)), msg))
        for cb in cbs:
            cb.on_train_end({})
        self.history = hist
        return hist


class Sequential(Model):
    """keras.models.Sequential: model.add(layer) chains (T2:747-778)."""

    def __init__(self, **kw):
        self._seq_layers, self._seq_out, self._kw = [], None, kw
        self._built = False
        L.reset_names()          # a fresh model: the layers created for the add() calls are conv2d_1, ... as in Keras

    def add(self, layer):
        if self._seq_out is None:
            shape = getattr(layer, "input_shape", None)
            if shape is None:
                raise ValueError("the first layer of a Sequential model needs input_shape=")
            self._seq_in = L.Input(shape)
            # the input layer is created after the first layer object: order it in front (Graph sorts by creation)
            self._seq_in.producer._seq = layer._seq - 0.5
            self._seq_out = layer(self._seq_in)
        else:
            self._seq_out = layer(self._seq_out)
        self._built = False

    def _ensure(self):
        if not self._built:
            Model.__init__(self, inputs=[self._seq_in], outputs=[self._seq_out], **self._kw)
            self._built = True

    def __getattr__(self, name):
        # first access to any Model attribute after add() finalises the graph
        if name in ("_seq_layers", "_seq_out", "_kw", "_built", "_seq_in"):
            raise AttributeError(name)
        if not self.__dict__.get("_built", False) and self.__dict__.get("_seq_out") is not None:
            self._ensure()
            return getattr(self, name)
        raise AttributeError(name)


# ---- Keras architecture JSON (model.to_json / model_from_json, T1H:1091) ------------------------------------------
_INIT = {"he_normal": {"class_name": "VarianceScaling", "config": {"scale": 2.0, "mode": "fan_in", "distribution": "normal", "seed": None}},
         "glorot_uniform": {"class_name": "VarianceScaling", "config": {"scale": 1.0, "mode": "fan_avg", "distribution": "uniform", "seed": None}}}
_ZEROS, _ONES = {"class_name": "Zeros", "config": {}}, {"class_name": "Ones", "config": {}}


def _layer_config(l):
    base = {"name": l.name, "trainable": True, "dtype": "float32"}
    reg = {"kernel_regularizer": None, "bias_regularizer": None, "activity_regularizer": None, "kernel_constraint": None,
           "bias_constraint": None}
    k = l.kind
    if k == "input":
        return "InputLayer", {"batch_input_shape": [None] + list(l.output.shape), "dtype": "float32", "sparse": False, "name": l.name}
    if k == "conv2d":
        return "Conv2D", dict(base, filters=l.filters, kernel_size=list(l.kernel_size), strides=[1, 1], padding=l.padding,
                              data_format="channels_last", dilation_rate=[1, 1], activation=l.activation or "linear",
                              use_bias=True, kernel_initializer=_INIT[l.kernel_initializer], bias_initializer=_ZEROS, **reg)
    if k == "conv2d_transpose":
        return "Conv2DTranspose", dict(base, filters=l.filters, kernel_size=[2, 2], strides=[2, 2], padding="same",
                                       data_format="channels_last", dilation_rate=[1, 1], activation="linear", use_bias=True,
                                       kernel_initializer=_INIT[l.kernel_initializer], bias_initializer=_ZEROS,
                                       output_padding=None, **reg)
    if k == "batch_normalization":
        return "BatchNormalization", dict(base, axis=-1, momentum=l.momentum, epsilon=l.epsilon, center=True, scale=True,
                                          beta_initializer=_ZEROS, gamma_initializer=_ONES, moving_mean_initializer=_ZEROS,
                                          moving_variance_initializer=_ONES, beta_regularizer=None, gamma_regularizer=None,
                                          beta_constraint=None, gamma_constraint=None)
    if k == "max_pooling2d":
        return "MaxPooling2D", dict(base, pool_size=[2, 2], padding="valid", strides=[2, 2], data_format="channels_last")
    if k == "dropout":
        return "Dropout", dict(base, rate=l.rate, noise_shape=None, seed=None)
    if k == "concatenate":
        return "Concatenate", dict(base, axis=3)
    if k == "flatten":
        return "Flatten", dict(base, data_format="channels_last")
    if k == "dense":
        return "Dense", dict(base, units=l.units, activation=l.activation or "linear", use_bias=True,
                             kernel_initializer=_INIT[l.kernel_initializer], bias_initializer=_ZEROS, **reg)
    raise ValueError("no Keras class for layer kind %r" % k)


def keras_config(graph, name="model_1"):
    layers = []
    for l in graph.layers:
        cls, cfg = _layer_config(l)
        inbound = [[[t.producer.name, 0, 0, {}] for t in l.inputs]] if l.inputs else []
        layers.append({"name": l.name, "class_name": cls, "config": cfg, "inbound_nodes": inbound})
    return {"class_name": "Model", "config": {"name": name, "layers": layers,
                                              "input_layers": [[graph.input.producer.name, 0, 0]],
                                              "output_layers": [[graph.output.producer.name, 0, 0]]},
            "keras_version": "2.3.1", "backend": "tensorflow"}


def _init_name(spec):
    if isinstance(spec, str):
        return spec
    c = (spec or {}).get("config", {})
    if (spec or {}).get("class_name") == "VarianceScaling" and c.get("mode") == "fan_in" and c.get("scale") == 2.0:
        return "he_normal"
    return "glorot_uniform"


def model_from_json(text, **model_kw):
    """keras.models.model_from_json for the layer classes the reference uses: rebuilds the Graph from a Keras
    functional-model ("Model") or Sequential JSON, keeping the layer names of the file."""
    cfg = json.loads(text)
    kind, conf = cfg["class_name"], cfg["config"]
    L.reset_names()
    tensors = {}
    entries = conf["layers"] if isinstance(conf, dict) else conf
    prev = None
    for e in entries:
        c, cls, name = e["config"], e["class_name"], e["config"].get("name", e.get("name"))
        if cls == "InputLayer":
            t = L.Input(tuple(c["batch_input_shape"][1:]), name=name)
            tensors[name] = prev = t
            continue
        if kind == "Sequential" and prev is None:
            prev = L.Input(tuple(c["batch_input_shape"][1:]))
            tensors[prev.producer.name] = prev
        if cls == "Conv2D":
            layer = L.Conv2D(c["filters"], tuple(c["kernel_size"]), activation=None if c["activation"] == "linear" else c["activation"],
                             padding=c["padding"], kernel_initializer=_init_name(c.get("kernel_initializer")), name=name)
        elif cls == "Conv2DTranspose":
            layer = L.Conv2DTranspose(c["filters"], tuple(c["kernel_size"]), strides=tuple(c["strides"]), padding=c["padding"],
                                      kernel_initializer=_init_name(c.get("kernel_initializer")), name=name)
        elif cls == "BatchNormalization":
            layer = L.BatchNormalization(momentum=c.get("momentum", 0.99), epsilon=c.get("epsilon", 1e-3), name=name)
        elif cls == "MaxPooling2D":
            layer = L.MaxPooling2D(tuple(c["pool_size"]), name=name)
        elif cls == "Dropout":
            layer = L.Dropout(c["rate"], name=name)
        elif cls == "Concatenate":
            layer = L.Concatenate(axis=c.get("axis", -1), name=name)
        elif cls == "Flatten":
            layer = L.Flatten(name=name)
        elif cls == "Dense":
            layer = L.Dense(c["units"], activation=None if c["activation"] == "linear" else c["activation"],
                            kernel_initializer=_init_name(c.get("kernel_initializer")), name=name)
        else:
            raise ValueError("model_from_json: layer class %r is outside the reference's layer set" % cls)
        if kind == "Sequential":
            ins = [prev]
        else:
            ins = [tensors[n[0]] for n in e["inbound_nodes"][0]]
        out = layer(ins if cls == "Concatenate" else ins[0])
        tensors[name] = prev = out
    if kind == "Sequential":
        first = next(t for t in tensors.values() if t.producer.kind == "input")
        return Model(inputs=[first], outputs=[prev], **model_kw)
    return Model(inputs=[tensors[conf["input_layers"][0][0]]], outputs=[tensors[conf["output_layers"][0][0]]], **model_kw)
