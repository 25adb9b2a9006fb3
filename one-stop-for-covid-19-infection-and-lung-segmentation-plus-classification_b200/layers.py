"""Keras-shaped graph declaration API (symbolic only -- no arithmetic happens here).

Mirrors the subset of `keras.layers` / `keras.models` the reference runners call (SURVEY.md 8b;
e.g. /root/reference/Scripts/task1_preprocessing_plus_unet_with_comments.py:853-915), with the same
constructor arguments and the same auto-naming (conv2d_1, batch_normalization_1, ...), so that the
graph sections of the runners read like the reference's.  A `Graph` is lowered to device op lists by
plan.py and executed by libb200unet.so; nothing in this file touches a tensor.
"""
from collections import OrderedDict
import math

_counters = {}


def reset_names():
    """keras.backend.clear_session(): restart auto-naming."""
    _counters.clear()


def _auto_name(kind):
    _counters[kind] = _counters.get(kind, 0) + 1
    return "%s_%d" % (kind, _counters[kind])


class SymTensor:
    """A symbolic NHWC activation (shape excludes the batch dim)."""

    def __init__(self, shape, producer=None):
        self.shape = tuple(int(s) for s in shape)
        self.producer = producer
        self.consumers = []

    @property
    def channels(self):
        return self.shape[-1]

    def __repr__(self):
        return "SymTensor(%s <- %s)" % (self.shape, self.producer.name if self.producer else "input")


class Layer:
    kind = "layer"
    _created = 0

    def __init__(self, name=None):
        self.name = name or _auto_name(self.kind)
        # keras orders model.layers by creation; the reference graphs are declared sequentially, so
        # creation order is both topological and what model.summary() prints
        Layer._created += 1
        self._seq = Layer._created
        self.inputs = []
        self.output = None
        self.weights = OrderedDict()      # short name -> (shape, initializer spec, trainable)

    def __call__(self, x):
        xs = list(x) if isinstance(x, (list, tuple)) else [x]
        self.inputs = xs
        for t in xs:
            t.consumers.append(self)
        self.output = SymTensor(self.build(xs), producer=self)
        return self.output

    def build(self, xs):
        raise NotImplementedError

    def count_params(self):
        return sum(int(math.prod(s)) for s, _, _ in self.weights.values())


class InputLayer(Layer):
    kind = "input"

    def __init__(self, shape, name=None):
        super().__init__(name)
        self.output = SymTensor(shape, producer=self)


def Input(shape, name=None):
    return InputLayer(shape, name).output


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


class Conv2D(Layer):
    kind = "conv2d"

    def __init__(self, filters, kernel_size, activation=None, padding="valid", kernel_initializer="glorot_uniform",
                 input_shape=None, name=None, **kw):
        super().__init__(name)
        self.filters = int(filters)
        self.kernel_size = _pair(kernel_size)
        self.activation = activation
        self.padding = padding
        self.kernel_initializer = kernel_initializer
        self.input_shape = input_shape
        if self.kernel_size not in ((3, 3), (1, 1)) or (self.kernel_size == (3, 3) and padding != "same"):
            raise ValueError("engine supports Conv2D 3x3 'same' and 1x1 only (got %s, %s)" % (self.kernel_size, padding))

    def build(self, xs):
        h, w, c = xs[0].shape
        kh, kw = self.kernel_size
        self.weights["kernel"] = ((kh, kw, c, self.filters), (self.kernel_initializer, kh * kw * c, kh * kw * self.filters), True)
        self.weights["bias"] = ((self.filters,), ("zeros",), True)
        return (h, w, self.filters)


class Conv2DTranspose(Layer):
    kind = "conv2d_transpose"

    def __init__(self, filters, kernel_size, strides=(1, 1), padding="valid", kernel_initializer="glorot_uniform",
                 name=None, **kw):
        super().__init__(name)
        self.filters = int(filters)
        if _pair(kernel_size) != (2, 2) or _pair(strides) != (2, 2):
            raise ValueError("engine supports Conv2DTranspose (2,2)/strides (2,2) only")
        self.kernel_initializer = kernel_initializer

    def build(self, xs):
        h, w, c = xs[0].shape
        # keras computes fans of the (kh,kw,Cout,Cin) kernel as fan_in = 4*Cout, fan_out = 4*Cin
        self.weights["kernel"] = ((2, 2, self.filters, c), (self.kernel_initializer, 4 * self.filters, 4 * c), True)
        self.weights["bias"] = ((self.filters,), ("zeros",), True)
        return (2 * h, 2 * w, self.filters)


class BatchNormalization(Layer):
    kind = "batch_normalization"

    def __init__(self, momentum=0.99, epsilon=1e-3, name=None, **kw):
        super().__init__(name)
        self.momentum, self.epsilon = float(momentum), float(epsilon)

    def build(self, xs):
        c = xs[0].shape[-1]
        self.weights["gamma"] = ((c,), ("ones",), True)
        self.weights["beta"] = ((c,), ("zeros",), True)
        self.weights["moving_mean"] = ((c,), ("zeros",), False)
        self.weights["moving_variance"] = ((c,), ("ones",), False)
        return xs[0].shape


class MaxPooling2D(Layer):
    kind = "max_pooling2d"

    def __init__(self, pool_size=(2, 2), name=None, **kw):
        super().__init__(name)
        if _pair(pool_size) != (2, 2):
            raise ValueError("engine supports MaxPooling2D((2,2)) only")

    def build(self, xs):
        h, w, c = xs[0].shape
        if h % 2 or w % 2:
            raise ValueError("MaxPooling2D needs even H, W (got %dx%d)" % (h, w))
        return (h // 2, w // 2, c)


class Dropout(Layer):
    kind = "dropout"

    def __init__(self, rate, name=None, **kw):
        super().__init__(name)
        self.rate = float(rate)

    def build(self, xs):
        return xs[0].shape


class Concatenate(Layer):
    kind = "concatenate"

    def __init__(self, axis=-1, name=None, **kw):
        super().__init__(name)
        if axis not in (-1, 3):
            raise ValueError("engine concatenates on the channel axis only")

    def build(self, xs):
        h, w, _ = xs[0].shape
        for t in xs:
            if t.shape[:2] != (h, w):
                raise ValueError("concatenate: spatial shapes differ")
        return (h, w, sum(t.shape[-1] for t in xs))


def concatenate(tensors, axis=-1, name=None):
    return Concatenate(axis=axis, name=name)(tensors)


class Flatten(Layer):
    kind = "flatten"

    def build(self, xs):
        return (int(math.prod(xs[0].shape)),)


class Dense(Layer):
    kind = "dense"

    def __init__(self, units, activation=None, kernel_initializer="glorot_uniform", name=None, **kw):
        super().__init__(name)
        self.units = int(units)
        self.activation = activation
        self.kernel_initializer = kernel_initializer

    def build(self, xs):
        (k,) = xs[0].shape
        self.weights["kernel"] = ((k, self.units), (self.kernel_initializer, k, self.units), True)
        self.weights["bias"] = ((self.units,), ("zeros",), True)
        return (self.units,)


class Graph:
    """The layer DAG between `inputs` and `outputs`, in creation (= topological) order."""

    def __init__(self, inputs, outputs):
        self.input = inputs[0] if isinstance(inputs, (list, tuple)) else inputs
        self.output = outputs[0] if isinstance(outputs, (list, tuple)) else outputs
        order, seen = [], set()

        def visit(t):
            l = t.producer
            if l is None or id(l) in seen:
                return
            seen.add(id(l))
            for i in l.inputs:
                visit(i)
            order.append(l)

        import sys
        old = sys.getrecursionlimit()
        sys.setrecursionlimit(10000)
        try:
            visit(self.output)
        finally:
            sys.setrecursionlimit(old)
        # a DFS post-order is topological but may interleave differently from keras' creation order
        self.layers = sorted(order, key=lambda l: l._seq)

    def weight_specs(self):
        """[(full name, shape, init spec, trainable)] in keras get_weights() order."""
        out = []
        for l in self.layers:
            for k, (shape, init, tr) in l.weights.items():
                out.append(("%s/%s" % (l.name, k), shape, init, tr))
        return out

    def count_params(self):
        tot = sum(int(math.prod(s)) for _, s, _, _ in self.weight_specs())
        tr = sum(int(math.prod(s)) for _, s, _, t in self.weight_specs() if t)
        return tot, tr, tot - tr

    def get_layer(self, name):
        for l in self.layers:
            if l.name == name:
                return l
        raise ValueError("No such layer: " + name)
