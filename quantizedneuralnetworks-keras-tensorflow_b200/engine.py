"""A minimal Keras-shaped host layer: the pieces of the Keras API the reference's model builders
touch (models/vgg.py:1-44, models/resnet.py:7-147, models/model_factory.py:1-72), re-hosted on
torch device memory + libqnnb200.  It is the drop-in seam described in SURVEY.md section 8(b):
same layer names, constructor arguments and weight ordering; ``model.predict`` runs a fused plan
(``plan.py``) of hand-written CUDA kernels instead of a TensorFlow session.

No training: regularizers, constraints, lr multipliers and initializer names are accepted and
stored but inert.
"""
from __future__ import annotations

import re

import numpy as np

F32 = np.float32
_NAME_COUNTS: dict = {}
_RNG = np.random.default_rng(0)
_WEIGHTS_EPOCH = 0


def weights_epoch() -> int:
    """Process-wide counter bumped by every ``layer.set_weights``; plans use it as an O(1) staleness test before
    comparing per-layer versions (plan.Plan._sync_weights)."""
    return _WEIGHTS_EPOCH


def set_seed(seed: int):
    """Seed the generator used for Keras-style random initialisation of new layers."""
    global _RNG
    _RNG = np.random.default_rng(seed)


def reset_names():
    """keras.backend.clear_session() analogue for auto-generated layer names."""
    _NAME_COUNTS.clear()


def _snake(name):
    s = re.sub("(.)([A-Z][a-z0-9]+)", r"\1_\2", name)
    return re.sub("([a-z])([A-Z])", r"\1_\2", s).lower()


def _unique(base):
    _NAME_COUNTS[base] = _NAME_COUNTS.get(base, 0) + 1
    return "%s_%d" % (base, _NAME_COUNTS[base])


class KTensor:
    """Symbolic tensor of the functional API (batch dimension is None)."""

    def __init__(self, shape, layer=None, inputs=()):
        self.shape = tuple(shape)
        self.layer = layer
        self.inputs = tuple(inputs)

    @property
    def _keras_shape(self):
        return self.shape


class ActProbe:
    """Passed through an activation callable to discover which quantiser op it applies, so that
    ``Activation(lambda x: quantized_tanh(x, nb=4))`` style closures (model_factory.py:19-20) can
    be fused.  The ops in layers/*_ops.py return a new probe carrying the op spec."""

    def __init__(self, spec=None):
        self.spec = spec


class Layer:
    _base_name = None

    def __init__(self, name=None, input_shape=None, batch_input_shape=None, trainable=True, dtype=None, **kwargs):
        if kwargs:
            # Keras would raise on unknown kwargs; the reference only passes known conv/dense ones.
            raise TypeError("%s: unexpected keyword arguments %s" % (type(self).__name__, sorted(kwargs)))
        base = self._base_name or _snake(type(self).__name__)
        self.name = name or _unique(base)
        self.trainable = trainable
        self.built = False
        self._input_shape_arg = None
        if batch_input_shape is not None:
            self._input_shape_arg = tuple(batch_input_shape)
        elif input_shape is not None:
            self._input_shape_arg = (None,) + tuple(input_shape)
        self.input_shape = None
        self.output_shape = None
        self._wversion = 0

    def _touch(self):
        """Weights changed: any plan that cached device copies derived from them must rebuild (plan.py)."""
        global _WEIGHTS_EPOCH
        self._wversion += 1
        _WEIGHTS_EPOCH += 1

    # ---- Keras protocol
    def build(self, input_shape):
        self.built = True

    def compute_output_shape(self, input_shape):
        return input_shape

    def call(self, inputs):
        raise NotImplementedError

    def get_weights(self):
        return []

    def set_weights(self, weights):
        if len(weights):
            raise ValueError("layer %s has no weights" % self.name)

    def weight_names(self):
        return []

    def count_params(self):
        return int(sum(int(np.prod(w.shape)) for w in self.get_weights()))

    def get_config(self):
        return {"name": self.name, "trainable": self.trainable}

    def _ensure_built(self, shape):
        if not self.built:
            self.build(shape)
            self.built = True
        if self.input_shape is None:
            self.input_shape = shape
            self.output_shape = self.compute_output_shape(shape)

    def __call__(self, inputs):
        if isinstance(inputs, KTensor) or (isinstance(inputs, (list, tuple)) and inputs and isinstance(inputs[0], KTensor)):
            ins = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
            shape = [t.shape for t in ins] if isinstance(inputs, (list, tuple)) else ins[0].shape
            self._ensure_built(shape)
            return KTensor(self.compute_output_shape(shape), self, ins)
        # eager call on device tensors
        from .kernels import QTensor
        if isinstance(inputs, (list, tuple)):
            shape = [(None,) + tuple(t.shape[1:]) for t in inputs]
        else:
            shape = (None,) + tuple(inputs.shape[1:])
        self._ensure_built(shape)
        return self.call(inputs)


class InputLayer(Layer):
    def __init__(self, input_shape=None, **kw):
        super().__init__(input_shape=input_shape, **kw)


def Input(shape=None, batch_shape=None, name=None, **_):
    layer = InputLayer(input_shape=shape if batch_shape is None else tuple(batch_shape[1:]), name=name)
    full = layer._input_shape_arg
    layer.input_shape = layer.output_shape = full
    layer.built = True
    return KTensor(full, layer, ())


class BatchNormalization(Layer):
    """keras.layers.BatchNormalization, inference form: ``x*inv + (beta - mean*inv)``.
    Weights ``[gamma, beta, moving_mean, moving_variance]`` (model_factory.py:91)."""

    def __init__(self, axis=-1, momentum=0.99, epsilon=1e-3, center=True, scale=True, **kw):
        super().__init__(**kw)
        if axis not in (-1, 3, 1):
            raise ValueError("BatchNormalization: only the channels_last axis is supported")
        self.axis, self.momentum, self.epsilon = axis, momentum, float(epsilon)
        self.center, self.scale = center, scale
        self.gamma = self.beta = self.moving_mean = self.moving_variance = None

    def build(self, input_shape):
        ch = input_shape[-1]
        if ch is None:
            raise ValueError("BatchNormalization: the channel dimension must be defined")
        self.gamma = np.ones(ch, F32)
        self.beta = np.zeros(ch, F32)
        self.moving_mean = np.zeros(ch, F32)
        self.moving_variance = np.ones(ch, F32)
        self.built = True

    def get_weights(self):
        return [self.gamma, self.beta, self.moving_mean, self.moving_variance]

    def weight_names(self):
        return ["gamma", "beta", "moving_mean", "moving_variance"]

    def set_weights(self, weights):
        if len(weights) != 4:
            raise ValueError("BatchNormalization expects 4 arrays [gamma, beta, mean, var]")
        arrs = [np.asarray(w, F32) for w in weights]
        for a in arrs:
            if a.shape != self.gamma.shape:
                raise ValueError("BatchNormalization %s: bad weight shape %s" % (self.name, a.shape))
        self.gamma, self.beta, self.moving_mean, self.moving_variance = arrs
        self._touch()

    def constants(self):
        from .kernels import bn_constants
        return bn_constants(self.gamma, self.beta, self.moving_mean, self.moving_variance, self.epsilon)

    def call(self, x):
        import torch
        from . import kernels as K
        xf = K.as_qtensor(x).to_float()
        inv, shift = self.constants()
        dev = xf.device
        return K.batchnorm(xf, torch.from_numpy(inv).to(dev), torch.from_numpy(shift).to(dev))

    def get_config(self):
        c = super().get_config()
        c.update({"axis": self.axis, "momentum": self.momentum, "epsilon": self.epsilon,
                  "center": self.center, "scale": self.scale})
        return c


class MaxPooling2D(Layer):
    def __init__(self, pool_size=(2, 2), strides=None, padding="valid", **kw):
        super().__init__(**kw)
        ps = (pool_size, pool_size) if isinstance(pool_size, int) else tuple(pool_size)
        st = ps if strides is None else ((strides, strides) if isinstance(strides, int) else tuple(strides))
        if ps != (2, 2) or st != (2, 2) or padding != "valid":
            raise ValueError("MaxPooling2D: only pool_size=(2,2), strides=(2,2), padding='valid' is on the path (models/vgg.py:23)")
        self.pool_size, self.strides, self.padding = ps, st, padding

    def compute_output_shape(self, s):
        return (s[0], s[1] // 2, s[2] // 2, s[3])

    def call(self, x):
        from . import kernels as K
        return K.maxpool2(K.as_qtensor(x).to_float())


class AveragePooling2D(Layer):
    def __init__(self, pool_size=(2, 2), strides=None, padding="valid", **kw):
        super().__init__(**kw)
        ps = (pool_size, pool_size) if isinstance(pool_size, int) else tuple(pool_size)
        self.pool_size = ps
        if strides not in (None, ps, ps[0]) or padding != "valid":
            raise ValueError("AveragePooling2D: only non-overlapping valid pooling is supported")

    def compute_output_shape(self, s):
        return (s[0], s[1] // self.pool_size[0], s[2] // self.pool_size[1], s[3])

    def call(self, x):
        raise NotImplementedError("AveragePooling2D is only available fused into the classifier head "
                                  "(AveragePooling2D -> Flatten -> Dense, models/resnet.py:134-140)")


class Flatten(Layer):
    def compute_output_shape(self, s):
        return (s[0], int(np.prod(s[1:])))

    def call(self, x):
        from . import kernels as K
        xf = K.as_qtensor(x).to_float()
        return xf.reshape(xf.shape[0], -1)


class ZeroPadding2D(Layer):
    def __init__(self, padding=(1, 1), **kw):
        super().__init__(**kw)
        p = (padding, padding) if isinstance(padding, int) else tuple(padding)
        self.padding = p

    def compute_output_shape(self, s):
        return (s[0], s[1] + 2 * self.padding[0], s[2] + 2 * self.padding[1], s[3])

    def call(self, x):
        import torch.nn.functional as Fn
        from . import kernels as K
        q = K.as_qtensor(x)
        if q.kind == "b1":
            raise ValueError("ZeroPadding2D on a bit-packed tensor is undefined (0 is not a +-1 level)")
        ph, pw = self.padding
        data = Fn.pad(q.data, (0, 0, pw, pw, ph, ph))       # level 0 == value 0 for u8 / i8 / f32
        return K.QTensor(q.kind, data.contiguous(), q.scale, q.channels)


class LeakyReLU(Layer):
    _base_name = "leaky_re_lu"

    def __init__(self, alpha=0.3, **kw):
        super().__init__(**kw)
        self.alpha = float(F32(alpha))

    def call(self, x):
        from . import kernels as K
        return K.leaky(K.as_qtensor(x).to_float(), self.alpha)

    def act_spec(self):
        return ("leaky", self.alpha)


class Activation(Layer):
    """keras.layers.Activation(callable).  The callable is probed once with an ``ActProbe`` so the
    quantiser it applies can be fused into the producing kernel's epilogue."""

    def __init__(self, activation, **kw):
        super().__init__(**kw)
        self.activation = activation
        self._spec = None
        if callable(activation):
            try:
                r = activation(ActProbe())
                if isinstance(r, ActProbe) and r.spec is not None:
                    self._spec = r.spec
            except Exception:
                self._spec = None
        elif activation in ("linear", None):
            self._spec = ("linear",)
        elif activation == "softmax":
            self._spec = ("softmax",)
        if self._spec is None:
            raise ValueError("Activation: unsupported activation %r (supported: the quantiser ops of layers/*_ops.py, "
                             "'softmax', 'linear')" % (activation,))

    def act_spec(self):
        return self._spec

    def call(self, x):
        from . import kernels as K
        if self._spec[0] == "linear":
            return x
        return self.activation(K.as_qtensor(x).to_float())


class Add(Layer):
    def compute_output_shape(self, shapes):
        return shapes[0]

    def call(self, xs):
        from . import kernels as K
        a, b = (K.as_qtensor(t).to_float() for t in xs)
        return a + b


def add(inputs, **kw):
    return Add(**kw)(inputs)


class Lambda(Layer):
    """keras.layers.Lambda.  The only use on the path is ``Lambda(lambda x: x * 0.5)``
    (models/resnet.py:128); the multiplier is recovered by probing the function."""

    def __init__(self, function, **kw):
        super().__init__(**kw)
        self.function = function
        try:
            one, two = float(function(1.0)), float(function(2.0))
        except Exception as e:
            raise ValueError("Lambda: only scalar multiples (x * c) are supported on the path") from e
        if abs(two - 2.0 * one) > 1e-12:
            raise ValueError("Lambda: only scalar multiples (x * c) are supported on the path")
        self.multiplier = one

    def call(self, x):
        from . import kernels as K
        return K.as_qtensor(x).to_float() * self.multiplier


# --------------------------------------------------------------------------- models
class _ModelBase:
    def __init__(self, name=None):
        self.name = name or _unique(_snake(type(self).__name__))
        self._plans = {}

    # to be provided by subclasses: self.layers (ordered), self._nodes() -> [(layer, [input node ids])]
    def _invalidate(self):
        """Drop the fused plans (graph structure or weights changed); their captured graphs, static buffers and
        cached constants are released right here, not whenever the garbage collector gets to them."""
        for pl in self._plans.values():
            pl.close()
        self._plans = {}

    def get_layer(self, name=None, index=None):
        if index is not None:
            return self.layers[index]
        for l in self.layers:
            if l.name == name:
                return l
        raise ValueError("No such layer: %s" % name)

    def get_weights(self):
        out = []
        for l in self.layers:
            out.extend(l.get_weights())
        return out

    def set_weights(self, weights):
        i = 0
        for l in self.layers:
            n = len(l.get_weights())
            l.set_weights(weights[i:i + n])
            i += n
        if i != len(weights):
            raise ValueError("set_weights: expected %d arrays, got %d" % (i, len(weights)))
        self._invalidate()

    def count_params(self):
        return int(sum(l.count_params() for l in self.layers))

    def summary(self, print_fn=print):
        line = "_" * 65
        print_fn(line)
        print_fn("%-29s%-26s%-10s" % ("Layer (type)", "Output Shape", "Param #"))
        print_fn("=" * 65)
        for l in self.layers:
            print_fn("%-29s%-26s%-10d" % ("%s (%s)" % (l.name, type(l).__name__), str(l.output_shape), l.count_params()))
        print_fn("=" * 65)
        print_fn("Total params: {:,}".format(self.count_params()))
        print_fn(line)

    def load_weights(self, path):
        from .hdf5_lite import load_keras_weights
        load_keras_weights(self, path)
        self._invalidate()

    # ---- inference
    def plan(self, impl=0, sm_share=0):
        """The fused launch plan for kernel selection ``impl`` (0 = automatic); ``sm_share`` = SMs per persistent kernel
        (0 = all; see plan.Plan)."""
        from .plan import Plan
        key = (int(impl), int(sm_share)) if sm_share else int(impl)
        if key not in self._plans:
            self._plans[key] = Plan(self, impl=int(impl), sm_share=int(sm_share))
        return self._plans[key]

    def predict(self, x, batch_size=None, verbose=0, impl=0, return_logits=False):
        """Forward pass.  ``x``: uint8 pixel levels or float32 values, NHWC; a numpy array (host;
        copied to ``cuda:0``, result returned as numpy) or a torch CUDA tensor (result stays on the
        device).  uint8 input is the deployed format: level/255 is what utils/load_data.py:40
        feeds the reference."""
        return self.plan(impl).predict(x, batch_size=batch_size, return_logits=return_logits)

    def evaluate(self, x, y, batch_size=None, verbose=0, loss=None):
        """``model.evaluate`` as used at train.py:163-169: returns ``[loss, accuracy]``.  ``y`` is one-hot
        (ResNet, categorical cross-entropy, train.py:101-107) or +-1 targets (VGG, squared hinge, train.py:68-100,
        utils/load_data.py:63-66); ``loss`` overrides the choice ('categorical_crossentropy' | 'squared_hinge').
        The forward pass runs on the GPU; the reduction over N x classes scores is host arithmetic."""
        import numpy as np
        scores = self.predict(x, batch_size=batch_size)
        scores = scores if isinstance(scores, np.ndarray) else scores.detach().cpu().numpy()
        y = np.asarray(y, dtype=np.float32)
        if loss is None:
            loss = "squared_hinge" if y.min() < 0 else "categorical_crossentropy"
        if loss == "squared_hinge":
            lv = float(np.mean(np.maximum(1.0 - y * scores, 0.0) ** 2))
        elif loss == "categorical_crossentropy":
            p = np.clip(scores / scores.sum(axis=1, keepdims=True), 1e-7, 1 - 1e-7)
            lv = float(np.mean(-(y * np.log(p)).sum(axis=1)))
        else:
            raise ValueError("unsupported loss %r" % loss)
        acc = float(np.mean(scores.argmax(1) == y.argmax(1)))
        return [lv, acc]

    def predict_async(self, x, impl=0):
        """Pipelined variant for host batches: enqueue H2D copy + CUDA-graph replay + D2H copy on one of the
        plan's streams and return a handle; ``handle.result()`` gives the fp32 (N, classes) CPU tensor."""
        return self.plan(impl).predict_async(x)



class Sequential(_ModelBase):
    def __init__(self, layers=None, name=None):
        super().__init__(name)
        self.layers = []
        self._out = None
        for l in layers or []:
            self.add(l)

    def add(self, layer):
        if self._out is None:
            if isinstance(layer, InputLayer):
                self._out = KTensor(layer._input_shape_arg, layer, ())
                self._in = self._out
                return
            if layer._input_shape_arg is None:
                raise ValueError("The first layer in a Sequential model must get an `input_shape` argument.")
            self._in = Input(batch_shape=layer._input_shape_arg)
            self._out = self._in
        self._out = layer(self._out)
        self.layers.append(layer)
        self._invalidate()

    @property
    def input_shape(self):
        return self._in.shape

    @property
    def output_shape(self):
        return self._out.shape

    def _graph(self):
        return [self._in], [self._out]


class Model(_ModelBase):
    def __init__(self, inputs, outputs, name=None):
        super().__init__(name)
        self._ins = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
        self._outs = list(outputs) if isinstance(outputs, (list, tuple)) else [outputs]
        if len(self._ins) != 1 or len(self._outs) != 1:
            raise ValueError("Model: exactly one input and one output are supported")
        self.input, self.output = self._ins[0], self._outs[0]
        order = topo_order(self._outs)
        self.layers = [t.layer for t in order if not isinstance(t.layer, InputLayer)]

    @property
    def input_shape(self):
        return self.input.shape

    @property
    def output_shape(self):
        return self.output.shape

    def _graph(self):
        return self._ins, self._outs


def topo_order(outputs):
    """KTensors in creation-consistent topological order (inputs first)."""
    seen, order = set(), []

    def visit(t):
        if id(t) in seen:
            return
        seen.add(id(t))
        for s in t.inputs:
            visit(s)
        order.append(t)

    for o in outputs:
        visit(o)
    return order
