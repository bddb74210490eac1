import re

import numpy as np
import torch
import torch.nn.functional as F

from .. import activations, initializers
from .. import backend as K

_counts = {}


def reset_names():
    _counts.clear()


def _snake(name):
    s = re.sub("(.)([A-Z][a-z0-9]+)", r"\1_\2", name)
    return re.sub("([a-z])([A-Z])", r"\1_\2", s).lower()


class InputSpec(object):
    def __init__(self, **kw):
        self.__dict__.update(kw)


class Tracer(object):
    """Functional-API tensor: carries a one-sample dummy value so layers can build eagerly."""

    def __init__(self, value, layer=None, parents=(), multi=False):
        self.value, self.layer, self.parents, self.multi = value, layer, tuple(parents), multi

    @property
    def shape(self):
        return self.value.shape


class Layer(object):
    def __init__(self, **kwargs):
        self._kw = kwargs
        base = _snake(type(self).__name__)
        if "name" in kwargs and kwargs["name"]:
            self.name = kwargs["name"]
        elif not hasattr(self, "name"):
            _counts[base] = _counts.get(base, 0) + 1
            self.name = "%s_%d" % (base, _counts[base])
        if not hasattr(self, "weights"):
            self.weights = []
        self.built = getattr(self, "built", False)

    def add_weight(self, shape=None, initializer=None, name=None, regularizer=None, constraint=None, **kw):
        init = initializers.get(initializer)
        w = torch.as_tensor(np.asarray(init(tuple(shape)), np.float32))
        self.weights.append(w)
        return w

    def build(self, input_shape):
        self.built = True

    def get_weights(self):
        return [w.numpy().copy() for w in self.weights]

    def set_weights(self, ws):
        assert len(ws) == len(self.weights)
        for w, v in zip(self.weights, ws):
            v = torch.as_tensor(np.asarray(v, np.float32))
            assert tuple(v.shape) == tuple(w.shape), (self.name, v.shape, w.shape)
            w.copy_(v)

    def _shape_of(self, x):
        return (None,) + tuple(x.shape[1:])

    def __call__(self, inputs):
        multi = isinstance(inputs, (list, tuple))
        ins = list(inputs) if multi else [inputs]
        traced = isinstance(ins[0], Tracer)
        vals = [i.value if traced else i for i in ins]
        if not self.built:
            self.build([self._shape_of(v) for v in vals] if multi else self._shape_of(vals[0]))
            self.built = True
        with torch.no_grad():
            out = self.call(vals if multi else vals[0])
        if traced:
            return Tracer(out, self, ins, multi)
        return out

    def call(self, x):
        return x

    def get_config(self):
        return {"name": self.name}


class InputLayer(Layer):
    pass


def Input(shape=None, **kw):
    return Tracer(torch.zeros((1,) + tuple(shape)), InputLayer(), ())


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


class Conv2D(Layer):
    def __init__(self, filters, kernel_size=(1, 1), strides=(1, 1), padding="valid", data_format=None, dilation_rate=(1, 1),
                 activation=None, use_bias=True, kernel_initializer="glorot_uniform", bias_initializer="zeros",
                 kernel_regularizer=None, bias_regularizer=None, activity_regularizer=None, kernel_constraint=None,
                 bias_constraint=None, **kwargs):
        super(Conv2D, self).__init__(**kwargs)
        self.filters, self.kernel_size, self.strides = filters, _pair(kernel_size), _pair(strides)
        self.padding, self.data_format, self.dilation_rate = padding, data_format or "channels_last", _pair(dilation_rate)
        self.activation, self.use_bias = activations.get(activation), use_bias
        self.kernel_initializer, self.bias_initializer = kernel_initializer, bias_initializer
        self.kernel_regularizer, self.bias_regularizer = kernel_regularizer, bias_regularizer
        self.activity_regularizer, self.kernel_constraint, self.bias_constraint = activity_regularizer, kernel_constraint, bias_constraint

    # keras.layers.Conv2D itself (network_type 'float', models/model_factory.py:24-25); the reference's custom layers
    # override build / call
    def build(self, input_shape):
        cin = int(input_shape[-1])
        self.kernel = self.add_weight(shape=self.kernel_size + (cin, self.filters), initializer=self.kernel_initializer, name="kernel")
        self.bias = self.add_weight(shape=(self.filters,), initializer=self.bias_initializer, name="bias") if self.use_bias else None
        self.built = True

    def call(self, x):
        from .. import backend as K
        y = K.conv2d(x, self.kernel, strides=self.strides, padding=self.padding, data_format=self.data_format, dilation_rate=self.dilation_rate)
        if self.use_bias:
            y = K.bias_add(y, self.bias, data_format=self.data_format)
        return self.activation(y) if self.activation is not None else y

    def get_config(self):
        return {"name": self.name, "filters": self.filters}


class Dense(Layer):
    def __init__(self, units, activation=None, use_bias=True, kernel_initializer="glorot_uniform", bias_initializer="zeros",
                 kernel_regularizer=None, bias_regularizer=None, activity_regularizer=None, kernel_constraint=None,
                 bias_constraint=None, **kwargs):
        super(Dense, self).__init__(**kwargs)
        self.units, self.activation, self.use_bias = units, activations.get(activation), use_bias
        self.kernel_initializer, self.bias_initializer = kernel_initializer, bias_initializer
        self.kernel_regularizer, self.bias_regularizer = kernel_regularizer, bias_regularizer
        self.activity_regularizer, self.kernel_constraint, self.bias_constraint = activity_regularizer, kernel_constraint, bias_constraint

    # keras.layers.Dense itself (network_type 'float')
    def build(self, input_shape):
        self.kernel = self.add_weight(shape=(int(input_shape[-1]), self.units), initializer=self.kernel_initializer, name="kernel")
        self.bias = self.add_weight(shape=(self.units,), initializer=self.bias_initializer, name="bias") if self.use_bias else None
        self.built = True

    def call(self, x):
        from .. import backend as K
        y = K.dot(x, self.kernel)
        if self.use_bias:
            y = K.bias_add(y, self.bias)
        return self.activation(y) if self.activation is not None else y

    def get_config(self):
        return {"name": self.name, "units": self.units}


class BatchNormalization(Layer):
    def __init__(self, axis=-1, momentum=0.99, epsilon=1e-3, **kwargs):
        super(BatchNormalization, self).__init__(**kwargs)
        self.momentum, self.epsilon = momentum, epsilon

    def build(self, input_shape):
        ch = input_shape[-1]
        for init in (np.ones, np.zeros, np.zeros, np.ones):          # gamma, beta, moving_mean, moving_variance
            self.weights.append(torch.as_tensor(init(ch, np.float32)))

    def call(self, x):
        g, b, mu, var = self.weights
        inv = torch.rsqrt(var + self.epsilon) * g                     # tf.nn.batch_normalization
        return x * inv + (b - mu * inv)


class Activation(Layer):
    def __init__(self, activation, **kwargs):
        super(Activation, self).__init__(**kwargs)
        self.activation = activations.get(activation)

    def call(self, x):
        return self.activation(x)


class MaxPooling2D(Layer):
    def __init__(self, pool_size=(2, 2), **kwargs):
        super(MaxPooling2D, self).__init__(**kwargs)
        self.pool_size = _pair(pool_size)

    def call(self, x):
        return F.max_pool2d(x.permute(0, 3, 1, 2), self.pool_size).permute(0, 2, 3, 1).contiguous()


class AveragePooling2D(Layer):
    def __init__(self, pool_size=(2, 2), **kwargs):
        super(AveragePooling2D, self).__init__(**kwargs)
        self.pool_size = _pair(pool_size)

    def call(self, x):
        return F.avg_pool2d(x.permute(0, 3, 1, 2), self.pool_size).permute(0, 2, 3, 1).contiguous()


class Flatten(Layer):
    def call(self, x):
        return x.reshape(x.shape[0], -1)


class ZeroPadding2D(Layer):
    def __init__(self, padding=(1, 1), **kwargs):
        super(ZeroPadding2D, self).__init__(**kwargs)
        self.padding = _pair(padding)

    def call(self, x):
        ph, pw = self.padding
        return F.pad(x, (0, 0, pw, pw, ph, ph))


class Lambda(Layer):
    def __init__(self, function, **kwargs):
        super(Lambda, self).__init__(**kwargs)
        self.function = function

    def call(self, x):
        return self.function(x)


class Add(Layer):
    def call(self, xs):
        out = xs[0]
        for t in xs[1:]:
            out = out + t
        return out


def add(inputs, **kw):
    return Add(**kw)(inputs)


class Reshape(Layer):
    pass


def concatenate(*a, **k):
    raise NotImplementedError


class SimpleRNN(Layer):          # imported (unused) by layers/ternary_layers.py:6
    pass
