"""Shared machinery of the six custom layers (Quantized/Binary/Ternary x Conv2D/Dense).

Mirrors the Keras ``Conv2D`` / ``Dense`` constructor surface the reference subclasses
(layers/quantized_layers.py:32-206, binary_layers.py:31-199, ternary_layers.py:30-186): same
keyword arguments, ``build(input_shape)``, ``call(inputs)``, ``get_config()``, weight order
``[kernel, bias]`` with the kernel in HWIO / (in, units) layout.  Training-only arguments
(regularizers, constraints, lr multipliers, initializer names) are stored but inert.
"""
from __future__ import annotations

import numpy as np

from .. import engine
from ..engine import Layer, F32


class Clip:
    """Weight-clipping constraint (layers/quantized_layers.py:13-29); training-only, kept as data."""

    def __init__(self, min_value, max_value=None):
        self.min_value = min_value
        self.max_value = max_value
        if not self.max_value:
            self.max_value = -self.min_value
        if self.min_value > self.max_value:
            self.min_value, self.max_value = self.max_value, self.min_value

    def __call__(self, p):
        return p.clip(self.min_value, self.max_value) if isinstance(p, np.ndarray) else p.clamp(self.min_value, self.max_value)

    def get_config(self):
        return {"name": "__call__", "min_value": self.min_value, "max_value": self.max_value}


def _pair(v):
    return (int(v), int(v)) if isinstance(v, int) else tuple(int(a) for a in v)


def _act_spec(activation):
    """Resolve a Keras ``activation=`` argument to a fusable spec (None == linear)."""
    if activation is None or activation == "linear":
        return None
    if activation == "softmax":
        return ("softmax",)
    if callable(activation):
        r = activation(engine.ActProbe())
        if isinstance(r, engine.ActProbe) and r.spec is not None:
            return r.spec
    raise ValueError("unsupported activation %r" % (activation,))


class QLinearBase(Layer):
    """Fields common to conv and dense: weight quantiser, H, multipliers."""
    WEIGHT_KIND = None          # 'quantized' | 'binary' | 'ternary' | 'float'

    def _init_common(self, H, nb, kernel_lr_multiplier, bias_lr_multiplier, use_bias, activation,
                     kernel_initializer, bias_initializer, kernel_regularizer, bias_regularizer,
                     activity_regularizer, kernel_constraint, bias_constraint):
        self.H = H
        self.nb = nb
        self.kernel_lr_multiplier = kernel_lr_multiplier
        self.bias_lr_multiplier = bias_lr_multiplier
        self.use_bias = bool(use_bias)
        self.activation = activation
        self._act = _act_spec(activation)
        self.kernel_initializer = kernel_initializer
        self.bias_initializer = bias_initializer
        self.kernel_regularizer = kernel_regularizer
        self.bias_regularizer = bias_regularizer
        self.activity_regularizer = activity_regularizer
        self.kernel_constraint = kernel_constraint
        self.bias_constraint = bias_constraint
        self.kernel = None
        self.bias = None
        self._packed = {}

    # -- weight quantiser of this layer family: (mode, nb, H, weight scale)
    def weight_mode(self):
        from .. import _lib as L
        if self.WEIGHT_KIND == "quantized":
            nb = int(self.nb)
            if not 2 <= nb <= 8:
                raise ValueError("%s: nb=%d outside 2..8 -- weight levels are stored as int8" % (self.name, nb))
            return L.W_QUANT, nb, 1.0, 1.0 / float(1 << (nb - 1))
        if self.WEIGHT_KIND == "binary":
            return L.W_BINARY, 1, float(self.H), float(self.H)
        if self.WEIGHT_KIND == "float":
            return L.W_FLOAT, 0, 1.0, 1.0
        return L.W_TERNARY, 2, float(self.H), float(self.H)

    def _resolve_glorot(self, nb_input, nb_output):
        # layers/quantized_layers.py:126-136 (conv), :49-54 (dense)
        if isinstance(self.H, str) and self.H == "Glorot":
            self.H = F32(np.sqrt(1.5 / (nb_input + nb_output)))
        if isinstance(self.kernel_lr_multiplier, str) and self.kernel_lr_multiplier == "Glorot":
            self.kernel_lr_multiplier = F32(1.0 / np.sqrt(1.5 / (nb_input + nb_output)))

    def _create_weights(self, kernel_shape, nout):
        H = float(self.H)
        self.kernel_constraint = Clip(-H, H)
        self.kernel_initializer = ("RandomUniform", -H, H)        # quantized_layers.py:140 overrides the kwarg
        self.kernel = engine._RNG.uniform(-H, H, size=kernel_shape).astype(F32)
        if self.use_bias:
            self.lr_multipliers = [self.kernel_lr_multiplier, self.bias_lr_multiplier]
            self.bias = np.zeros((nout,), F32)
        else:
            self.lr_multipliers = [self.kernel_lr_multiplier]
            self.bias = None

    def get_weights(self):
        return [self.kernel] + ([self.bias] if self.use_bias else [])

    def weight_names(self):
        return ["kernel"] + (["bias"] if self.use_bias else [])

    def set_weights(self, weights):
        want = 2 if self.use_bias else 1
        if len(weights) != want:
            raise ValueError("%s expects %d weight arrays, got %d" % (self.name, want, len(weights)))
        k = np.asarray(weights[0], F32)
        if k.shape != self.kernel.shape:
            raise ValueError("%s: kernel shape %s != %s" % (self.name, k.shape, self.kernel.shape))
        self.kernel = np.ascontiguousarray(k)
        if self.use_bias:
            b = np.asarray(weights[1], F32)
            if b.shape != self.bias.shape:
                raise ValueError("%s: bias shape %s != %s" % (self.name, b.shape, self.bias.shape))
            self.bias = np.ascontiguousarray(b)
        self._packed = {}
        self._touch()

    # -- device-side caches
    def packed_kernel(self, device, wfmt, kernel_override=None, tag=""):
        """Quantise + pack once per (device, format); K0 runs on the GPU."""
        import torch
        from .. import kernels as K
        key = (str(device), int(wfmt), tag)
        if key not in self._packed:
            mode, nb, H, _ = self.weight_mode()
            src = self.kernel if kernel_override is None else kernel_override
            kt = torch.from_numpy(np.ascontiguousarray(src)).to(device)
            self._packed[key] = K.pack_weights(kt, mode, nb, H, wfmt)
        return self._packed[key]

    def bias_tensor(self, device):
        import torch
        if not self.use_bias:
            return None
        key = (str(device), "bias")
        if key not in self._packed:
            self._packed[key] = torch.from_numpy(self.bias).to(device)
        return self._packed[key]

    def _common_config(self):
        return {"H": float(self.H) if not isinstance(self.H, str) else self.H,
                "kernel_lr_multiplier": (float(self.kernel_lr_multiplier)
                                         if not isinstance(self.kernel_lr_multiplier, (str, type(None))) else self.kernel_lr_multiplier),
                "bias_lr_multiplier": self.bias_lr_multiplier}


class QConv2DBase(QLinearBase):
    def __init__(self, filters, kernel_size=(1, 1), strides=(1, 1), padding="valid", data_format=None,
                 dilation_rate=(1, 1), activation=None, use_bias=True, kernel_initializer="glorot_uniform",
                 bias_initializer="zeros", kernel_regularizer=None, bias_regularizer=None,
                 activity_regularizer=None, kernel_constraint=None, bias_constraint=None,
                 kernel_lr_multiplier="Glorot", bias_lr_multiplier=None, H=1.0, nb=16, **kwargs):
        super().__init__(**kwargs)
        self.filters = int(filters)
        self.kernel_size = _pair(kernel_size)
        self.strides = _pair(strides)
        self.padding = padding
        self.data_format = data_format or "channels_last"
        self.dilation_rate = _pair(dilation_rate)
        self._init_common(H, nb, kernel_lr_multiplier, bias_lr_multiplier, use_bias, activation,
                          kernel_initializer, bias_initializer, kernel_regularizer, bias_regularizer,
                          activity_regularizer, kernel_constraint, bias_constraint)
        if self.data_format != "channels_last":
            raise ValueError("only data_format='channels_last' is supported (reference README.md:17)")
        if self.dilation_rate != (1, 1):
            raise ValueError("dilation_rate != 1 is not on the path")
        if self.strides[0] != self.strides[1] or self.strides[0] not in (1, 2):
            raise ValueError("strides must be (1,1) or (2,2), got %s" % (self.strides,))
        if max(self.kernel_size) > 3:
            raise ValueError("kernel_size up to 3x3 is supported, got %s" % (self.kernel_size,))
        if self.padding != "same" and not (self.padding == "valid" and self.kernel_size == (1, 1)):
            raise ValueError("only padding='same' is on the path (models/vgg.py:9, models/resnet.py:51)")

    def build(self, input_shape):
        channel_axis = -1
        if input_shape[channel_axis] is None:
            raise ValueError("The channel dimension of the inputs should be defined. Found `None`.")
        input_dim = int(input_shape[channel_axis])
        kernel_shape = self.kernel_size + (input_dim, self.filters)
        base = self.kernel_size[0] * self.kernel_size[1]
        self._resolve_glorot(int(input_dim * base), int(self.filters * base))
        self._create_weights(kernel_shape, self.filters)
        self.built = True

    def compute_output_shape(self, s):
        st = self.strides[0]
        return (s[0], -(-s[1] // st), -(-s[2] // st), self.filters)

    def call(self, inputs):
        """Un-fused forward: conv + bias (+ activation), fp32 values out -- what the reference's
        ``call`` returns.  Inside ``model.predict`` the fused plan is used instead."""
        from .. import _lib as L, kernels as K
        x = K.as_qtensor(inputs)
        wfmt = L.WFMT_B1 if x.kind == "b1" else L.WFMT_I8
        if wfmt == L.WFMT_B1 and self.WEIGHT_KIND != "binary":
            x = K.QTensor("f32", x.to_float(), 1.0, x.channels)
            wfmt = L.WFMT_I8
        if self.WEIGHT_KIND == "float":
            x = x if x.kind == "f32" else K.QTensor("f32", x.to_float(), 1.0, x.channels)
            wfmt = L.WFMT_F32
        dev = x.data.device
        _, _, _, wscale = self.weight_mode()
        scale = K.acc_scale(x.scale if x.kind in ("u8", "i8") else 1.0, wscale)
        act, abits, alpha, post = L.ACT_NONE, 0, 0.3, None
        if self._act is not None:
            post = self._act
        epi = K.make_epilogue(scale, bias=self.bias_tensor(dev), act=act, abits=abits, leaky_alpha=alpha)
        y = K.conv2d(x, self.packed_kernel(dev, wfmt), self.kernel_size[0], self.kernel_size[1], self.filters,
                     self.strides[0], epi).data
        if post is not None:
            y = self.activation(y)
        return y

    def get_config(self):
        c = super().get_config()
        c.update({"filters": self.filters, "kernel_size": self.kernel_size, "strides": self.strides,
                  "padding": self.padding, "data_format": self.data_format, "dilation_rate": self.dilation_rate,
                  "activation": getattr(self.activation, "__name__", self.activation), "use_bias": self.use_bias})
        c.update(self._common_config())
        return c


class QDenseBase(QLinearBase):
    def __init__(self, units, activation=None, use_bias=True, kernel_initializer="glorot_uniform",
                 bias_initializer="zeros", kernel_regularizer=None, bias_regularizer=None,
                 activity_regularizer=None, kernel_constraint=None, bias_constraint=None,
                 H=1.0, nb=16, kernel_lr_multiplier="Glorot", bias_lr_multiplier=None, **kwargs):
        super().__init__(**kwargs)
        self.units = int(units)
        self._init_common(H, nb, kernel_lr_multiplier, bias_lr_multiplier, use_bias, activation,
                          kernel_initializer, bias_initializer, kernel_regularizer, bias_regularizer,
                          activity_regularizer, kernel_constraint, bias_constraint)

    def build(self, input_shape):
        assert len(input_shape) >= 2
        input_dim = int(input_shape[1])
        self._resolve_glorot(input_dim, self.units)
        self._create_weights((input_dim, self.units), self.units)
        self.built = True

    def compute_output_shape(self, s):
        return (s[0], self.units)

    def call(self, inputs):
        from .. import _lib as L, kernels as K
        x = K.as_qtensor(inputs)
        if x.kind == "u8":
            x = K.QTensor("f32", x.to_float(), 1.0, x.channels)
        wfmt = L.WFMT_B1 if x.kind == "b1" else L.WFMT_I8
        if wfmt == L.WFMT_B1 and (self.WEIGHT_KIND != "binary" or x.channels % 32):
            x = K.QTensor("f32", x.to_float(), 1.0, x.channels)
            wfmt = L.WFMT_I8
        if self.WEIGHT_KIND == "float":
            x = x if x.kind == "f32" else K.QTensor("f32", x.to_float(), 1.0, x.channels)
            wfmt = L.WFMT_F32
        dev = x.data.device
        _, _, _, wscale = self.weight_mode()
        scale = K.acc_scale(x.scale if x.kind == "i8" else 1.0, wscale)
        epi = K.make_epilogue(scale, bias=self.bias_tensor(dev))
        softmax = self._act is not None and self._act[0] == "softmax"
        y, _ = K.dense(x, self.packed_kernel(dev, wfmt), self.units, epi, softmax=softmax)
        if self._act is not None and not softmax:
            y = self.activation(y)
        return y

    def get_config(self):
        c = super().get_config()
        c.update({"units": self.units, "activation": getattr(self.activation, "__name__", self.activation),
                  "use_bias": self.use_bias})
        c.update(self._common_config())
        return c
