"""Quantiser ops -- same names and argument meaning as the reference's
``layers/quantized_ops.py``, evaluated by libqnnb200 CUDA kernels on torch CUDA tensors.

On the path: ``round_through`` (:8-14), ``clip_through`` (:17-31),
``quantize`` (:49-66), ``quantized_tanh`` (:87-100; this is the activation the models use,
bound as ``quantize_op`` in models/model_factory.py:9).  The other quantisers of that file are
never referenced by ``model_factory`` and raise ``NotImplementedError``.

Every op also accepts an ``engine.ActProbe`` and returns a probe describing itself, which is
how ``Activation(lambda x: quantized_tanh(x, nb=4))`` gets fused into a conv epilogue.
"""
from __future__ import annotations

from ..engine import ActProbe


def _k():
    from .. import kernels
    return kernels


def round_through(x):
    """Forward value of the straight-through rounding: tf.round, i.e. round-half-to-even."""
    return _k().round_half_even(x)


def clip_through(x, min, max):
    """Forward value of the straight-through clip."""
    return x.clamp(min, max)


def quantize(W, nb=16, clip_through=False):
    """``clip(round(W * 2^(nb-1)), -2^(nb-1), 2^(nb-1)-1) / 2^(nb-1)`` as fp32 values.
    (The reference's ``clip_through=True`` branch calls a bool and can never run.)"""
    if isinstance(W, ActProbe):
        return ActProbe(("quant", int(nb)))
    if clip_through:
        raise TypeError("'bool' object is not callable")     # what the reference does (quantized_ops.py:61-62)
    if not 2 <= int(nb) <= 8:
        raise ValueError("quantize: nb=%d outside 2..8 (levels are stored as int8)" % nb)
    return _k().quantize_act(W, int(nb)).to_float()


def quantized_tanh(W, nb=16):
    """Signed symmetric activation quantiser; same formula as :func:`quantize`."""
    return quantize(W, nb=nb)


def _off_path(name):
    def fn(*a, **k):
        raise NotImplementedError("%s is never referenced by models/model_factory.py and is outside the "
                                  "accelerated path (SURVEY.md section 2 row 1)" % name)
    fn.__name__ = name
    return fn


quantized_relu = _off_path("quantized_relu")
quantized_leakyrelu = _off_path("quantized_leakyrelu")
quantized_maxrelu = _off_path("quantized_maxrelu")
quantized_leakymaxrelu = _off_path("quantized_leakymaxrelu")
xnorize = _off_path("xnorize")
