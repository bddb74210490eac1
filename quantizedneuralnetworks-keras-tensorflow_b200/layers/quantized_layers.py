"""``QuantizedConv2D`` / ``QuantizedDense`` -- drop-in for the reference's
``layers/quantized_layers.py`` (constructors :37-43 and :104-112, ``build`` :45-77 / :114-162,
``call`` :79-88 / :164-194, ``get_config`` :91-96 / :196-201).

``call`` computes ``conv2d(x, quantize(kernel, nb)) + bias``.  The reference additionally wraps
the conv in a gradient-scaling identity (:167-180) that only matters for the backward pass; it is
a value identity and is dropped here (its fp32 rounding noise is quantified in SURVEY.md App. A.5).
"""
from ._base import Clip, QConv2DBase, QDenseBase
from .quantized_ops import quantize, clip_through  # noqa: F401  (re-exported like the reference)


class QuantizedDense(QDenseBase):
    """n-bit weight Dense layer.  NB: models/model_factory.py:31 builds it with ``nb=cf.abits``."""
    WEIGHT_KIND = "quantized"

    def __init__(self, units, H=1., nb=16, kernel_lr_multiplier='Glorot', bias_lr_multiplier=None, **kwargs):
        super().__init__(units, H=H, nb=nb, kernel_lr_multiplier=kernel_lr_multiplier,
                         bias_lr_multiplier=bias_lr_multiplier, **kwargs)


class QuantizedConv2D(QConv2DBase):
    """n-bit weight Conv2D layer (channels_last)."""
    WEIGHT_KIND = "quantized"

    def __init__(self, filters, kernel_regularizer=None, activity_regularizer=None, kernel_lr_multiplier='Glorot',
                 bias_lr_multiplier=None, H=1., nb=16, **kwargs):
        super().__init__(filters, kernel_regularizer=kernel_regularizer, activity_regularizer=activity_regularizer,
                         kernel_lr_multiplier=kernel_lr_multiplier, bias_lr_multiplier=bias_lr_multiplier,
                         H=H, nb=nb, **kwargs)


# Aliases
QuantizedConvolution2D = QuantizedConv2D
