"""``TernaryConv2D`` / ``TernaryDense`` -- drop-in for the reference's ``layers/ternary_layers.py``
(constructors :37-41 and :100-105, ``call`` :77-84 / :156-174): ``conv2d(x, ternarize(kernel, H)) + bias``
with the whole-tensor cutoff ``0.7*mean|W/H|`` (layers/ternary_ops.py:22-28)."""
from ._base import Clip, QConv2DBase, QDenseBase
from .ternary_ops import ternarize  # noqa: F401


class TernaryDense(QDenseBase):
    WEIGHT_KIND = "ternary"

    def __init__(self, units, H=1., kernel_lr_multiplier='Glorot', bias_lr_multiplier=None, **kwargs):
        super().__init__(units, H=H, nb=2, kernel_lr_multiplier=kernel_lr_multiplier,
                         bias_lr_multiplier=bias_lr_multiplier, **kwargs)


class TernaryConv2D(QConv2DBase):
    WEIGHT_KIND = "ternary"

    def __init__(self, filters, kernel_lr_multiplier='Glorot', bias_lr_multiplier=None, H=1., **kwargs):
        super().__init__(filters, kernel_lr_multiplier=kernel_lr_multiplier, bias_lr_multiplier=bias_lr_multiplier,
                         H=H, nb=2, **kwargs)


# Aliases
TernaryConvolution2D = TernaryConv2D
