"""``BinaryConv2D`` / ``BinaryDense`` -- drop-in for the reference's ``layers/binary_layers.py``
(constructors :36-42 and :100-107, ``call`` :78-85 / :160-187): ``conv2d(x, binarize(kernel, H)) + bias``.
With bit-packed +-1 inputs the product is XNOR + popcount (zero padding contributes 0)."""
from ._base import Clip, QConv2DBase, QDenseBase
from .binary_ops import binarize  # noqa: F401


class BinaryDense(QDenseBase):
    WEIGHT_KIND = "binary"

    def __init__(self, units, H=1., kernel_lr_multiplier='Glorot', bias_lr_multiplier=None, **kwargs):
        super().__init__(units, H=H, nb=1, kernel_lr_multiplier=kernel_lr_multiplier,
                         bias_lr_multiplier=bias_lr_multiplier, **kwargs)


class BinaryConv2D(QConv2DBase):
    WEIGHT_KIND = "binary"

    def __init__(self, filters, kernel_regularizer=None, activity_regularizer=None, kernel_lr_multiplier='Glorot',
                 bias_lr_multiplier=None, H=1., **kwargs):
        super().__init__(filters, kernel_regularizer=kernel_regularizer, activity_regularizer=activity_regularizer,
                         kernel_lr_multiplier=kernel_lr_multiplier, bias_lr_multiplier=bias_lr_multiplier,
                         H=H, nb=1, **kwargs)


# Aliases
BinaryConvolution2D = BinaryConv2D
