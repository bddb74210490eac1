"""``Conv2D`` / ``Dense`` -- the plain Keras layers ``build_model`` uses for ``network_type == 'float'``
(models/model_factory.py:24-27): no weight quantiser, fp32 kernel x fp32 activations.  They share the
constructor surface, weight order and fused-plan hooks of the six custom layers; the kernel is re-laid out once
(QNNB_W_FLOAT / QNNB_WFMT_F32) and runs on the FFMA kernels."""
import numpy as np

from ._base import QConv2DBase, QDenseBase
from .. import engine
from ..engine import F32


class _GlorotInit:
    """keras glorot_uniform: U(-l, l), l = sqrt(6 / (fan_in + fan_out)); no clipping constraint."""

    def _create_weights(self, kernel_shape, nout):
        receptive = int(np.prod(kernel_shape[:-2])) if len(kernel_shape) > 2 else 1
        fan_in, fan_out = kernel_shape[-2] * receptive, kernel_shape[-1] * receptive
        lim = float(np.sqrt(6.0 / (fan_in + fan_out)))
        self.kernel = engine._RNG.uniform(-lim, lim, size=kernel_shape).astype(F32)
        self.bias = np.zeros((nout,), F32) if self.use_bias else None

    def _resolve_glorot(self, nb_input, nb_output):       # H / lr multipliers do not exist on the plain layers
        pass

    def _common_config(self):
        return {}


class Conv2D(_GlorotInit, QConv2DBase):
    WEIGHT_KIND = "float"


class Dense(_GlorotInit, QDenseBase):
    WEIGHT_KIND = "float"


Convolution2D = Conv2D
