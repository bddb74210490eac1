"""``build_model(cf)`` -- the reference's ``network_type`` switch (models/model_factory.py:18-72).

``cf`` is any attribute bag with the fields the reference reads: ``network_type, wbits, abits,
architecture, dataset, dim, channels, classes, nla, nfa, nlb, nfb, nlc, nfc, nres, pfilt,
kernel_initializer, kernel_regularizer`` (config/config.py, test_resnet.py:48-62).

Differences from the reference, all deliberate:
* ``Fc`` accepts ``units`` positionally: models/vgg.py:41 calls ``Fc(cf.classes)``, which the
  reference's ``lambda **kwargs`` (model_factory.py:31) cannot take (SURVEY.md finding 6);
* ``full-tnn`` raises ``NotImplementedError``: ``ternary_tanh`` thresholds on a whole-BATCH mean
  (ternary_ops.py:52-54), which no batch-sharded or chunked forward can reproduce;
* ``model.summary()`` is only printed when ``cf.verbose`` is truthy.
"""
from ..engine import Activation, LeakyReLU
from ..layers.quantized_layers import QuantizedConv2D, QuantizedDense
from ..layers.quantized_ops import quantized_tanh as quantize_op
from ..layers.binary_layers import BinaryConv2D, BinaryDense
from ..layers.binary_ops import binary_tanh
from ..layers.ternary_layers import TernaryConv2D, TernaryDense
from ..layers.float_layers import Conv2D, Dense
from .resnet import ResNet18
from .vgg import Vgg

NETWORK_TYPES = ('float', 'qnn', 'full-qnn', 'bnn', 'qbnn', 'full-bnn', 'tnn', 'qtnn')


def build_model(cf, legacy_resnet=False):
    def quantized_relu(x):                      # historical name; it is quantized_tanh (model_factory.py:9,19-20)
        return quantize_op(x, nb=cf.abits)

    leaky = lambda: LeakyReLU()
    quant = lambda: Activation(quantized_relu)
    kind = cf.network_type

    if kind == 'float':
        Conv, Fc, Act = Conv2D, Dense, leaky                                 # model_factory.py:24-27
    elif kind in ('qnn', 'full-qnn'):
        Conv = lambda **kw: QuantizedConv2D(H=1, nb=cf.wbits, **kw)
        Fc = lambda *a, **kw: QuantizedDense(*a, nb=cf.abits, **kw)      # nb=abits as in the reference
        Act = leaky if kind == 'qnn' else quant
    elif kind in ('bnn', 'qbnn', 'full-bnn'):
        Conv = lambda **kw: BinaryConv2D(H=1, **kw)
        Fc = BinaryDense
        Act = {'bnn': leaky, 'qbnn': quant, 'full-bnn': lambda: Activation(binary_tanh)}[kind]
    elif kind in ('tnn', 'qtnn'):
        Conv = lambda **kw: TernaryConv2D(H=1, **kw)
        Fc = TernaryDense
        Act = leaky if kind == 'tnn' else quant
    elif kind == 'full-tnn':
        raise NotImplementedError("network_type %r is outside the accelerated path (batch-global ternary_tanh); "
                                  "supported: %s" % (kind, ", ".join(NETWORK_TYPES)))
    else:
        raise ValueError('wrong network type, the supported network types in this repo are float, qnn, full-qnn, bnn and full-bnn')

    if cf.architecture == "VGG":
        model = Vgg(Conv, Act, Fc, cf)
    elif cf.architecture == "RESNET":
        model = ResNet18(Conv, Act, Fc, cf, legacy=legacy_resnet)
    else:
        raise ValueError("Error: type " + str(cf.architecture) + " is not supported")

    if getattr(cf, "verbose", False):
        model.summary()
    return model
