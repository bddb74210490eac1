"""B200-native (sm_100a) inference path for the quantized / binary / ternary layers of
victorjoos/QuantizedNeuralNetworks-Keras-Tensorflow.

Package layout mirrors the reference (``layers/``, ``models/``); compute goes through
``libqnnb200.so`` (hand-written CUDA, C ABI in ``include/qnnb200.h``).  There is no CPU path.

The directory name contains a hyphen, so import it through the ``qnn_b200`` alias module at the
repository root (``import qnn_b200 as q; q.models.model_factory.build_model(cf)``) or with
``importlib.import_module("quantizedneuralnetworks-keras-tensorflow_b200")``.
"""
from . import _lib                     # noqa: F401  (ctypes binding; loads lazily)
from . import engine                   # noqa: F401
from .engine import (Sequential, Model, Input, BatchNormalization, MaxPooling2D, AveragePooling2D, Flatten,  # noqa: F401
                     Activation, LeakyReLU, ZeroPadding2D, Lambda, Add, add, set_seed, reset_names)
from . import layers, models           # noqa: F401
from .models.model_factory import build_model  # noqa: F401

__version__ = "0.1.0"


def lib_path():
    return _lib.LIB_PATH
