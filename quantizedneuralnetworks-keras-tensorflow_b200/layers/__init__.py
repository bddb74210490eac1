"""Custom layers and quantiser ops, named as in the reference's ``layers/`` package."""
from . import quantized_ops, binary_ops, ternary_ops          # noqa: F401
from . import quantized_layers, binary_layers, ternary_layers  # noqa: F401
