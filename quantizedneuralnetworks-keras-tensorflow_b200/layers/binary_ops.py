"""BinaryNet ops -- same names as the reference's ``layers/binary_ops.py``, evaluated by
libqnnb200 CUDA kernels.  ``binary_tanh(x) = +1 iff x > 2^-24`` in fp32 (0 maps to -1), which is
what ``2*round(clip(0.5x+0.5, 0, 1)) - 1`` evaluates to (binary_ops.py:16-24, 37-51)."""
from __future__ import annotations

from ..engine import ActProbe


def _k():
    from .. import kernels
    return kernels


def round_through(x):
    return _k().round_half_even(x)


def binary_tanh(x):
    if isinstance(x, ActProbe):
        return ActProbe(("binary",))
    return _k().sign_act(x).to_float()


def binary_sigmoid(x):
    """round(hard_sigmoid(x)) in {0, 1}  (binary_ops.py:27-34)."""
    return (binary_tanh(x) + 1.0) * 0.5


def binarize(W, H=1):
    """``H * binary_tanh(W / H)``  (binary_ops.py:54-64)."""
    H = float(H)
    return binary_tanh(W / H if H != 1.0 else W) * H


def xnorize(*a, **k):
    raise NotImplementedError("xnorize is unused by the reference models (SURVEY.md section 2 row 3)")
