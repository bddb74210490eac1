"""Ternary weight ops -- same names as the reference's ``layers/ternary_ops.py``.

``_ternarize`` (ternary_ops.py:15-30): ``W/H``; ``cutoff = 0.7*mean|W/H|`` over the whole tensor;
``+1 if W > cutoff; -1 if W <= -cutoff; else 0``; times ``H``.  ``ternarize`` (:33-41) has the same
forward value.  ``ternary_tanh`` (:52-54) uses a whole-batch statistic and is outside the
accelerated path."""
from __future__ import annotations

import torch

from .. import _lib as L


def _ternarize(W, H=1):
    from .. import kernels as K
    flat = W.contiguous().reshape(-1, 1)                       # (cin = numel, cout = 1)
    lv = K.pack_weights(flat, L.W_TERNARY, 2, float(H), L.WFMT_I8)   # int8 [1][1][1][numel_pad]
    lv = lv.reshape(-1)[: W.numel()].reshape(W.shape)
    q = K.QTensor("i8", lv.contiguous(), float(H), int(W.shape[-1]))
    return q.to_float()


def ternarize(W, H=1):
    return _ternarize(W, H)


def ternarize_dot(*a, **k):
    raise NotImplementedError("ternarize_dot is unused by the reference models")


def ternary_tanh(x):
    raise NotImplementedError("ternary_tanh / full-tnn uses a whole-batch mean and is outside the accelerated path "
                              "(SURVEY.md section 2 row 5)")
