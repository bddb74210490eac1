import numpy as np
import torch
import torch.nn.functional as F


def backend():
    return "tensorflow"


def _t(x):
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(np.asarray(x, dtype=np.float32))


def round(x):
    return torch.round(_t(x))          # half-to-even, like tf.round


def clip(x, lo, hi):
    return torch.clamp(_t(x), float(lo), float(hi))


def stop_gradient(x):
    return x


def abs(x):
    return torch.abs(_t(x))


def mean(x, axis=None, keepdims=False):
    x = _t(x)
    return torch.mean(x) if axis is None else torch.mean(x, dim=axis, keepdim=keepdims)


def ones_like(x):
    return torch.ones_like(_t(x))


def zeros_like(x):
    return torch.zeros_like(_t(x))


def same_pads(size, k, stride):
    out = -(-size // stride)
    total = max((out - 1) * stride + k - size, 0)
    return total // 2, total - total // 2


def conv2d(x, kernel, strides=(1, 1), padding="valid", data_format=None, dilation_rate=(1, 1)):
    assert data_format in (None, "channels_last") and tuple(dilation_rate) == (1, 1)
    x, kernel = _t(x), _t(kernel)
    sh, sw = strides
    kh, kw = kernel.shape[0], kernel.shape[1]
    xc = x.permute(0, 3, 1, 2)
    if padding == "same":
        pt, pb = same_pads(x.shape[1], kh, sh)
        pl, pr = same_pads(x.shape[2], kw, sw)
        xc = F.pad(xc, (pl, pr, pt, pb))
    y = F.conv2d(xc, kernel.permute(3, 2, 0, 1).contiguous(), stride=(sh, sw))
    return y.permute(0, 2, 3, 1).contiguous()


def dot(x, y):
    return _t(x) @ _t(y)


def bias_add(x, bias, data_format=None):
    return _t(x) + _t(bias)


def softmax(x):
    return torch.softmax(_t(x), dim=-1)
